#!/bin/bash
# 4 GPUs: BASELINE configs[3] weak (86400 elements per GPU), configs[4] weak (49152 per GPU, nlev=128) and strong
# (393216 elements cut over 4 GPUs = 98304 per GPU = 32.4 GB each)
set -u
OUT=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29531 bench.py --gpus 4 --steps 20 --warmup 3 --no-cpu-baseline > $OUT/r2q_bench_ne120_weak_n4.json 2> $OUT/r2q_bench_ne120_weak_n4.err; echo "rc=$?" >> $OUT/r2q_bench_ne120_weak_n4.err
timeout 600 $TR --master-port 29532 bench.py --gpus 4 --nelem 49152 --nlev 128 --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/r2q_bench_ne256_weak_n4.json 2> $OUT/r2q_bench_ne256_weak_n4.err; echo "rc=$?" >> $OUT/r2q_bench_ne256_weak_n4.err
timeout 900 $TR --master-port 29533 bench.py --gpus 4 --scaling strong --nelem 393216 --nlev 128 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/r2q_bench_ne256_strong_n4.json 2> $OUT/r2q_bench_ne256_strong_n4.err; echo "rc=$?" >> $OUT/r2q_bench_ne256_strong_n4.err
