#!/bin/bash
# 8 GPUs: BASELINE configs[3] strong scaling (ne=120 cut over 8 GPUs) and configs[4] (ne=256, nlev=128: 49152 elements per GPU)
set -u
OUT=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29521 bench.py --gpus 8 --scaling strong --nelem 86400 --steps 200 --warmup 5 --no-cpu-baseline --no-e2e > $OUT/r2n_bench_ne120_strong_n8.json 2> $OUT/r2n_bench_ne120_strong_n8.err
echo "rc=$?" >> $OUT/r2n_bench_ne120_strong_n8.err
timeout 900 $TR --master-port 29522 bench.py --gpus 8 --nelem 49152 --nlev 128 --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 2 > $OUT/r2n_bench_ne256_n8.json 2> $OUT/r2n_bench_ne256_n8.err
echo "rc=$?" >> $OUT/r2n_bench_ne256_n8.err
