#!/bin/bash
# 2 GPUs: the multi-GPU tests, a default-size 2-rank bench (weak) and the C++ driver on 2 GPUs
set -u
OUT=gpurun_out
timeout 1200 python -m pytest tests/test_multigpu_gpu.py tests/test_host_gpu.py -m gpu -x -q > $OUT/r2i_pytest_2gpu.log 2>&1
echo "pytest rc=$?" >> $OUT/r2i_pytest_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > $OUT/r2i_bench_n2.json 2> $OUT/r2i_bench_n2.err
echo "rc=$?" >> $OUT/r2i_bench_n2.err
