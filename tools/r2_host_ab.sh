#!/bin/bash
# caar_run_host: one copy-in stream against two (CAAR_HOST_IN_STREAMS), ne=120 on one GPU
OUT=gpurun_out/r2s_host_ab.log
: > $OUT
for rep in 1 2; do
for S in 1 2; do
  CAAR_HOST_IN_STREAMS=$S timeout 400 python tools/kernel_sweep.py --nelem 86400 --nlev 72 --steps 5 --repeat 1 --host-steps 3 --chunk 0 --tag in$S >> $OUT 2>&1
done
done
CAAR_HOST_IN_STREAMS=2 CAAR_HOST_ROW=18432 timeout 400 python tools/kernel_sweep.py --nelem 86400 --nlev 72 --steps 5 --repeat 1 --host-steps 3 --chunk 0 --tag in2_row18432 >> $OUT 2>&1
CAAR_HOST_IN_STREAMS=2 timeout 400 python tools/kernel_sweep.py --nelem 86400 --nlev 72 --steps 5 --repeat 1 --host-steps 3 --chunk 900 3600 --tag in2_chunks >> $OUT 2>&1
timeout 300 python -m pytest tests/test_host_gpu.py tests/test_parity_gpu.py -q -m gpu -x -k "host" > gpurun_out/r2s_host_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2s_host_pytest.log
CAAR_HOST_IN_STREAMS=2 timeout 300 python -m pytest tests/test_host_gpu.py tests/test_parity_gpu.py -q -m gpu -x -k "host" > gpurun_out/r2s_host2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2s_host2_pytest.log
