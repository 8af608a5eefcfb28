#!/bin/bash
# Eulerian instances: carry + column total in one pass (default) against two calls (variant wtsep)
OUT=gpurun_out/r2q_wt_ab2.log
: > $OUT
timeout 900 python -m pytest tests/test_parity_gpu.py -q -m gpu -x -k "eulerian or Eulerian" > gpurun_out/r2q_wt2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2q_wt2_pytest.log
V=tools/_variants/libcaar_b200_wtsep.so
for rep in 1 2 3; do
for cfg in "21600 72" "12288 128" "21600 96" "21600 30"; do
  set -- $cfg
  python tools/kernel_sweep.py --nelem $1 --nlev $2 --eulerian --steps 20 --tag new_eul_L$2 >> $OUT 2>&1
  python tools/kernel_sweep.py --nelem $1 --nlev $2 --eulerian --steps 20 --tag old_eul_L$2 --lib $V >> $OUT 2>&1
done
done
