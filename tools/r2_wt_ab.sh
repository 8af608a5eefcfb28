#!/bin/bash
# warp_totals: column sums (default) against the butterfly (variant wt1), Lagrangian and Eulerian instances
OUT=gpurun_out/r2q_wt_ab.log
: > $OUT
timeout 900 python -m pytest tests/test_parity_gpu.py -q -m gpu -x -k "nlev or eulerian or Eulerian or fused or default" > gpurun_out/r2q_wt_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2q_wt_pytest.log
V=tools/_variants/libcaar_b200_wt1.so
for rep in 1 2; do
for cfg in "21600 72" "12288 128" "21600 30" "21600 88" "21600 104" "5400 72"; do
  set -- $cfg
  python tools/kernel_sweep.py --nelem $1 --nlev $2 --steps 20 --tag new_L$2_E$1 >> $OUT 2>&1
  python tools/kernel_sweep.py --nelem $1 --nlev $2 --steps 20 --tag old_L$2_E$1 --lib $V >> $OUT 2>&1
done
for cfg in "21600 72" "12288 128" "21600 96"; do
  set -- $cfg
  python tools/kernel_sweep.py --nelem $1 --nlev $2 --eulerian --steps 20 --tag new_eul_L$2 >> $OUT 2>&1
  python tools/kernel_sweep.py --nelem $1 --nlev $2 --eulerian --steps 20 --tag old_eul_L$2 --lib $V >> $OUT 2>&1
done
done
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r2q_bench_after_wt.json 2> gpurun_out/r2q_bench_after_wt.err
timeout 300 python -m pytest tests/test_host_gpu.py tests/test_parity_gpu.py -q -m gpu -x -k "host" > gpurun_out/r2q_host_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2q_host_pytest.log
