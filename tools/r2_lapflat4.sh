#!/bin/bash
# thread-per-level laplacians with register prefetch + L2 prefetch cursor: parity, sizes, waves, variants, ncu
set -u
OUT=gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -q -m gpu -k "weak_form or biharmonic or linear" > $OUT/lf4_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/lf4_pytest.log
: > $OUT/lf4_bench.jsonl
run() { echo "# $*" >> $OUT/lf4_bench.jsonl; timeout 300 "$@" >> $OUT/lf4_bench.jsonl 2>> $OUT/lf4_bench.err; }
for W in 1 2 3 4 6; do
  export CAAR_LEVELOP_WAVES=$W
  echo "# waves=$W" >> $OUT/lf4_bench.jsonl
  run python tools/levelop_bench.py --ops lap,lapt --modes fast --nlev 72
  run python tools/levelop_bench.py --ops lap,lapt --modes fast --nlev 128 --nelem 49152
done
unset CAAR_LEVELOP_WAVES
run python tools/levelop_bench.py --ops lap,lapt --modes fast --nlev 128
run python tools/levelop_bench.py --ops lap,lapt --modes fast --nlev 30
run python tools/levelop_bench.py --ops lap,lapt --modes fast --nlev 72 --nelem 86400
for v in pd6 w8 a; do
  run python tools/levelop_bench.py --ops lap,lapt --modes fast --nlev 72 --lib tools/_variants/libcaar_b200_lf_$v.so
  run python tools/levelop_bench.py --ops lap,lapt --modes fast --nlev 128 --nelem 49152 --lib tools/_variants/libcaar_b200_lf_$v.so
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:laplace_flat --launch-skip 3 --launch-count 1 \
  -o $OUT/r2g_lapflat3_L128 -f python tools/levelop_bench.py --ops lap --modes fast --nlev 128 --steps 2 > $OUT/lf4_ncu.log 2>&1
