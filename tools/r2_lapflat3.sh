#!/bin/bash
# thread-per-level laplacians with the geometry prefetch: parity, throughput at three sizes, grid in waves, one variant
set -u
OUT=gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -q -m gpu -k "weak_form or biharmonic or linear" > $OUT/lf3_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/lf3_pytest.log
: > $OUT/lf3_bench.jsonl
run() { echo "# $*" >> $OUT/lf3_bench.jsonl; timeout 300 "$@" >> $OUT/lf3_bench.jsonl 2>> $OUT/lf3_bench.err; }
run python tools/levelop_bench.py --ops lap,lapt --modes fast --nlev 72
run python tools/levelop_bench.py --ops lap,lapt --modes fast --nlev 128
run python tools/levelop_bench.py --ops lap,lapt --modes fast --nlev 30
run python tools/levelop_bench.py --ops lap,lapt --modes fast --nlev 26
run python tools/levelop_bench.py --ops lap,lapt --modes fast --nlev 72 --nelem 86400
run python tools/levelop_bench.py --ops lap,lapt --modes fast --nlev 128 --nelem 49152
for W in 2 3; do
  CAAR_LEVELOP_WAVES=$W run python tools/levelop_bench.py --ops lap,lapt --modes fast --nlev 128 --nelem 49152
  echo "# ^ waves=$W" >> $OUT/lf3_bench.jsonl
done
for L in 72 128; do
  run python tools/levelop_bench.py --ops lap,lapt --modes fast --nlev $L --lib tools/_variants/libcaar_b200_lf_a.so
done
run python tools/levelop_bench.py --ops lap,lapt --modes fast --nlev 128 --nelem 49152 --lib tools/_variants/libcaar_b200_lf_a.so
run python tools/levelop_bench.py --ops lap --modes fast --nlev 128 --nelem 49152 --lib tools/_variants/libcaar_b200_lf_copy.so
timeout 600 ncu --set full --clock-control none --import-source on -k regex:laplace_flat --launch-skip 3 --launch-count 1 \
  -o $OUT/r2g_lapflat2_L128 -f python tools/levelop_bench.py --ops lap --modes fast --nlev 128 --steps 2 > $OUT/lf3_ncu.log 2>&1
