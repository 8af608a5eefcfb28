#!/bin/bash
# thread-per-level laplacians, two-stage geometry prefetch + waves heuristic
set -u
OUT=gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -q -m gpu -k "weak_form or biharmonic or linear" > $OUT/lf5_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/lf5_pytest.log
: > $OUT/lf5_bench.jsonl
run() { echo "# $*" >> $OUT/lf5_bench.jsonl; timeout 300 "$@" >> $OUT/lf5_bench.jsonl 2>> $OUT/lf5_bench.err; }
echo "# default waves" >> $OUT/lf5_bench.jsonl
run python tools/levelop_bench.py --ops lap,lapt --modes fast --nlev 72
run python tools/levelop_bench.py --ops lap,lapt --modes fast --nlev 128
run python tools/levelop_bench.py --ops lap,lapt --modes fast --nlev 30
run python tools/levelop_bench.py --ops lap,lapt --modes fast --nlev 72 --nelem 5400
run python tools/levelop_bench.py --ops lap,lapt --modes fast --nlev 72 --nelem 86400
run python tools/levelop_bench.py --ops lap,lapt --modes fast --nlev 128 --nelem 49152
for W in 1 2 3 4 6 8 12 16; do
  export CAAR_LEVELOP_WAVES=$W
  echo "# waves=$W" >> $OUT/lf5_bench.jsonl
  run python tools/levelop_bench.py --ops lap --modes fast --nlev 72
  run python tools/levelop_bench.py --ops lap --modes fast --nlev 128 --nelem 49152
done
unset CAAR_LEVELOP_WAVES
timeout 600 ncu --set full --clock-control none --import-source on -k regex:laplace_flat --launch-skip 3 --launch-count 1 \
  -o $OUT/r2g_lapflat4_L128 -f python tools/levelop_bench.py --ops lap --modes fast --nlev 128 --steps 2 > $OUT/lf5_ncu.log 2>&1
