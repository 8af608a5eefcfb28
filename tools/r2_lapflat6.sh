#!/bin/bash
# thread-per-level laplacians with the dynamic chunk scheduler
set -u
OUT=gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -q -m gpu -k "weak_form or biharmonic or linear" > $OUT/lf6_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/lf6_pytest.log
: > $OUT/lf6_bench.jsonl
run() { echo "# $*" >> $OUT/lf6_bench.jsonl; timeout 300 "$@" >> $OUT/lf6_bench.jsonl 2>> $OUT/lf6_bench.err; }
echo "# default chunk" >> $OUT/lf6_bench.jsonl
run python tools/levelop_bench.py --ops lap,lapt --modes fast --nlev 72
run python tools/levelop_bench.py --ops lap,lapt --modes fast --nlev 128
run python tools/levelop_bench.py --ops lap,lapt --modes fast --nlev 30
run python tools/levelop_bench.py --ops lap,lapt --modes fast --nlev 26
run python tools/levelop_bench.py --ops lap,lapt --modes fast --nlev 72 --nelem 5400
run python tools/levelop_bench.py --ops lap,lapt --modes fast --nlev 72 --nelem 86400
run python tools/levelop_bench.py --ops lap,lapt --modes fast --nlev 128 --nelem 49152
for C in 1 2 4 8 16 32; do
  export CAAR_LAPLACE_CHUNK=$C
  echo "# chunk=$C" >> $OUT/lf6_bench.jsonl
  run python tools/levelop_bench.py --ops lap --modes fast --nlev 72 --nelem 5400
  run python tools/levelop_bench.py --ops lap --modes fast --nlev 72
  run python tools/levelop_bench.py --ops lap --modes fast --nlev 128 --nelem 49152
done
unset CAAR_LAPLACE_CHUNK
timeout 600 ncu --set full --clock-control none --import-source on -k regex:laplace_flat --launch-skip 3 --launch-count 1 \
  -o $OUT/r2g_lapflat5_L128 -f python tools/levelop_bench.py --ops lap --modes fast --nlev 128 --steps 2 > $OUT/lf6_ncu.log 2>&1
