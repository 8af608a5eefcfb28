#!/bin/bash
# thread-per-level laplacians, dynamic chunks with a static first chunk: parity, sizes; the other operators at other sizes
set -u
OUT=gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -q -m gpu -k "weak_form or biharmonic or linear" > $OUT/lf7_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/lf7_pytest.log
: > $OUT/lf7_bench.jsonl
run() { echo "# $*" >> $OUT/lf7_bench.jsonl; timeout 300 "$@" >> $OUT/lf7_bench.jsonl 2>> $OUT/lf7_bench.err; }
echo "# default chunk" >> $OUT/lf7_bench.jsonl
run python tools/levelop_bench.py --ops lap,lapt --modes fast --nlev 72
run python tools/levelop_bench.py --ops lap,lapt --modes fast --nlev 128
run python tools/levelop_bench.py --ops lap,lapt --modes fast --nlev 30
run python tools/levelop_bench.py --ops lap,lapt --modes fast --nlev 26
run python tools/levelop_bench.py --ops lap,lapt --modes fast --nlev 72 --nelem 5400
run python tools/levelop_bench.py --ops lap,lapt --modes fast --nlev 72 --nelem 86400
run python tools/levelop_bench.py --ops lap,lapt --modes fast --nlev 128 --nelem 49152
for C in 2 3 4 6; do
  export CAAR_LAPLACE_CHUNK=$C
  echo "# chunk=$C" >> $OUT/lf7_bench.jsonl
  run python tools/levelop_bench.py --ops lap --modes fast --nlev 72 --nelem 5400
  run python tools/levelop_bench.py --ops lap --modes fast --nlev 72
  run python tools/levelop_bench.py --ops lap --modes fast --nlev 72 --nelem 86400
done
unset CAAR_LAPLACE_CHUNK
for W in 0 1 2 4 8 16; do
  if [ $W -gt 0 ]; then export CAAR_LEVELOP_WAVES=$W; fi
  echo "# first-generation skeleton, waves=$W (0 = default)" >> $OUT/lf7_bench.jsonl
  run python tools/levelop_bench.py --ops euler,divwk --modes fast --nlev 72 --nelem 5400
  run python tools/levelop_bench.py --ops euler,divwk --modes fast --nlev 72 --nelem 86400
done
unset CAAR_LEVELOP_WAVES
timeout 600 ncu --set full --clock-control none --import-source on -k regex:laplace_flat --launch-skip 3 --launch-count 1 \
  -o $OUT/r2g_lapflat6_L72 -f python tools/levelop_bench.py --ops lap --modes fast --nlev 72 --steps 2 > $OUT/lf7_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:laplace_flat --launch-skip 3 --launch-count 1 \
  -o $OUT/r2g_lapflat6t_L128 -f python tools/levelop_bench.py --ops lapt --modes fast --nlev 128 --steps 2 >> $OUT/lf7_ncu.log 2>&1
