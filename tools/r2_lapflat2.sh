#!/bin/bash
# thread-per-level laplacians: where the time goes (copy-only pipeline, larger arrays, ncu)
set -u
OUT=gpurun_out
: > $OUT/lf2_bench.jsonl
run() { echo "# $*" >> $OUT/lf2_bench.jsonl; timeout 300 "$@" >> $OUT/lf2_bench.jsonl 2>> $OUT/lf2_bench.err; }
run python tools/levelop_bench.py --ops lap,lapt --modes fast --nlev 128 --lib tools/_variants/libcaar_b200_lf_copy.so
run python tools/levelop_bench.py --ops lap,lapt --modes fast --nlev 72 --lib tools/_variants/libcaar_b200_lf_copy.so
run python tools/levelop_bench.py --ops lap,lapt,divwk --modes fast --nlev 128 --nelem 49152
run python tools/levelop_bench.py --ops lap --modes fast --nlev 128 --nelem 49152 --lib tools/_variants/libcaar_b200_lf_copy.so
run python tools/levelop_bench.py --ops lap,lapt,divwk --modes fast --nlev 72 --nelem 86400
timeout 600 ncu --set full --clock-control none --import-source on -k regex:laplace_flat --launch-skip 3 --launch-count 1 \
  -o $OUT/r2g_lapflat_L128 -f python tools/levelop_bench.py --ops lap --modes fast --nlev 128 --steps 2 > $OUT/lf2_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:laplace_flat --launch-skip 3 --launch-count 1 \
  -o $OUT/r2g_lapflat_copy_L128 -f python tools/levelop_bench.py --ops lap --modes fast --nlev 128 --steps 2 --lib tools/_variants/libcaar_b200_lf_copy.so >> $OUT/lf2_ncu.log 2>&1
