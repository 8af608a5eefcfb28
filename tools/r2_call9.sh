#!/bin/bash
set -u
OUT=gpurun_out
{
for lib in default ns6 ns8no3 w8 w2 mb5; do
  L=""; [ $lib != default ] && L="--lib tools/_variants/libcaar_b200_$lib.so"
  for w in 1 2 4 8; do
    CAAR_LEVELOP_WAVES=$w timeout 300 python tools/levelop_bench.py --nelem 21600 --nlev 72 --qsize 4 --ops euler,lap --modes fast $L
  done
  CAAR_LEVELOP_WAVES=4 timeout 300 python tools/levelop_bench.py --nelem 2700 --nlev 72 --qsize 35 --ops euler --modes fast $L
done
} > $OUT/r2g_levelop_sweep.log 2>&1
