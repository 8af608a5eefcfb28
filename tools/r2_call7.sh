#!/bin/bash
set -u
OUT=gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
python tools/levelop_bench.py --nelem 21600 --nlev 72 --qsize 4 --ops euler --modes fast --steps 2 > $OUT/r2f_plain.log 2>&1 && \
$NCU -k regex:levelop_kernel -s 1 -c 1 -f -o $OUT/r2f_euler_q4 python tools/levelop_bench.py --nelem 21600 --nlev 72 --qsize 4 --ops euler --modes fast --steps 2 > $OUT/r2f_ncu1.log 2>&1
python tools/levelop_bench.py --nelem 43200 --nlev 72 --ops lap --modes fast --steps 2 >> $OUT/r2f_plain.log 2>&1 && \
$NCU -k regex:levelop_kernel -s 1 -c 1 -f -o $OUT/r2f_lap python tools/levelop_bench.py --nelem 43200 --nlev 72 --ops lap --modes fast --steps 2 > $OUT/r2f_ncu2.log 2>&1
