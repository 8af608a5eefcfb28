#!/bin/bash
# Final evidence of the round for the kernels as committed: launch list of the bench command, ncu --set full of the
# headline kernel (ne=120 quarter: 21600 elements) and of the level-local operators, saxpby sweep, ne=30 / ne=256 lines.
set -u
OUT=gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
CMD="python bench.py --steps 2 --warmup 3 --nelem 21600 --no-e2e --no-cpu-baseline --no-parity --no-clock-topup"
$CMD > $OUT/r2j_plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/r2j_launches.csv $CMD > $OUT/r2j_ncu_launches.log 2>&1
$CMD >> $OUT/r2j_plain_bench.log 2>&1 && \
$NCU -k regex:caar_fused_kernel -s 3 -c 1 -f -o $OUT/r2j_fused_L72 $CMD > $OUT/r2j_ncu_fused.log 2>&1
C2="python tools/levelop_bench.py --nelem 21600 --nlev 72 --qsize 4 --ops euler,divwk,lap,lapt --modes fast --steps 2"
$C2 > $OUT/r2j_plain_levelop.log 2>&1 && \
$NCU -k regex:levelop_kernel -c 16 -f -o $OUT/r2j_levelops $C2 > $OUT/r2j_ncu_levelops.log 2>&1
python tools/saxpby_sweep.py --out $OUT/r2j_saxpby_sweep.json > $OUT/r2j_saxpby.log 2>&1
python bench.py > $OUT/r2j_bench_n1.json 2> $OUT/r2j_bench_n1.err
python bench.py --nelem 5400 --steps 100 --no-cpu-baseline > $OUT/r2j_bench_ne30.json 2> $OUT/r2j_bench_ne30.err
python bench.py --nelem 49152 --nlev 128 --steps 10 --no-cpu-baseline > $OUT/r2j_bench_ne256slice.json 2> $OUT/r2j_bench_ne256slice.err
python bench.py --nelem 10800 --steps 20 --no-cpu-baseline --no-e2e > $OUT/r2j_bench_strong_slice.json 2> $OUT/r2j_bench_strong_slice.err
python bench.py --nelem 100000 --nlev 30 --steps 20 --no-cpu-baseline > $OUT/r2j_bench_nlev30.json 2> $OUT/r2j_bench_nlev30.err
