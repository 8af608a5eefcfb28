#!/bin/bash
# After the DMMA change: ncu --set full of the fused kernel (Lagrangian and Eulerian nlev 72), the default bench line, the
# launch list of a bench run.
set -u
OUT=gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
timeout 80 $NCU -k regex:caar_fused_kernel -s 2 -c 1 -f -o $OUT/r2v_fused_L72 python tools/kernel_sweep.py --nelem 21600 --nlev 72 --steps 3 --warmup 2 --repeat 1 > $OUT/r2v_ncu.log 2>&1
timeout 140 python bench.py > $OUT/r2v_bench.json 2> $OUT/r2v_bench.err; echo "rc=$?" >> $OUT/r2v_bench.err
timeout 80 $NCU -k regex:caar_fused_kernel -s 2 -c 1 -f -o $OUT/r2v_eul_L72 python tools/kernel_sweep.py --nelem 21600 --nlev 72 --eulerian --steps 3 --warmup 2 --repeat 1 >> $OUT/r2v_ncu.log 2>&1
timeout 60 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/r2v_final_launches.csv python bench.py --steps 2 --warmup 3 --nelem 21600 --no-e2e --no-cpu-baseline --no-parity --no-clock-topup >> $OUT/r2v_ncu.log 2>&1
