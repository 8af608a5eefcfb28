#!/bin/bash
# ncu --set full of the fused kernel under the aliased and the dry control (21600 x 72), next to the distinct-control
# capture in profiles/r2r_fused_ncu_summary.txt
set -u
OUT=gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
for V in aliased dry; do
  timeout 300 $NCU -k regex:caar_fused_kernel -s 3 -c 1 -f -o $OUT/r2t_fused_L72_$V \
    python tools/kernel_sweep.py --nelem 21600 --nlev 72 --variants $V --steps 3 --warmup 2 --repeat 1 > $OUT/r2t_ncu_$V.log 2>&1
done
