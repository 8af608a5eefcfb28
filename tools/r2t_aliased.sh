#!/bin/bash
# Time-level aliasing and the dry case as roofline records: n0 = np1 = nm1 reads the aliased level
# once (17 compulsory level-fields instead of 21), qn0 = -1 reads no Qdp (20). One process per size, every variant on
# the same resident state; then the DRAM bytes of one aliased launch (ncu, targeted metrics).
set -u
OUT=gpurun_out
J=$OUT/r2t_aliased.jsonl
: > $J
V="distinct aliased dry aliased_dry"
timeout 300 python tools/kernel_sweep.py --nelem 86400 --nlev 72 --steps 20 --variants $V --tag ne120 >> $J 2>> $OUT/r2t_aliased.err
timeout 200 python tools/kernel_sweep.py --nelem 12288 --nlev 128 --steps 20 --variants $V --tag L128 >> $J 2>> $OUT/r2t_aliased.err
timeout 200 python tools/kernel_sweep.py --nelem 21600 --nlev 72 --steps 20 --eulerian --variants $V --tag eul >> $J 2>> $OUT/r2t_aliased.err
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,gpu__time_duration.sum --clock-control none \
  -k regex:caar_fused_kernel -s 3 -c 1 --csv --log-file $OUT/r2t_aliased_ncu.csv \
  python tools/kernel_sweep.py --nelem 21600 --nlev 72 --variants aliased --steps 3 --warmup 2 --repeat 1 > $OUT/r2t_aliased_ncu.log 2>&1
echo done >> $OUT/r2t_aliased.err
