#!/bin/bash
# Round-2 ncu evidence for every kernel that is NOT the headline one: plain run first (must exit
# 0), then one `ncu --set full` capture of one launch each. Run under gpurun, one GPU:
#   gpurun --timeout 1500 -- 'bash tools/r2_profile_secondary.sh [tag]'
# Outputs: gpurun_out/<tag>_*.ncu-rep, <tag>_plain.log. Summaries are made here with tools/ncu_summarize.sh.
set -u
TAG=${1:-r2}
OUT=gpurun_out
mkdir -p $OUT
NCU="ncu --set full --clock-control none --import-source on"
PLAIN=$OUT/${TAG}_plain.log
: > $PLAIN

cap() {  # cap <name> <kernel regex> <skip> -- <cmd...>
  local name=$1 regex=$2 skip=$3
  shift 4
  echo "=== $name: $*" >> $PLAIN
  if "$@" >> $PLAIN 2>&1; then
    $NCU -k regex:$regex -s $skip -c 1 -f -o $OUT/${TAG}_$name "$@" > $OUT/${TAG}_${name}_ncu.log 2>&1 || echo "ncu failed for $name" >> $PLAIN
  else
    echo "plain run FAILED for $name" >> $PLAIN
  fi
}

cap euler_q1  euler_step_kernel 0 -- python tools/euler_bench.py --nelem 43200 --nlev 72 --qsize 1 --steps 3
cap euler_q4  euler_step_kernel 0 -- python tools/euler_bench.py --nelem 21600 --nlev 72 --qsize 4 --steps 3
cap euler_q35 euler_step_kernel 0 -- python tools/euler_bench.py --nelem 2700 --nlev 72 --qsize 35 --steps 3
cap eul_L72   caar_fused_kernel 2 -- python tools/kernel_sweep.py --nelem 21600 --nlev 72 --eulerian --steps 3 --warmup 2 --repeat 1
cap eul_L128  caar_fused_kernel 2 -- python tools/kernel_sweep.py --nelem 12288 --nlev 128 --eulerian --steps 3 --warmup 2 --repeat 1
cap strict    caar_strict_kernel 1 -- python tools/kernel_sweep.py --nelem 5400 --nlev 72 --mode strict --steps 3 --warmup 2 --repeat 1
cap saxpby    saxpby_kernel 3 -- python tools/saxpby_sweep.py --out $OUT/${TAG}_saxpby_1g.json --min-gb 1 --max-gb 1 --no-cpu
echo done >> $PLAIN
