#!/bin/bash
# The controls again in another order, with the SM clock read after each timed loop (is the order / clock state part
# of the differences in r2t_time_levels.jsonl?)
set -u
OUT=gpurun_out
timeout 300 python tools/kernel_sweep.py --nelem 86400 --nlev 72 --steps 20 --variants dry aliased distinct aliased_dry dry distinct aliased --tag ne120_reordered > $OUT/r2t_time_levels_reordered.jsonl 2> $OUT/r2t_aliased2.err
