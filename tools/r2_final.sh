#!/bin/bash
# End of round 2: the whole -m gpu suite, smoke, the default bench line + reference arm, every fused instance once
# (Lagrangian and Eulerian), and fresh ncu captures of the kernels touched last (fused L72 / Eulerian L72 after the
# warp_totals change).
set -u
OUT=gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > $OUT/final_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/final_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/final_smoke.log 2>&1; echo "smoke rc=$?" >> $OUT/final_smoke.log
timeout 900 python bench.py > $OUT/final_bench.json 2> $OUT/final_bench.err; echo "rc=$?" >> $OUT/final_bench.err
timeout 900 python bench.py --impl reference > $OUT/final_bench_ref.json 2> $OUT/final_bench_ref.err; echo "rc=$?" >> $OUT/final_bench_ref.err
: > $OUT/final_instances.jsonl
for L in 8 16 24 32 40 48 56 64 72 80 96 112 120 128; do
  E=$(( 1555200 / L ))
  timeout 200 python tools/kernel_sweep.py --nelem $E --nlev $L --steps 20 --tag lag_L$L >> $OUT/final_instances.jsonl 2>> $OUT/final_instances.err
  timeout 200 python tools/kernel_sweep.py --nelem $E --nlev $L --eulerian --steps 20 --tag eul_L$L >> $OUT/final_instances.jsonl 2>> $OUT/final_instances.err
done
NCU="ncu --set full --clock-control none --import-source on"
timeout 400 $NCU -k regex:caar_fused_kernel -s 2 -c 1 -f -o $OUT/r2r_fused_L72 python tools/kernel_sweep.py --nelem 21600 --nlev 72 --steps 3 --warmup 2 --repeat 1 > $OUT/final_ncu.log 2>&1
timeout 400 $NCU -k regex:caar_fused_kernel -s 2 -c 1 -f -o $OUT/r2r_eul_L72 python tools/kernel_sweep.py --nelem 21600 --nlev 72 --eulerian --steps 3 --warmup 2 --repeat 1 >> $OUT/final_ncu.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/r2r_final_launches.csv python bench.py --steps 2 --warmup 3 --nelem 21600 --no-e2e --no-cpu-baseline --no-parity --no-clock-topup >> $OUT/final_ncu.log 2>&1
