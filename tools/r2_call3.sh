#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
python bench.py > $OUT/r2c_bench_n1.json 2> $OUT/r2c_bench_n1.err; echo "rc=$?" >> $OUT/r2c_bench_n1.err
python bench.py --impl reference > $OUT/r2c_bench_ref.json 2> $OUT/r2c_bench_ref.err; echo "rc=$?" >> $OUT/r2c_bench_ref.err
python bench.py --nelem 5400 --steps 100 --no-cpu-baseline > $OUT/r2c_bench_ne30.json 2> $OUT/r2c_bench_ne30.err
python bench.py --nelem 49152 --nlev 128 --steps 10 --no-cpu-baseline > $OUT/r2c_bench_ne256slice.json 2> $OUT/r2c_bench_ne256slice.err
python bench.py --nelem 10800 --steps 20 --no-cpu-baseline --no-e2e > $OUT/r2c_bench_10800.json 2> $OUT/r2c_bench_10800.err
tools/_variants/pcie_probe_multi > $OUT/r2c_pcie_multi_1gpu.txt 2>&1
