#!/bin/bash
# What the driver runs at round end, on one GPU: the -m gpu suite, smoke(), the default bench line and the reference arm.
set -u
OUT=gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > $OUT/verify_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/verify_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/verify_smoke.log 2>&1; echo "smoke rc=$?" >> $OUT/verify_smoke.log
timeout 900 python bench.py > $OUT/verify_bench.json 2> $OUT/verify_bench.err; echo "rc=$?" >> $OUT/verify_bench.err
timeout 900 python bench.py --impl reference > $OUT/verify_bench_ref.json 2> $OUT/verify_bench_ref.err; echo "rc=$?" >> $OUT/verify_bench_ref.err
