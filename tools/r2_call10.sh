#!/bin/bash
set -u
OUT=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/r2h_pytest.log 2>&1
echo "pytest rc=$?" >> $OUT/r2h_pytest.log
{
for q in 1 4 8; do timeout 300 python tools/levelop_bench.py --nelem 21600 --nlev 72 --qsize $q --ops euler --modes fast; done
timeout 300 python tools/levelop_bench.py --nelem 2700 --nlev 72 --qsize 35 --ops euler --modes fast
timeout 300 python tools/levelop_bench.py --nelem 10800 --nlev 128 --qsize 8 --ops euler --modes fast
timeout 300 python tools/levelop_bench.py --nelem 43200 --nlev 72 --qsize 1 --ops divwk,lap,lapt --modes fast
timeout 300 python tools/levelop_bench.py --nelem 24000 --nlev 128 --qsize 1 --ops divwk,lap,lapt --modes fast
timeout 300 python tools/levelop_bench.py --nelem 100000 --nlev 30 --qsize 1 --ops lap --modes fast
} > $OUT/r2h_levelops.log 2>&1
