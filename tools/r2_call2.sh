#!/bin/bash
# GPU call 2: the whole -m gpu suite with the padded instances / 3-D tensor maps / new tests, then a few timings.
set -u
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -x -q --timeout=900 > $OUT/r2b_pytest.log 2>&1
echo "pytest rc=$?" >> $OUT/r2b_pytest.log
{
python tools/kernel_sweep.py --nelem 86400 --nlev 72 --steps 20 --tag ne120
python tools/kernel_sweep.py --nelem 5400 --nlev 72 --steps 100 --tag ne30
python tools/kernel_sweep.py --nelem 49152 --nlev 128 --steps 10 --tag ne256slice
python tools/kernel_sweep.py --nelem 100000 --nlev 30 --steps 20 --tag nlev30
python tools/kernel_sweep.py --nelem 100000 --nlev 26 --steps 20 --tag nlev26
python tools/kernel_sweep.py --nelem 60000 --nlev 88 --steps 20 --tag nlev88
python tools/kernel_sweep.py --nelem 50000 --nlev 104 --steps 20 --tag nlev104
python tools/kernel_sweep.py --nelem 21600 --nlev 72 --eulerian --steps 20 --tag eul72
} > $OUT/r2b_sweep.log 2>&1
