#!/bin/bash
# Eulerian instances: prefetch distance / cluster policy sweeps and the load-hoisting variant (tools/build_variant.sh eulhoist "-DCAAR_EUL_HOIST=1")
OUT=gpurun_out/r2p_eul_sweep.log
: > $OUT
for rep in 1 2; do
python tools/kernel_sweep.py --nelem 21600 --nlev 72 --eulerian --steps 20 --tag base72 >> $OUT 2>&1
python tools/kernel_sweep.py --nelem 21600 --nlev 72 --eulerian --steps 20 --tag hoist72 --lib tools/_variants/libcaar_b200_eulhoist.so >> $OUT 2>&1
python tools/kernel_sweep.py --nelem 12288 --nlev 128 --eulerian --steps 20 --tag base128 >> $OUT 2>&1
python tools/kernel_sweep.py --nelem 12288 --nlev 128 --eulerian --steps 20 --tag hoist128 --lib tools/_variants/libcaar_b200_eulhoist.so >> $OUT 2>&1
done
for pf in 0 18 37 111 148; do CAAR_PF_DIST=$pf python tools/kernel_sweep.py --nelem 21600 --nlev 72 --eulerian --steps 20 --tag pf$pf >> $OUT 2>&1; done
for pol in 1 2; do CAAR_CLUSTER_POLICY=$pol python tools/kernel_sweep.py --nelem 21600 --nlev 72 --eulerian --steps 20 --tag pol$pol >> $OUT 2>&1; CAAR_CLUSTER_POLICY=$pol python tools/kernel_sweep.py --nelem 12288 --nlev 128 --eulerian --steps 20 --tag L128pol$pol >> $OUT 2>&1; done
for pf in 8 16 64; do CAAR_PF_DIST=$pf python tools/kernel_sweep.py --nelem 12288 --nlev 128 --eulerian --steps 20 --tag L128pf$pf >> $OUT 2>&1; done
