#!/bin/bash
# A/B of the igp derivatives on the FP64 tensor core (-DCAAR_DERIV_MMA=1, tools/_variants/libcaar_b200_mma.so) against the
# committed shuffle + DFMA form, interleaved on one box; then the whole -m gpu suite with the variant library swapped in
# (the box's copy of the repository is scratch).
set -u
OUT=gpurun_out
J=$OUT/r2u_mma_ab.jsonl
: > $J
MMA=tools/_variants/libcaar_b200_mma.so
ab() {  # nelem nlev tag extra...
  timeout 200 python tools/kernel_sweep.py --nelem "$1" --nlev "$2" --steps 20 --variants distinct aliased distinct "${@:4}" --tag "$3_base" >> $J 2>> $OUT/r2u_mma.err
  timeout 200 python tools/kernel_sweep.py --lib $MMA --nelem "$1" --nlev "$2" --steps 20 --variants distinct aliased distinct "${@:4}" --tag "$3_mma" >> $J 2>> $OUT/r2u_mma.err
}
ab 86400 72 ne120
ab 86400 72 ne120_again
ab 24576 128 L128
ab 21600 72 eul72 --eulerian
ab 12288 128 eul128 --eulerian
ab 51840 30 L30
cp $MMA tinman_sandbox_b200/libcaar_b200.so
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/r2u_mma_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/r2u_mma_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/r2u_mma_smoke.log 2>&1; echo "smoke rc=$?" >> $OUT/r2u_mma_smoke.log
