#!/bin/bash
# tools/build_variant.sh NAME "-DFLAG ..." : builds tools/_variants/libcaar_b200_NAME.so from csrc/ with extra
# nvcc flags (development A/B runs with tools/kernel_sweep.py --lib; *.so is git-ignored but ships to the GPU box)
set -e
name=$1; flags=$2
here=$(cd "$(dirname "$0")" && pwd)
src=$here/../tinman_sandbox_b200/csrc
out=$here/_variants
mkdir -p $out/obj_$name
for f in caar_capi caar_fused caar_fused_more caar_strict caar_aux caar_euler caar_levelops; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC \
    -I$here/../include -I$src $flags -Xptxas -v -c $src/$f.cu -o $out/obj_$name/$f.o 2> $out/obj_$name/$f.ptxas.log &
done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/libcaar_b200_$name.so $out/obj_$name/*.o -cudart static
grep -h "Used\|spill" $out/obj_$name/caar_fused.ptxas.log | sed "s/^/[$name] /"
