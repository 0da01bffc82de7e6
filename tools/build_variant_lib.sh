#!/bin/bash
# A/B builds: the same sources with extra nvcc flags into svit_b200/libsvit_sm100_<name>.so (git-ignored, travels with gpurun).
# Use: tools/build_variant_lib.sh p3 -DSVIT_ATTN_EXP_POLY_MASK=3 ; SVIT_LIB=svit_b200/libsvit_sm100_p3.so python tools/...
set -e
cd "$(dirname "$0")/.."
name=$1; shift
obj=/tmp/svit_var_obj_$name
mkdir -p $obj
for f in svit_b200/csrc/*.cu; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr "$@" -c $f -o $obj/$(basename $f .cu).o &
done
wait
/usr/local/cuda/bin/nvcc -shared -o svit_b200/libsvit_sm100_$name.so $obj/*.o -gencode arch=compute_100a,code=sm_100a -lcudart
echo built svit_b200/libsvit_sm100_$name.so
