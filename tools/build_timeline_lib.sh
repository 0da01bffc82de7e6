#!/bin/bash
# Instrumented build (-DSVIT_TIMELINE: clock64 stamps of CTA 0) of the same sources into svit_b200/libsvit_sm100_tl.so.
# Use: SVIT_LIB=svit_b200/libsvit_sm100_tl.so python tools/pool_timeline.py
set -e
cd "$(dirname "$0")/.."
mkdir -p /tmp/svit_tl_obj
for f in svit_b200/csrc/*.cu; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -DSVIT_TIMELINE -c $f -o /tmp/svit_tl_obj/$(basename $f .cu).o &
done
wait
/usr/local/cuda/bin/nvcc -shared -o svit_b200/libsvit_sm100_tl.so /tmp/svit_tl_obj/*.o -gencode arch=compute_100a,code=sm_100a -lcudart
echo built svit_b200/libsvit_sm100_tl.so
