#!/bin/bash
# Builds libneuma_b200 with NEUMA_GATHER_VARIANT = 0 / 1 / 2 (csrc/mpm_common.cuh) into build_variants/
# and times tools/fuse_ab.py with each (run the timing part on the GPU box):
#   tools/gather_variants.sh build     (here: nvcc cross-compiles)
#   tools/gather_variants.sh run       (GPU box)
set -e
cd "$(dirname "$0")/.."
mkdir -p build_variants
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr"
if [ "$1" = build ]; then
  for v in 0 1 2; do
    objs=""
    for f in neuma_b200/csrc/*.cu; do
      o=build_variants/$(basename ${f%.cu})_v$v.o
      extra=""; [ "$(basename $f)" = raster_pre.cu ] && extra="-fmad=false"
      nvcc $FLAGS $extra -DNEUMA_GATHER_VARIANT=$v -c -o $o $f &
      objs="$objs $o"
    done
    wait
    nvcc -shared -gencode arch=compute_100a,code=sm_100a -o build_variants/libneuma_b200_v$v.so $objs
    rm -f $objs
  done
else
  for v in 0 1 2; do
    echo "variant $v"
    NEUMA_B200_LIB=$PWD/build_variants/libneuma_b200_v$v.so python tools/fuse_ab.py | tail -1
  done
fi
