#!/bin/bash
# Builds variants of libb200cv.so that differ in the compile-time knobs of the tile list
# (tiles.cuh) for A/B timing on the GPU box: colvars_b200/libb200cv_<name>.so, selected with
# B200CV_LIB=...  Usage: tools/build_tile_variants.sh name "-DB200CV_TILE_UNIT=1 ..." [name flags]...
set -e
cd "$(dirname "$0")/../colvars_b200/csrc"
make -j8 > /dev/null
NVCC=/usr/local/cuda/bin/nvcc
FLAGS="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Xcompiler -fopenmp"
while [ $# -ge 2 ]; do
  name=$1; defs=$2; shift 2
  dir=/tmp/b200cv_variant_$name
  mkdir -p $dir
  for f in tiles api planner calc graph_step comm host_path measure; do
    $NVCC $FLAGS $defs -c $f.cu -o $dir/$f.o &
  done
  wait
  $NVCC -gencode arch=compute_100a,code=sm_100a -shared -o ../libb200cv_$name.so \
    $dir/*.o kernels.o sweep_inst.o fit.o cells.o -ldl -lgomp
  echo "built libb200cv_$name.so ($defs)"
done
