#!/bin/bash
# Variants of libb200cv.so that differ in compile-time knobs of one translation unit, for A/B
# timing on the GPU box (B200CV_LIB=... selects one):
#   tools/build_variants.sh name file.cu "-DKNOB=..." [name file.cu flags]...
set -e
cd "$(dirname "$0")/../colvars_b200/csrc"
make -j8 > /dev/null
NVCC=/usr/local/cuda/bin/nvcc
FLAGS="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Xcompiler -fopenmp"
OBJS="kernels sweep_inst api planner calc graph_step comm host_path measure fit cells cellsort keysort rows"
while [ $# -ge 3 ]; do
  name=$1; src=$2; defs=$3; shift 3
  $NVCC $FLAGS $defs -c $src -o /tmp/variant_$name.o
  objs=""
  for o in $OBJS; do
    if [ "$o.cu" == "$src" ]; then objs="$objs /tmp/variant_$name.o"; else objs="$objs $o.o"; fi
  done
  $NVCC -gencode arch=compute_100a,code=sm_100a -shared -o ../libb200cv_$name.so $objs -ldl -lgomp
  echo "built libb200cv_$name.so ($src $defs)"
done
