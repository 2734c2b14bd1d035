#!/bin/bash
# tools/link_variant.sh <name> <source.cu> [nvcc -D flags...]: compile csrc/<source.cu> with the flags and link it with the
# other objects of the in-tree build into indirect_learning_pose-shape_b200/ab/lib_<name>.so -- A/B builds for
# tools/seg_ab.py / bench_configs (binding.LIB_PATH), never loaded by the package itself.
set -e
P=indirect_learning_pose-shape_b200
name=$1; src=$2; shift; shift
base=${src%.cu}
mkdir -p $P/ab
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xptxas -v "$@" -c $P/csrc/$src -o $P/ab/${base}_$name.o 2> $P/ab/${base}_$name.log
objs=$(ls $P/build/*.o | grep -v "/$base.o")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC -o $P/ab/lib_$name.so $objs $P/ab/${base}_$name.o
echo built $P/ab/lib_$name.so
