#!/bin/bash
# tools/link_variant.sh <name> [nvcc -D flags...]: compile csrc/seg_kernels.cu with the flags and link it with the other
# objects of the in-tree build into indirect_learning_pose-shape_b200/ab/lib_<name>.so (A/B builds for tools/seg_ab.py).
set -e
P=indirect_learning_pose-shape_b200
name=$1; shift
mkdir -p $P/ab
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xptxas -v "$@" -c $P/csrc/seg_kernels.cu -o $P/ab/seg_$name.o 2> $P/ab/seg_$name.log
objs=$(ls $P/build/*.o | grep -v seg_kernels.o)
nvcc -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC -o $P/ab/lib_$name.so $objs $P/ab/seg_$name.o
echo built $P/ab/lib_$name.so
