#!/bin/bash
# usage: build_variant.sh <name> [-Dflags...]: libg753 built with extra flags into ginger-lib_b200/variants/ (A/B runs)
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
out=$root/ginger-lib_b200/variants; obj=$out/obj_$name
mkdir -p $obj
cd $root/ginger-lib_b200/csrc
for f in capi msm_g0 msm_g1 msm_g2 msm_g3; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC "$@" -c -o $obj/$f.o $f.cu &
done
wait
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o $out/libg753_$name.so $obj/*.o
rm -rf $obj
ls -la $out/libg753_$name.so
