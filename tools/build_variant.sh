#!/bin/bash
# Build an experiment variant of the library next to the product one:  tools/build_variant.sh NAME -DFLAG ...
# -> valle2_b200/lib/libvalle_b200_NAME.so ; run with VALLE_B200_LIB=$PWD/valle2_b200/lib/libvalle_b200_NAME.so
set -e
name=$1; shift
d=valle2_b200/build_$name
mkdir -p $d
for f in valle2_b200/csrc/*.cu; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr "$@" -c $f -o $d/$(basename ${f%.cu}).o &
done
wait
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o valle2_b200/lib/libvalle_b200_$name.so $d/*.o -cudart static
echo built valle2_b200/lib/libvalle_b200_$name.so
