#!/bin/bash
# quick_variant.sh NAME "-DFLAG=.. ..." -> build/variants/libnlos_q_NAME.so: render_kernels.cu rebuilt with NLOS_QUICK_BUILD (headline instantiations
# only, ~15 s) + the flags, linked with the in-tree objects of the other translation units.  For kernel A/B experiments (tools/quick_probe.py).
set -e
cd "$(dirname "$0")/../nlos_surface_optimization_b200/csrc"
OUT=../../build/variants; mkdir -p $OUT
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC -ccbin /usr/bin/g++ -Xptxas -v -DNLOS_QUICK_BUILD $2 -c render_kernels.cu -o $OUT/q_$1.o 2> $OUT/q_$1.log
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -ccbin /usr/bin/g++ -o $OUT/libnlos_q_$1.so $OUT/q_$1.o lbvh.o group_grid.o render_kernels_ext.o mesh_kernels.o nlos_abi.o -lcudart
echo "$1: $(grep -A2 'k_forward_gridILb0ELb0ELb0ELb0ELb1ELi0ELb0ELi4' $OUT/q_$1.log | grep -o 'Used [0-9]* registers\|[0-9]* bytes spill stores' | tr '\n' ' ')"
