#!/bin/bash
# build_variant.sh NAME "-DFLAG=.. ..."  -> gpurun_out/variants/libnlos_NAME.so (kernel A/B experiments)
set -e
cd "$(dirname "$0")/../nlos_surface_optimization_b200/csrc"
OUT=../../build/variants; mkdir -p $OUT/obj_$1
for f in lbvh group_grid render_kernels render_kernels_ext mesh_kernels nlos_abi; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC -ccbin /usr/bin/g++ -Xptxas -v $2 -c $f.cu -o $OUT/obj_$1/$f.o 2> $OUT/obj_$1/$f.log &
done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -ccbin /usr/bin/g++ -o $OUT/libnlos_$1.so $OUT/obj_$1/*.o -lcudart
grep -A1 "k_forwardILb0ELb0ELb0ELb0ELb1ELi0" $OUT/obj_$1/render_kernels.log | grep -o "Used [0-9]* registers" | head -1
grep -A1 "k_forwardILb0ELb0ELb0ELb0ELb1ELi0" $OUT/obj_$1/render_kernels.log | grep -o "[0-9]* bytes spill stores" | head -1
grep -A1 "k_gradientILb0ELb0ELb0ELi0ELb1E" $OUT/obj_$1/render_kernels.log | grep -o "Used [0-9]* registers\|[0-9]* bytes spill stores" | head -2
