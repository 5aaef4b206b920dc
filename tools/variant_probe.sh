#!/bin/bash
# variant_probe.sh "<grid_res list>" name1 name2 ... : runs tools/grid_probe.py with the default library and with each build/variants/libnlos_<name>.so
LIB=nlos_surface_optimization_b200/libnlos_b200.so
cp $LIB /tmp/libnlos_default.so
R=$1; shift
echo "== default"; python tools/grid_probe.py 64 $R | grep -v "^bvh" 
for n in "$@"; do
  cp build/variants/libnlos_$n.so $LIB
  echo "== $n"; python tools/grid_probe.py 64 $R | grep -v "^bvh"
done
cp /tmp/libnlos_default.so $LIB
