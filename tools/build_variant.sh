#!/bin/bash
# usage: tools/build_variant.sh NAME "-DMACRO1 -DMACRO2" [file.cu ...]
# Builds variants/lib_NAME.so: the listed translation units (default: knn_screen.cu) recompiled with the extra flags,
# the other objects taken from the current build.  For A/B runs on the GPU box with tools/ab_variants.sh (development aid).
set -e
name=$1; flags=$2; shift 2
files=${@:-knn_screen.cu}
root=$(cd "$(dirname "$0")/.." && pwd)
cd "$root/matternet-rs_b200/csrc"
make -s
mkdir -p "$root/variants" /tmp/variant_$name
objs=""
for f in api knn knn_exact knn_screen laplacian lambda pipeline comm bc project; do
  if [[ " $files " == *" $f.cu "* ]]; then
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --fmad=false -Xcompiler -fPIC $flags -c $f.cu -o /tmp/variant_$name/$f.o
    objs="$objs /tmp/variant_$name/$f.o"
  else
    objs="$objs $f.o"
  fi
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "$root/variants/lib_$name.so" $objs -lcudart_static -ldl -lpthread -lrt
echo "built variants/lib_$name.so"
