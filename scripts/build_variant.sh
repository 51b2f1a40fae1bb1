#!/bin/bash
# build a probe variant of libglia_rd.so: scripts/build_variant.sh <tag> <extra nvcc -D flags...>
# -> glia_b200/lib/libglia_rd_<tag>.so (use with GLIA_RD_LIB=...); objects under /tmp
set -e
tag=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
out=/tmp/glia_variant_$tag; mkdir -p $out
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -I$root/include $*"
for tu in c_api engine_f32 engine_f64; do
  nvcc $F -c $root/glia_b200/csrc/$tu.cu -o $out/$tu.o &
done
wait
nvcc -shared -o $root/glia_b200/lib/libglia_rd_$tag.so $out/c_api.o $out/engine_f32.o $out/engine_f64.o -gencode arch=compute_100a,code=sm_100a
echo built $root/glia_b200/lib/libglia_rd_$tag.so
