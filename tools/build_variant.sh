#!/bin/bash
# Build build_variants/librvcb200_<tag>.so from the working tree with conv_tc.cu taken from git rev $1 (or "wt" = working tree).
set -e
rev=$1; tag=${2:-$1}
root=$(cd "$(dirname "$0")/.." && pwd)
tmp=$(mktemp -d)
mkdir -p $tmp/comfy_rvc_b200/csrc $tmp/include $root/build_variants
cp $root/comfy_rvc_b200/csrc/*.cu $root/comfy_rvc_b200/csrc/*.cuh $tmp/comfy_rvc_b200/csrc/
cp $root/include/rvcb200.h $tmp/include/
if [ "$rev" != "wt" ]; then git -C $root show $rev:comfy_rvc_b200/csrc/conv_tc.cu > $tmp/comfy_rvc_b200/csrc/conv_tc.cu; fi
objs=""
for f in $tmp/comfy_rvc_b200/csrc/*.cu; do
  o=${f%.cu}.o; objs="$objs $o"
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --fmad=true -cudart static -c $f -o $o &
done
wait
nvcc -shared -o $root/build_variants/librvcb200_$tag.so $objs -cudart static -Xcompiler -fPIC -gencode arch=compute_100a,code=sm_100a
rm -rf $tmp
echo built build_variants/librvcb200_$tag.so
