#!/bin/bash
# Build a kernel-variant library for A/B experiments:  scripts/build_variant.sh NAME -DFOO=1 ...
# -> frustum_convnet_b200/variants/libfrustum_b200_NAME.so   (select it with FCN_LIB_PATH=...)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
out=frustum_convnet_b200/variants; obj=$out/obj_$name
mkdir -p $obj
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --extended-lambda -Xcompiler -fPIC -Xcompiler -fvisibility=hidden"
pids=()
for f in frustum_convnet_b200/csrc/*.cu; do
  nvcc $FLAGS "$@" -c $f -o $obj/$(basename ${f%.cu}).o & pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o $out/libfrustum_b200_$name.so $obj/*.o -Xcompiler -fPIC
rm -rf $obj
echo $out/libfrustum_b200_$name.so
