#!/bin/sh
# Build a tuning variant of libsage_splat.so: tools/build_variant.sh <tag> [extra nvcc flags...]
# -> sage-3d_official_b200/libsage_splat_<tag>.so (select with SAGE_SPLAT_LIB=<path>).  Experiments only.
set -e
TAG=$1; shift
HERE=$(cd "$(dirname "$0")" && pwd)
CSRC="$HERE/../sage-3d_official_b200/csrc"
OBJ=/tmp/sage_variant_$TAG
mkdir -p "$OBJ"
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC,-O2 $*"
for f in api scene_api frame project sort bin composite collide jpeg png plan; do
    nvcc $FLAGS -c "$CSRC/$f.cu" -o "$OBJ/$f.o" &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "$HERE/../sage-3d_official_b200/libsage_splat_$TAG.so" "$OBJ"/*.o -lcudart
echo "$HERE/../sage-3d_official_b200/libsage_splat_$TAG.so"
