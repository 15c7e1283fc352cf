#!/bin/sh
# Build the emulated render pipeline (development only): every render-path translation unit compiled
# with g++ under tools/cuda_emu/cuda_emu.h, linked into $1 (default /tmp/emu/librender_emu.so).
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
CSRC="$HERE/../../sage-3d_official_b200/csrc"
OUT=${1:-/tmp/emu/librender_emu.so}
DIR=$(dirname "$OUT")
mkdir -p "$DIR"
FLAGS="-std=c++20 -O1 -DSAGE_CUDA_EMU -DSAGE_EMU_POISON -fPIC -pthread -w -ffp-contract=off -fno-fast-math -mfma $SAGE_EMU_EXTRA"
for f in project sort bin composite frame scene_api jpeg png; do
    g++ $FLAGS -x c++ -c "$CSRC/$f.cu" -o "$DIR/emu_$f.o" &
done
g++ $FLAGS -c "$HERE/render_emu.cpp" -o "$DIR/emu_main.o" &
wait
g++ -shared -o "$OUT" "$DIR"/emu_*.o -pthread
