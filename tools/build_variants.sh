#!/bin/bash
# Kernel-variant experiments: builds libf3d with different -D settings into build/variants/ (git-ignored, shipped by gpurun).
# usage: tools/build_variants.sh name1:"-DFUSE_NB8=8 -DFUSE_MINB8=3" name2:"..."
set -e
cd "$(dirname "$0")/.."
PKG=3d-point-cloud-segmentation-using-2d-img-segmentation_b200
mkdir -p build/variants
for spec in "$@"; do
  name="${spec%%:*}"; defs="${spec#*:}"
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-O2,-fno-fast-math,-ffp-contract=off \
    -Xptxas -v -shared -cudart static $defs -o build/variants/libf3d_$name.so $PKG/csrc/c_api.cu $PKG/csrc/frame_setup.cu \
    $PKG/csrc/fuse_project_vote.cu $PKG/csrc/vote_resolve.cu $PKG/csrc/vote_exchange.cu $PKG/csrc/box_merge.cu > build/variants/$name.log 2>&1 &
done
wait
for spec in "$@"; do name="${spec%%:*}"; echo "== $name"; grep -A2 "fuse_kernelILi0ELi0ELi1E" build/variants/$name.log | grep -E "spill|Used"; done
