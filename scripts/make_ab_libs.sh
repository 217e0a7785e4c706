#!/bin/bash
# Build the two libraries of an A/B measurement here (CPU container, nvcc cross-compiles):
#   scratch_libs/libbdrt_prev.so = the working tree as it is, scratch_libs/libbdrt_new.so = working tree + PATCH.
# The working tree is left unchanged.  Then:  gpurun -- 'bash scripts/gpu_ab.sh [map|nuts|both]'
set -e
PATCH=$1
[ -f "$PATCH" ] || { echo "usage: $0 <patch file>"; exit 2; }
cd "$(dirname "$0")/.."
mkdir -p scratch_libs
bash bayes_drt_b200/csrc/build.sh > /dev/null
cp bayes_drt_b200/libbdrt.so scratch_libs/libbdrt_prev.so
git apply "$PATCH"
trap 'git apply -R "$PATCH"; bash bayes_drt_b200/csrc/build.sh > /dev/null' EXIT
bash bayes_drt_b200/csrc/build.sh | tail -1
cp bayes_drt_b200/libbdrt.so scratch_libs/libbdrt_new.so
echo "built scratch_libs/libbdrt_prev.so and scratch_libs/libbdrt_new.so"
