#!/bin/sh
# Standalone hardware self-test of the patch-extraction kernel (tools/warp_selftest.cu); the binary is git-ignored
# and travels to the GPU box with the snapshot:  gpurun -- ./tools/bin/warp_selftest
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/bin
/usr/local/cuda/bin/nvcc -O2 -std=c++17 -Wno-deprecated-gpu-targets -o tools/bin/warp_selftest tools/warp_selftest.cu \
  -Lvae_gan_mark_b200 -lvaegan_b200 -Xlinker -rpath -Xlinker '$ORIGIN/../../vae_gan_mark_b200'
