#!/bin/bash
# Builds the native measurement / bring-up binaries into tools/bin/ (git-ignored, NOT gpurun-ignored: they travel to the GPU box).
# Run in the build container (nvcc cross-compiles for sm_100a without a GPU) after `python -c "import __graft_entry__ as g; g.build()"`.
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/bin
ARCH="-gencode arch=compute_100a,code=sm_100a"
nvcc $ARCH -O2 -std=c++17 -o tools/bin/pair_probe tools/pair_probe.cu
nvcc $ARCH -O2 -std=c++17 -o tools/bin/l2_stream_probe tools/l2_stream_probe.cu
nvcc -O2 -std=c++17 -Wno-deprecated-gpu-targets -o tools/bin/tc_native_check tools/tc_native_check.cu -ldl
ls -la tools/bin
