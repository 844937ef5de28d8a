#!/bin/bash
# Builds tools/probe_kstep (stand-alone microbenchmark of the chain kernels' k-step pacing; see probe_kstep.cu).  Not part of build().
set -e
cd "$(dirname "$0")/.."
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
$NVCC -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -I i2sdf_b200/csrc tools/probe_kstep.cu -o tools/probe_kstep
echo "built tools/probe_kstep"
