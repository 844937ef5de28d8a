#!/bin/bash
# Build libi2sdf_b200.so in-tree for sm_100a.  Usage: build.sh [outdir]
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="${1:-$HERE/..}"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
ARCH="-gencode arch=compute_100a,code=sm_100a"
COMMON="-O3 -std=c++17 -lineinfo -Xcompiler -fPIC $ARCH"
mkdir -p "$HERE/build"
rm -f "$HERE"/build/*.o
# the translation units compile in parallel
$NVCC $COMMON -Xptxas -v -c "$HERE/mlp_simt.cu" -o "$HERE/build/mlp_simt.o" 2> "$HERE/build/mlp_simt.ptxas.txt" &
$NVCC $COMMON -Xptxas -v -c "$HERE/mlp_tc3.cu" -o "$HERE/build/mlp_tc3.o" 2> "$HERE/build/mlp_tc3.ptxas.txt" &
$NVCC $COMMON -Xptxas -v -c "$HERE/mlp_tc_bwd.cu" -o "$HERE/build/mlp_tc_bwd.o" 2> "$HERE/build/mlp_tc_bwd.ptxas.txt" &
$NVCC $COMMON -Xptxas -v -c "$HERE/tc_gemm.cu" -o "$HERE/build/tc_gemm.o" 2> "$HERE/build/tc_gemm.ptxas.txt" &
$NVCC $COMMON -Xptxas -v -c "$HERE/wgrad_planes.cu" -o "$HERE/build/wgrad_planes.o" 2> "$HERE/build/wgrad_planes.ptxas.txt" &
$NVCC $COMMON -fmad=false -Xptxas -v -c "$HERE/sampler.cu" -o "$HERE/build/sampler.o" 2> "$HERE/build/sampler.ptxas.txt" &
$NVCC $COMMON -Xptxas -v -c "$HERE/backward.cu" -o "$HERE/build/backward.o" 2> "$HERE/build/backward.ptxas.txt" &
$NVCC $COMMON -Xptxas -v -c "$HERE/loss.cu" -o "$HERE/build/loss.o" 2> "$HERE/build/loss.ptxas.txt" &
$NVCC $COMMON -Xptxas -v -c "$HERE/wnorm.cu" -o "$HERE/build/wnorm.o" 2> "$HERE/build/wnorm.ptxas.txt" &
$NVCC $COMMON -Xptxas -v -c "$HERE/adam.cu" -o "$HERE/build/adam.o" 2> "$HERE/build/adam.ptxas.txt" &
$NVCC $COMMON -c "$HERE/c_abi.cu" -o "$HERE/build/c_abi.o" &
# check build (test infrastructure): the same ABI + the fp32 SIMT kernel / layer-wise backward as a selectable cross-check backend
$NVCC $COMMON -DI2SDF_CHECK_BUILD -c "$HERE/c_abi.cu" -o "$HERE/build/c_abi_check.o" &
$NVCC $COMMON -DI2SDF_CHECK_BUILD -c "$HERE/mlp_tc3.cu" -o "$HERE/build/mlp_tc3_check.o" &
wait
for f in mlp_simt mlp_tc3 mlp_tc_bwd tc_gemm wgrad_planes sampler backward loss wnorm adam c_abi c_abi_check mlp_tc3_check; do [ -s "$HERE/build/$f.o" ] || { echo "build failed: $f"; cat "$HERE/build/$f.ptxas.txt" 2>/dev/null | grep -i error; exit 1; }; done
# product: no mlp_simt.o
$NVCC -shared $ARCH -o "$OUT/libi2sdf_b200.so" "$HERE/build/mlp_tc3.o" "$HERE/build/mlp_tc_bwd.o" "$HERE/build/tc_gemm.o" "$HERE/build/wgrad_planes.o" "$HERE/build/sampler.o" "$HERE/build/backward.o" "$HERE/build/loss.o" "$HERE/build/wnorm.o" "$HERE/build/adam.o" "$HERE/build/c_abi.o" -lcudart
$NVCC -shared $ARCH -o "$OUT/libi2sdf_b200_check.so" "$HERE/build/mlp_simt.o" "$HERE/build/mlp_tc3_check.o" "$HERE/build/mlp_tc_bwd.o" "$HERE/build/tc_gemm.o" "$HERE/build/wgrad_planes.o" "$HERE/build/sampler.o" "$HERE/build/backward.o" "$HERE/build/loss.o" "$HERE/build/wnorm.o" "$HERE/build/adam.o" "$HERE/build/c_abi_check.o" -lcudart
echo "built $OUT/libi2sdf_b200.so (+ libi2sdf_b200_check.so)"
