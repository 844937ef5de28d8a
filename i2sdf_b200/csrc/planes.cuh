// "Plane slots": how per-point activations / adjoints travel through HBM between the training kernels.
//
// A slot holds one [M][256] (or [M][48]) matrix split into bf16 HI and LO planes (value = hi + lo, ~16 mantissa bits), in
// 32-point sub tiles that ARE the SMEM image the weight-gradient tensor-core kernel consumes (canonical no-swizzle UMMA
// core matrices: 8 points x 8 columns, 16 B per point):
//     tile (128 points) = 4 sub tiles ; sub tile = [hi plane | lo plane] ; plane = [chunk k/8][32 rows][16 B]
//     element (row r, col k), plane p -> (r/32) * 2*C*512 + p * C*512 + (k/8) * 512 + (r%32) * 16 + (k%8) * 2      (C chunks)
// Two kinds of reader / writer:
//   * chain kernels (mlp_tc3.cu / mlp_tc_bwd.cu): epilogue warps (one warp = the 32 rows of one sub tile) store and load
//     16 B segments directly, 512 B contiguous per warp access;
//   * weight-gradient kernel (wgrad_planes.cu): one cp.async.bulk per sub tile, MN-major operands (MN = features, K =
//     points: SBO = 512 between 8-column groups, LBO = 128 between 8-point groups);
// so nothing is ever transposed or re-split between the forward, the backward chain and the weight gradients.
// Rows >= M of the last tile are written as zeros by whoever produces an adjoint slot (P, Q, FB, PC): the weight
// gradients sum over whole tiles.
#pragma once
#include <stdint.h>

namespace i2sdf {
namespace planes {

constexpr int TM = 128;
constexpr int BIG_CHUNKS = 32;                              // 256 columns
constexpr int SMALL_CHUNKS = 6;                             // 48 columns (positional encodings)
constexpr int SUB_ROWS = 32;
constexpr size_t SUB_CHUNK = (size_t)SUB_ROWS * 16;         // 512 B: one chunk of one sub tile
constexpr size_t BIG_PLANE = BIG_CHUNKS * SUB_CHUNK;        // 16384: one plane of one sub tile
constexpr size_t SMALL_PLANE = SMALL_CHUNKS * SUB_CHUNK;    // 3072
constexpr size_t BIG_SUB = 2 * BIG_PLANE;                   // hi + lo of one sub tile
constexpr size_t SMALL_SUB = 2 * SMALL_PLANE;
constexpr size_t BIG_TILE = 4 * BIG_SUB;                    // 131072
constexpr size_t SMALL_TILE = 4 * SMALL_SUB;                // 24576

// byte offset (from the slot base) of the HI 16-byte segment of (point m, chunk kc); the LO segment is chunks * 512 further
__host__ __device__ inline size_t seg(long long m, int kc, int chunks) {
    return (size_t)(m >> 5) * (size_t)(2 * chunks) * SUB_CHUNK + (size_t)kc * SUB_CHUNK + (size_t)(m & 31) * 16;
}

__host__ __device__ inline long long ntiles(long long M) { return (M + TM - 1) / TM; }
__host__ __device__ inline size_t big_slot_bytes(long long M) { return (size_t)ntiles(M) * BIG_TILE; }
__host__ __device__ inline size_t small_slot_bytes(long long M) { return (size_t)ntiles(M) * SMALL_TILE; }

// ---- slot directory of one saved forward state (byte offsets from the base of the buffer) ----------------------
// forward (written by tc3::tc_mlp_kernel<true> in save mode):
//   H[l]  l = 0..NL-1 : input of SDF layer l+1  = h~_l  (skip concat and 1/sqrt2 applied)
//   Q[l]  l = 0..NL-1 : adjoint of a_l in the reverse sweep = softplus'(a_l) * (W_{l+1}^T q_{l+1})       (rows >= M zero)
//   E                 : PE(x), 48 columns (input of SDF layer 0)
//   CF, C[0..Lc-2]    : radiance stack: features (input of layer 0), post-ReLU hidden activations
//   DV                : PE(view dir), 48 columns
// backward workspace (written by tcb::tc_bwd_kernel):
//   HD[l] : tangent h~dot_l ; P[l] : adjoint p_l ; ED : tangent of PE(x) ; FB : adjoint of the features ;
//   PC[l] : adjoint of the radiance pre-activations
struct SavedDir {
    size_t H[12], Q[12], E, CF, C[12], DV, total;
};
struct BwdDir {
    size_t HD[12], P[12], ED, FB, PC[12], total;
};
inline SavedDir saved_dir(long long M, int NL, int Lc, bool with_color) {
    SavedDir d{};
    size_t off = 0;
    const size_t big = big_slot_bytes(M), small = small_slot_bytes(M);
    for (int l = 0; l < NL; ++l) { d.H[l] = off; off += big; }
    for (int l = 0; l < NL; ++l) { d.Q[l] = off; off += big; }
    d.E = off; off += small;
    if (with_color) {
        d.CF = off; off += big;
        for (int l = 0; l < Lc - 1; ++l) { d.C[l] = off; off += big; }
        d.DV = off; off += small;
    }
    d.total = off;
    return d;
}
inline BwdDir bwd_dir(long long M, int NL, int Lc, bool with_color) {
    BwdDir d{};
    size_t off = 0;
    const size_t big = big_slot_bytes(M), small = small_slot_bytes(M);
    for (int l = 0; l < NL; ++l) { d.HD[l] = off; off += big; }
    for (int l = 0; l < NL; ++l) { d.P[l] = off; off += big; }
    d.ED = off; off += small;
    d.FB = off; off += big;
    if (with_color) for (int l = 0; l < Lc - 1; ++l) { d.PC[l] = off; off += big; }
    d.total = off;
    return d;
}

}  // namespace planes
}  // namespace i2sdf
