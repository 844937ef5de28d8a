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

// the same for a slot with `nplanes` planes per sub tile (1: HI plane only)
__host__ __device__ inline size_t segp(long long m, int kc, int chunks, int nplanes) {
    return (size_t)(m >> 5) * (size_t)(nplanes * chunks) * SUB_CHUNK + (size_t)kc * SUB_CHUNK + (size_t)(m & 31) * 16;
}
// Planes per slot kind (1 = HI plane only).  Adjoints (P, PC, FB) are written by the backward chain and read ONLY by the weight-gradient
// kernel, so their LO plane can be dropped for half the bytes of those streams.  Measured in round 2 (profiles/r02_summary.txt): chain
// 1.54 -> 1.46 ms, weight gradients 0.68 -> 0.59 ms, training step -3 % - but the parameter-gradient error against the reference grew from
// 2.0e-4 to 6.7e-4 (L2, 32-ray fixture; 3.6e-5 -> 3.5e-4 on the light config): a 2^-9 rounding per adjoint averages out only over many
// points.  Rejected: both planes stay.  The plane count is a per-slot-kind constant so that the trade can be re-made in one place.
constexpr int kPlanesAdj = 2;
constexpr int kPlanesHD = 2;

__host__ __device__ inline long long ntiles(long long M) { return (M + TM - 1) / TM; }
__host__ __device__ inline size_t big_slot_bytes(long long M) { return (size_t)ntiles(M) * BIG_TILE; }
__host__ __device__ inline size_t small_slot_bytes(long long M) { return (size_t)ntiles(M) * SMALL_TILE; }

// ---- slot directory (byte offsets) ---------------------------------------------------------------------------------
// saved forward state (written by tc3::tc_mlp_kernel<true> in save mode), NL = SDF hidden layers, Lc = radiance layers:
//   H(l)  l = 0..NL-1 : input of SDF layer l+1  = h~_l  (skip concat and 1/sqrt2 applied)
//   Q(l)  l = 0..NL-1 : adjoint of a_l in the reverse sweep = softplus'(a_l) * (W_{l+1}^T q_{l+1})       (rows >= M zero)
//   E                 : PE(x), 48 columns (input of SDF layer 0)
//   CF, C(0..Lc-2)    : radiance stack: features (input of layer 0), post-ReLU hidden activations
//   DV                : PE(view dir), 32 columns
// backward workspace (written by tcb::tc_bwd_kernel):
//   HD(l) : tangent h~dot_l ; P(l) : adjoint p_l ; ED : tangent of PE(x) ; FB : adjoint of the features ;
//   PC(l) : adjoint of the radiance pre-activations
constexpr int DV_CHUNKS = 4;
struct Layout {
    uint8_t* base;      // saved forward state (null: nothing is saved)
    uint8_t* wbase;     // backward workspace
    size_t big, small, dvb;
    size_t adj, hd;     // bytes of one adjoint slot (P, PC, FB) / one tangent slot (HD): big * planes / 2
    int NL, Lc, color;
    __host__ __device__ size_t H(int l) const { return (size_t)l * big; }
    __host__ __device__ size_t Q(int l) const { return (size_t)(NL + l) * big; }
    __host__ __device__ size_t E() const { return (size_t)2 * NL * big; }
    __host__ __device__ size_t CF() const { return E() + small; }
    __host__ __device__ size_t C(int l) const { return CF() + (size_t)(1 + l) * big; }
    __host__ __device__ size_t DV() const { return CF() + (size_t)Lc * big; }
    __host__ __device__ size_t saved_total() const { return color ? DV() + dvb : E() + small; }
    __host__ __device__ size_t HD(int l) const { return (size_t)l * hd; }
    __host__ __device__ size_t P(int l) const { return (size_t)NL * hd + (size_t)l * adj; }
    __host__ __device__ size_t ED() const { return (size_t)NL * hd + (size_t)NL * adj; }
    __host__ __device__ size_t FB() const { return ED() + small; }
    __host__ __device__ size_t PC(int l) const { return FB() + (size_t)(1 + l) * adj; }
    __host__ __device__ size_t bwd_total() const { return FB() + (size_t)(color ? Lc : 1) * adj; }
};
inline Layout make_layout(long long M, int NL, int Lc, bool color, void* base, void* wbase) {
    Layout L{};
    L.base = (uint8_t*)base; L.wbase = (uint8_t*)wbase;
    L.big = big_slot_bytes(M); L.small = small_slot_bytes(M); L.dvb = (size_t)ntiles(M) * 4 * 2 * DV_CHUNKS * SUB_CHUNK;
    L.adj = L.big * kPlanesAdj / 2; L.hd = L.big * kPlanesHD / 2;
    L.NL = NL; L.Lc = Lc; L.color = color ? 1 : 0;
    return L;
}

}  // namespace planes
}  // namespace i2sdf
