// Tensor-core GEMMs for the training backward (tcgen05 / TMEM, sm_100a), fp32 in / fp32 out with the same bf16 hi/lo
// 3-product scheme as the forward kernels (mlp_tc3.cu).
//
//  tc_gemm_pw     C[M, n] = act(A[M, K] . B + bias)     "points x weights": A is an fp32 activation array in global
//                 memory, staged per 128-row tile into the SMEM A operand (split hi/lo on the fly); B is a PRE-PACKED weight
//                 block (forward or transposed, see tc_block()) streamed by cp.async.bulk; D in TMEM, two buffers
//                 ping-ponged across tiles so the store of tile t overlaps the MMAs of tile t+1.
//  tc_gemm_wgrad  dW[n1, n2] += P[M, n1]^T . X[M, n2]    reduction over the points: both operands are fp32 activation
//                 arrays; 16 points per k step are transposed into the K-major operand layout by the loader warps
//                 (coalesced reads along the feature axis), every CTA accumulates its slice of the points in TMEM
//                 (2 x 128 feature rows x 256 columns) and writes one partial; a small kernel reduces the partials.
// Both are HBM-bound at these shapes (an [M,256] fp32 array is read or written once per pass), which is what replaces
// the FMA-pipe SGEMM of backward.cu when the tensor-core path is enabled.
#include "common.cuh"
#include "tc_common.cuh"

namespace i2sdf {
namespace tcg {

using namespace tc;

constexpr int TM = 128;
constexpr int NSTAGE = 4;
constexpr int STAGE_MAX = 16384;
constexpr int A_CHUNKS = 36;
constexpr int A_PART_BYTES = A_CHUNKS * TM * 16;
constexpr int N_EPI_WARPS = 16;
constexpr int NTHREADS = (2 + N_EPI_WARPS) * 32;
constexpr int N_READY = 9;
constexpr uint32_t LBO_A = TM * 16, SBO = 128;
constexpr size_t kSmemPW = 1024 + 2 * (size_t)A_PART_BYTES + NSTAGE * STAGE_MAX + 256;

struct PWArgs {
    const float* A; int lda; int kvalid;       // A[m][k] valid for k < kvalid (zero beyond)
    long long M;
    const uint8_t* B; int ksteps; int n;       // packed weight block: ksteps x (n*64) bytes
    float* C; int ldc; int ncols;              // write columns < ncols
    const float* bias;                         // optional [n]
    int relu;
    // optional fused input transform (tangent forward):  A'[m][k] = softplus'(A[m][k]) * A2[m][k]
    //   (+ skip concat: k >= nsplit -> E2[m][k - nsplit]; everything / sqrt2)
    const float* A2; int lda2; int is_skip; int nsplit; const float* E2;
    int in_relu;                               // A'[m][k] = max(A[m][k], 0)  (light head: relu(features))
};

// fast softplus_100 and its derivative (same approximations as the forward tensor-core kernels)
__device__ __forceinline__ float sp_fast(float a) {
    const float e = ex2_approx(-fabsf(a) * 144.26950408889634f);
    return fmaf(lg2_approx(1.0f + e), 0.0069314718055994531f, fmaxf(a, 0.0f));
}
__device__ __forceinline__ float dsp_fast(float a) {
    const float e = ex2_approx(-fabsf(a) * 144.26950408889634f);
    const float rr = rcp_approx(1.0f + e);
    return (a >= 0.f) ? rr : e * rr;
}

__device__ __forceinline__ void store_a16(uint8_t* A_hi, uint8_t* A_lo, int row, int kc0, const float (&hv)[16]) {
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        uint32_t h[4], lo[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) split_bf16x2(hv[s * 8 + 2 * i], hv[s * 8 + 2 * i + 1], h[i], lo[i]);
        const uint32_t off = seg_off<TM>(row, kc0 + s);
        *reinterpret_cast<uint4*>(A_hi + off) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4*>(A_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
}
__device__ __forceinline__ void publish(uint64_t* bar, int lane) {
    fence_proxy_async();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar);
}

__global__ void __launch_bounds__(NTHREADS, 1) gemm_pw_kernel(const PWArgs G) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_align1024(smem_raw);
    uint8_t* A_hi = smem;
    uint8_t* A_lo = smem + A_PART_BYTES;
    uint8_t* ring = smem + 2 * A_PART_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(ring + NSTAGE * STAGE_MAX);
    uint64_t* full = bars;
    uint64_t* empty = bars + NSTAGE;
    uint64_t* a_ready = bars + 2 * NSTAGE;       // [N_READY], 4 arrivals (one loader warp per lane quarter)
    uint64_t* d_full = a_ready + N_READY;        // [2]  MMAs of a tile done (A operand free, accumulator complete)
    uint64_t* d_empty = d_full + 2;              // [2]  accumulator stored (8 storer warps)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d_empty + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long ntiles = (G.M + TM - 1) / TM;
    const int nchunks = (G.ksteps + 1) >> 1;     // 32-column A chunks the MMA waits on
    const uint32_t sb = (uint32_t)G.n * 64u;

    if (tid == 0) {
        for (int i = 0; i < NSTAGE; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < N_READY; ++i) mbar_init(&a_ready[i], 4);
        mbar_init(&d_full[0], 1);
        mbar_init(&d_full[1], 1);
        mbar_init(&d_empty[0], 8);
        mbar_init(&d_empty[1], 8);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                for (int ks = 0; ks < G.ksteps; ++ks) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    mbar_arrive_expect_tx(&full[stage], sb);
                    bulk_g2s(ring + stage * STAGE_MAX, G.B + (size_t)ks * sb, sb, &full[stage]);
                    if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // all 32 lanes in converged control flow, one elected lane issues (see chain_mma in tc_chain.cuh for why)
        {
            const uint32_t a_hi_s = smem_u32(A_hi), a_lo_s = smem_u32(A_lo), ring_s = smem_u32(ring);
            const uint32_t idesc = instr_desc_bf16(TM, G.n);
            const uint32_t lbo_b = (uint32_t)G.n * 16u, lo_off = (uint32_t)G.n * 32u;
            const uint64_t dA_hi0 = smem_desc(a_hi_s, LBO_A, SBO), dA_lo0 = smem_desc(a_lo_s, LBO_A, SBO), dB0 = smem_desc(ring_s, lbo_b, SBO);
            uint32_t stage = 0, phase = 0, aphase = 0, t = 0;
            for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++t) {
                const uint32_t d_tmem = tmem_base + (t & 1u) * 256u;
                if (t >= 2) {                                   // accumulator of tile t-2 must have been stored
                    mbar_wait(&d_empty[t & 1u], ((t >> 1) - 1) & 1u);
                    tc_fence_after();
                }
                for (int ks = 0; ks < G.ksteps; ++ks) {
                    if ((ks & 1) == 0) {
                        const int c = ks >> 1;
                        mbar_wait(&a_ready[c], (aphase >> c) & 1u);
                        aphase ^= (1u << c);
                    }
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint64_t da_hi = dA_hi0 + (uint64_t)(((uint32_t)ks * 2u * LBO_A) >> 4);
                    const uint64_t da_lo = dA_lo0 + (uint64_t)(((uint32_t)ks * 2u * LBO_A) >> 4);
                    const uint64_t db_hi = dB0 + (uint64_t)((stage * (uint32_t)STAGE_MAX) >> 4);
                    const uint64_t db_lo = db_hi + (uint64_t)(lo_off >> 4);
                    if (elect_one_sync()) {
                        mma_bf16_ss(d_tmem, da_hi, db_hi, idesc, ks > 0 ? 1u : 0u);
                        mma_bf16_ss(d_tmem, da_lo, db_hi, idesc, 1u);
                        mma_bf16_ss(d_tmem, da_hi, db_lo, idesc, 1u);
                        mma_commit(&empty[stage]);
                    }
                    __syncwarp();
                    if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
                }
                if (elect_one_sync()) mma_commit(&d_full[t & 1u]);
                __syncwarp();
            }
        }
    } else {
        const int q = warp & 3;
        const int sub = (warp - 2) >> 2;              // 0,1: storers (TMEM -> global)   2,3: loaders (global -> SMEM operand)
        const int row = q * 32 + lane;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        const int kpad = G.ksteps * 16;
        const bool vec_ok = ((G.lda & 3) == 0) && ((reinterpret_cast<uintptr_t>(G.A) & 15) == 0);
        const bool cvec_ok = ((G.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(G.C) & 15) == 0);

        if (sub >= 2) {
            // ================= loaders: stage tile t's A operand as soon as the MMAs of tile t-1 released it =================
            const int s2 = sub - 2;
            auto fetch_item = [&](const float* arow, long long m, bool rowok, int col0, float (&hv)[16]) {
                if (rowok && vec_ok && col0 + 16 <= G.kvalid) {
#pragma unroll
                    for (int j4 = 0; j4 < 4; ++j4) {
                        const float4 v = *reinterpret_cast<const float4*>(arow + col0 + j4 * 4);
                        hv[j4 * 4] = v.x; hv[j4 * 4 + 1] = v.y; hv[j4 * 4 + 2] = v.z; hv[j4 * 4 + 3] = v.w;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) hv[j] = (rowok && col0 + j < G.kvalid) ? arow[col0 + j] : 0.f;
                }
                if (G.in_relu) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) hv[j] = fmaxf(hv[j], 0.f);
                }
                if (G.A2) {          // fused: softplus'(a) * adot  (+ skip concat)
                    const float* a2 = G.A2 + (size_t)(rowok ? m : 0) * G.lda2 + col0;
#pragma unroll
                    for (int j4 = 0; j4 < 4; ++j4) {
                        const float4 t = rowok ? *reinterpret_cast<const float4*>(a2 + j4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
                        hv[j4 * 4] = dsp_fast(hv[j4 * 4]) * t.x; hv[j4 * 4 + 1] = dsp_fast(hv[j4 * 4 + 1]) * t.y;
                        hv[j4 * 4 + 2] = dsp_fast(hv[j4 * 4 + 2]) * t.z; hv[j4 * 4 + 3] = dsp_fast(hv[j4 * 4 + 3]) * t.w;
                    }
                    if (G.is_skip) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const int k = col0 + j;
                            if (k >= G.nsplit) hv[j] = rowok ? G.E2[(size_t)m * 40 + (k - G.nsplit)] : 0.f;
                            hv[j] *= 0.70710678118654752f;
                        }
                    }
                }
            };
            uint32_t t = 0, fphase = 0;
            for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++t) {
                if (t >= 1) {                                   // MMAs of the previous tile finished reading the A operand
                    const uint32_t b = (t - 1) & 1u;
                    mbar_wait(&d_full[b], (fphase >> b) & 1u);
                    fphase ^= (1u << b);
                }
                const long long m = tile * TM + row;
                const bool rowok = m < G.M;
                const float* arow = G.A + (size_t)(rowok ? m : 0) * G.lda;
                for (int c = s2; c < nchunks; c += 2) {         // this warp: every other 32-column chunk, two 16-column items each
                    float h0[16], h1[16];
                    const int col0 = c * 32;
                    const bool v0 = col0 < kpad, v1 = col0 + 16 < kpad;
                    if (v0) fetch_item(arow, m, rowok, col0, h0);
                    if (v1) fetch_item(arow, m, rowok, col0 + 16, h1);
                    if (v0) store_a16(A_hi, A_lo, row, col0 >> 3, h0);
                    if (v1) store_a16(A_hi, A_lo, row, (col0 + 16) >> 3, h1);
                    publish(&a_ready[c], lane);
                }
            }
        } else {
            // ================= storers: accumulator of tile t -> global, overlapping the staging / MMAs of tile t+1 =================
            uint32_t t = 0, dphase = 0;
            for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++t) {
                const uint32_t b = t & 1u;
                mbar_wait(&d_full[b], (dphase >> b) & 1u);
                dphase ^= (1u << b);
                tc_fence_after();
                const long long m = tile * TM + row;
                for (int it = 0; it < 8; ++it) {
                    const int col0 = it * 32 + sub * 16;
                    if (col0 >= G.n) break;
                    uint32_t v[16];
                    tmem_ld16(tmem_base + lane_base + b * 256u + (uint32_t)col0, v);
                    tmem_ld_wait();
                    if (m < G.M && col0 < G.ncols) {
                        float o[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            float x = __uint_as_float(v[j]);
                            if (G.bias) x += __ldg(G.bias + col0 + j);
                            if (G.relu) x = fmaxf(x, 0.f);
                            o[j] = x;
                        }
                        float* crow = G.C + (size_t)m * G.ldc + col0;
                        if (cvec_ok && col0 + 16 <= G.ncols) {
#pragma unroll
                            for (int j4 = 0; j4 < 4; ++j4) *reinterpret_cast<float4*>(crow + j4 * 4) = make_float4(o[j4 * 4], o[j4 * 4 + 1], o[j4 * 4 + 2], o[j4 * 4 + 3]);
                        } else {
#pragma unroll
                            for (int j = 0; j < 16; ++j) if (col0 + j < G.ncols) crow[j] = o[j];
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&d_empty[b]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// =================================================================================================
// weight gradient:  partial[cta][n1pad][256] = sum over the CTA's points of  P[m][i] * X[m][j]
// =================================================================================================
constexpr int WG_STAGES = 4;
constexpr int WG_STAGE_BYTES = 32768;     // A_hi 8K | A_lo 8K | B_hi 8K | B_lo 8K   (256 rows x 2 k-chunks x 16 B each)
constexpr size_t kSmemWG = 1024 + (size_t)WG_STAGES * WG_STAGE_BYTES + 256;

struct WGArgs {
    const float* P0; int ldp0; const float* X0; int ldx0;     // first source pair
    const float* P1; int ldp1; const float* X1; int ldx1;     // optional second pair (null = none), same shapes
    int n1, n2;                                               // valid feature counts (<= 256 each)
    long long M;
    float* partial;                                           // [grid][256][256]
    // optional fused X transform: X0 := softplus(Aprev) and X1 := softplus'(Aprev) * Adprev  (ld 256), with the skip concat
    // (columns >= nsplit come from E0 / E1 [M][40]; everything / sqrt2).  When Aprev is set X0 / X1 pointers are ignored.
    const float* Aprev; const float* Adprev; int is_skip; int nsplit; const float* E0; const float* E1;
    int x_relu;                                               // X := max(X, 0) while staging (light head: relu(features))
};

__global__ void __launch_bounds__(NTHREADS, 1) gemm_wgrad_kernel(const WGArgs G) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_align1024(smem_raw);
    uint8_t* ring = smem;
    uint64_t* bars = reinterpret_cast<uint64_t*>(ring + WG_STAGES * WG_STAGE_BYTES);
    uint64_t* full = bars;                   // [WG_STAGES] 16 arrivals (loader warps)
    uint64_t* empty = bars + WG_STAGES;      // [WG_STAGES] 1 arrival (tcgen05.commit)
    uint64_t* d_full = empty + WG_STAGES;    // [1]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d_full + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // this CTA's slice of the points, in units of 16
    const long long nk16 = (G.M + 15) / 16;
    const long long per = (nk16 + gridDim.x - 1) / gridDim.x;
    const long long k_lo = (long long)blockIdx.x * per, k_hi = (k_lo + per < nk16) ? k_lo + per : nk16;
    const int npairs = G.P1 ? 2 : 1;
    const long long nsteps = (k_hi > k_lo ? (k_hi - k_lo) : 0) * npairs;
    const int n2pad = (G.n2 + 15) & ~15;
    const int mt_count = G.n1 > 128 ? 2 : 1;

    if (tid == 0) {
        for (int i = 0; i < WG_STAGES; ++i) { mbar_init(&full[i], N_EPI_WARPS); mbar_init(&empty[i], 1); }
        mbar_init(d_full, 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 1) {
        // all 32 lanes converged, one elected lane issues (see chain_mma in tc_chain.cuh)
        {
            const uint32_t ring_s = smem_u32(ring);
            const uint32_t idesc = instr_desc_bf16(TM, n2pad);
            const uint32_t lbo = 256u * 16u;
            uint32_t stage = 0, phase = 0;
            for (long long st = 0; st < nsteps; ++st) {
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                const uint32_t base = ring_s + stage * WG_STAGE_BYTES;
                const uint64_t db_hi = smem_desc(base + 16384, lbo, SBO);
                const uint64_t db_lo = smem_desc(base + 24576, lbo, SBO);
                if (elect_one_sync()) {
                    for (int mt = 0; mt < mt_count; ++mt) {
                        const uint32_t d_tmem = tmem_base + (uint32_t)mt * 256u;
                        const uint64_t da_hi = smem_desc(base + (uint32_t)mt * 2048u, lbo, SBO);
                        const uint64_t da_lo = smem_desc(base + 8192 + (uint32_t)mt * 2048u, lbo, SBO);
                        mma_bf16_ss(d_tmem, da_hi, db_hi, idesc, st > 0 ? 1u : 0u);
                        mma_bf16_ss(d_tmem, da_lo, db_hi, idesc, 1u);
                        mma_bf16_ss(d_tmem, da_hi, db_lo, idesc, 1u);
                    }
                    mma_commit(&empty[stage]);
                }
                __syncwarp();
                if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
            }
            if (elect_one_sync()) mma_commit(d_full);
            __syncwarp();
        }
    } else if (warp >= 2) {
        // ---- loaders: thread e = (kc, r): feature row r (0..255), k chunk kc (8 points) of the A (P) and B (X) operands
        const int e = tid - 64;                  // 0..511
        const int kc = e >> 8, r = e & 255;
        uint32_t stage = 0, phase = 0;
        // fp32 operands of one k step (8 points of feature row r) -> registers
        auto fetch = [&](long long st, float (&a)[8], float (&bv)[8]) {
            const int pair = (npairs == 2) ? (int)(st & 1) : 0;
            const long long k16 = k_lo + (npairs == 2 ? (st >> 1) : st);
            const float* Pp = pair ? G.P1 : G.P0;
            const float* Xp = pair ? G.X1 : G.X0;
            const int ldp = pair ? G.ldp1 : G.ldp0, ldx = pair ? G.ldx1 : G.ldx0;
            const long long p0 = k16 * 16 + kc * 8;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const long long p = p0 + i;
                const bool ok = p < G.M;
                a[i] = (ok && r < G.n1) ? __ldg(Pp + (size_t)p * ldp + r) : 0.f;
                if (!G.Aprev) {
                    bv[i] = (ok && r < G.n2) ? __ldg(Xp + (size_t)p * ldx + r) : 0.f;
                    if (G.x_relu) bv[i] = fmaxf(bv[i], 0.f);
                } else if (!(ok && r < G.n2)) {
                    bv[i] = 0.f;
                } else if (G.is_skip && r >= G.nsplit) {
                    bv[i] = __ldg((pair ? G.E1 : G.E0) + (size_t)p * 40 + (r - G.nsplit)) * 0.70710678118654752f;
                } else {
                    const float av = __ldg(G.Aprev + (size_t)p * 256 + r);
                    float x = pair ? dsp_fast(av) * __ldg(G.Adprev + (size_t)p * 256 + r) : sp_fast(av);
                    bv[i] = G.is_skip ? x * 0.70710678118654752f : x;
                }
            }
        };
        auto commit = [&](const float (&a)[8], const float (&bv)[8]) {
            mbar_wait(&empty[stage], phase ^ 1);
            uint8_t* base = ring + stage * WG_STAGE_BYTES;
            const uint32_t off = (uint32_t)kc * 4096u + (uint32_t)(r >> 3) * 128u + (uint32_t)(r & 7) * 16u;
            uint32_t h[4], lo[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) split_bf16x2(a[2 * i], a[2 * i + 1], h[i], lo[i]);
            *reinterpret_cast<uint4*>(base + off) = make_uint4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<uint4*>(base + 8192 + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
#pragma unroll
            for (int i = 0; i < 4; ++i) split_bf16x2(bv[2 * i], bv[2 * i + 1], h[i], lo[i]);
            *reinterpret_cast<uint4*>(base + 16384 + off) = make_uint4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<uint4*>(base + 24576 + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&full[stage]);
            if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
        };
        for (long long st = 0; st < nsteps; st += 2) {       // two k steps of loads in flight per thread
            float a0[8], b0[8], a1[8], b1[8];
            fetch(st, a0, b0);
            const bool two = st + 1 < nsteps;
            if (two) fetch(st + 1, a1, b1);
            commit(a0, b0);
            if (two) commit(a1, b1);
        }
        // ---- epilogue: TMEM -> partial
        mbar_wait(d_full, 0);
        tc_fence_after();
        const int q = warp & 3, sub = (warp - 2) >> 2;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        float* part = G.partial + (size_t)blockIdx.x * 256 * 256;
        for (int mt = 0; mt < mt_count; ++mt) {
            const int i = mt * 128 + q * 32 + lane;            // output row (feature of P)
            for (int it = 0; it < 4; ++it) {
                const int col0 = (sub * 4 + it) * 16;
                if (col0 >= n2pad) break;
                uint32_t v[16];
                tmem_ld16(tmem_base + lane_base + (uint32_t)mt * 256u + (uint32_t)col0, v);
                tmem_ld_wait();
                float* dst = part + (size_t)i * 256 + col0;
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4)
                    *reinterpret_cast<float4*>(dst + j4 * 4) = make_float4(nsteps > 0 ? __uint_as_float(v[j4 * 4]) : 0.f, nsteps > 0 ? __uint_as_float(v[j4 * 4 + 1]) : 0.f,
                                                                           nsteps > 0 ? __uint_as_float(v[j4 * 4 + 2]) : 0.f, nsteps > 0 ? __uint_as_float(v[j4 * 4 + 3]) : 0.f);
            }
        }
        tc_fence_before();
    }
    if (warp == 0 || nsteps == 0) { /* idle */ }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// dW[i][j] += sum_c partial[c][i][j]
__global__ void reduce_partials_kernel(float* __restrict__ dW, int ldw, int n1, int n2, const float* __restrict__ partial, int nparts) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n1 * n2) return;
    int i = idx / n2, j = idx % n2;
    float s = 0.f;
    for (int c = 0; c < nparts; ++c) s += partial[((size_t)c * 256 + i) * 256 + j];
    dW[(size_t)i * ldw + j] += s;
}

}  // namespace tcg

size_t tc_wgrad_ws_floats(const i2sdf_handle* h) { return (size_t)h->num_sms * 256 * 256; }

int tc_gemm_pw_ex(const i2sdf_handle* h, cudaStream_t st, long long M, const float* A, int lda, int kvalid, const TcBlock& blk, float* C, int ldc,
                  int ncols, const float* bias, int relu, const float* A2, int lda2, int is_skip, int nsplit, const float* E2);
int tc_gemm_pw(const i2sdf_handle* h, cudaStream_t st, long long M, const float* A, int lda, int kvalid, const TcBlock& blk, float* C, int ldc,
               int ncols, const float* bias, int relu) {
    return tc_gemm_pw_ex(h, st, M, A, lda, kvalid, blk, C, ldc, ncols, bias, relu, nullptr, 0, 0, 0, nullptr);
}
int tc_gemm_pw_ex(const i2sdf_handle* h, cudaStream_t st, long long M, const float* A, int lda, int kvalid, const TcBlock& blk, float* C, int ldc,
                  int ncols, const float* bias, int relu, const float* A2, int lda2, int is_skip, int nsplit, const float* E2) {
    using namespace tcg;
    static PerDeviceOnce once;
    if (once.need()) {
        I2SDF_CUDA_CHECK(cudaFuncSetAttribute(gemm_pw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemPW));
        I2SDF_CUDA_CHECK(cudaFuncSetAttribute(gemm_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemWG));
    }
    if (M <= 0) return I2SDF_OK;
    if (!blk.ptr || blk.ksteps * 16 > A_CHUNKS * 8 || blk.n > 256) { set_error("tc_gemm_pw: bad weight block"); return I2SDF_E_INVALID; }
    PWArgs G;
    G.A = A; G.lda = lda; G.kvalid = kvalid; G.M = M; G.B = blk.ptr; G.ksteps = blk.ksteps; G.n = blk.n; G.C = C; G.ldc = ldc; G.ncols = ncols;
    G.bias = bias; G.relu = relu;
    G.A2 = A2; G.lda2 = lda2; G.is_skip = is_skip; G.nsplit = nsplit; G.E2 = E2;
    G.in_relu = (relu & 2) ? 1 : 0; G.relu = relu & 1;       // relu bit 0: ReLU on the output, bit 1: ReLU on the input
    long long ntiles = (M + TM - 1) / TM;
    int grid = (int)(ntiles < (long long)h->num_sms ? ntiles : (long long)h->num_sms);
    gemm_pw_kernel<<<grid, NTHREADS, kSmemPW, st>>>(G);
    I2SDF_CUDA_CHECK(cudaGetLastError());
    return I2SDF_OK;
}

// dW[n1][ldw] (+=) P0^T X0 (+ P1^T X1).  ws: tc_wgrad_ws_floats(h) floats.
int tc_gemm_wgrad_ex(const i2sdf_handle* h, cudaStream_t st, long long M, const float* P0, int ldp0, const float* X0, int ldx0, const float* P1,
                     int ldp1, const float* X1, int ldx1, int n1, int n2, float* dW, int ldw, float* ws, const float* Aprev, const float* Adprev,
                     int is_skip, int nsplit, const float* E0, const float* E1);
int tc_gemm_wgrad(const i2sdf_handle* h, cudaStream_t st, long long M, const float* P0, int ldp0, const float* X0, int ldx0, const float* P1,
                  int ldp1, const float* X1, int ldx1, int n1, int n2, float* dW, int ldw, float* ws) {
    return tc_gemm_wgrad_ex(h, st, M, P0, ldp0, X0, ldx0, P1, ldp1, X1, ldx1, n1, n2, dW, ldw, ws, nullptr, nullptr, 0, 0, nullptr, nullptr);
}
// Aprev != null: X0 / X1 are computed on the fly from the saved pre-activations / tangents of the previous layer
int tc_gemm_wgrad_ex(const i2sdf_handle* h, cudaStream_t st, long long M, const float* P0, int ldp0, const float* X0, int ldx0, const float* P1,
                     int ldp1, const float* X1, int ldx1, int n1, int n2, float* dW, int ldw, float* ws, const float* Aprev, const float* Adprev,
                     int is_skip, int nsplit, const float* E0, const float* E1) {
    using namespace tcg;
    static PerDeviceOnce once;
    if (once.need()) {
        I2SDF_CUDA_CHECK(cudaFuncSetAttribute(gemm_pw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemPW));
        I2SDF_CUDA_CHECK(cudaFuncSetAttribute(gemm_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemWG));
    }
    if (M <= 0) return I2SDF_OK;
    if (n1 > 256 || n2 > 256 || n1 < 1 || n2 < 1) { set_error("tc_gemm_wgrad: n1/n2 out of range"); return I2SDF_E_INVALID; }
    WGArgs G;
    G.P0 = P0; G.ldp0 = ldp0; G.X0 = X0; G.ldx0 = ldx0; G.P1 = P1; G.ldp1 = ldp1; G.X1 = X1; G.ldx1 = ldx1; G.n1 = n1; G.n2 = n2; G.M = M; G.partial = ws;
    G.Aprev = Aprev; G.Adprev = Adprev; G.is_skip = is_skip; G.nsplit = nsplit; G.E0 = E0; G.E1 = E1;
    G.x_relu = (is_skip & 2) ? 1 : 0; G.is_skip = is_skip & 1;      // is_skip bit 1 (only without Aprev): ReLU on X
    long long nk16 = (M + 15) / 16;
    int grid = (int)(nk16 < (long long)h->num_sms ? nk16 : (long long)h->num_sms);
    gemm_wgrad_kernel<<<grid, NTHREADS, kSmemWG, st>>>(G);
    I2SDF_CUDA_CHECK(cudaGetLastError());
    reduce_partials_kernel<<<(n1 * n2 + 255) / 256, 256, 0, st>>>(dW, ldw, n1, n2, ws, grid);
    I2SDF_CUDA_CHECK(cudaGetLastError());
    return I2SDF_OK;
}

}  // namespace i2sdf
