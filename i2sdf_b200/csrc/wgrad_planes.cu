// Weight gradients of the training backward on tcgen05, straight from plane slots (planes.cuh), all layers in ONE launch.
//
//   dW_job[out][in] += sum over points m of  P[m][out] * X[m][in]        (one or two (P, X) terms per job)
//
// The reduction dimension is the POINT index, so both operands are consumed MN-major: a 32-point sub tile of a slot
// ([hi|lo][chunk][32 rows][16 B], 32 KB contiguous) is ONE cp.async.bulk into a 3-stage SMEM ring and is described to the
// tensor core with a_major = b_major = MN (SBO = chunk stride, LBO = 8-row group stride); nothing is transposed or
// converted on the way.
// Per 16-point k step and per 128-row half of `out`: 3 MMAs (P_hi X_hi + P_lo X_hi + P_hi X_lo), fp32 accumulators
// fill all 512 TMEM columns (two 128 x 256 halves).  Work = (term, tile) units split evenly over the CTAs in job order; a
// CTA flushes its accumulator (atomic adds into dW) whenever the job changes.  Idle time of the eight flush warps is
// used for the bias gradients: column sums of the P operand of flagged terms, taken from the staged SMEM tiles.
//
// Replaces (reference): the dW = grad_out^T @ input products autograd runs for every nn.Linear of ImplicitNetwork
// (mlp.py:84-105, twice: first- and second-order graph) and RenderingNetwork (mlp.py:208-229).
#include "common.cuh"
#include "planes.cuh"
#include "tc_common.cuh"
#include "wgrad_planes.cuh"

namespace i2sdf {
namespace wgp {

using namespace tc;

constexpr int NSTAGE = 3;
constexpr int SUB_ROWS = 32;
constexpr int SUB_CHUNK = SUB_ROWS * 16;          // 512 B per (chunk, 32 rows)
constexpr int PLANE_STAGE = 32 * SUB_CHUNK;       // 16 KB
constexpr int STAGE_BYTES = 4 * PLANE_STAGE;      // P_hi | P_lo | X_hi | X_lo
constexpr int N_FLUSH_WARPS = 8;
constexpr int NTHREADS = (2 + N_FLUSH_WARPS) * 32;
constexpr int XPOSE_FLOATS = 32 * 33;             // per flush warp: 32 x 32 accumulator block, padded
constexpr size_t kSmemBytes = 1024 + (size_t)NSTAGE * STAGE_BYTES + (size_t)N_FLUSH_WARPS * XPOSE_FLOATS * 4 + 256;

__device__ __forceinline__ bool flush_after(const WgArgs& A, long long u, long long u1) {
    if (u + 1 >= u1) return true;
    return A.terms[(u + 1) / A.ntiles].job != A.terms[u / A.ntiles].job;
}

__global__ void __launch_bounds__(NTHREADS, 1) wgrad_planes_kernel(const WgArgs A) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_align1024(smem_raw);
    uint8_t* ring = smem;
    float* xpose = reinterpret_cast<float*>(ring + NSTAGE * STAGE_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(xpose + N_FLUSH_WARPS * XPOSE_FLOATS);
    uint64_t* full = bars;                 // [NSTAGE] producer -> MMA (tx bytes)
    uint64_t* empty = bars + NSTAGE;       // [NSTAGE] MMA commit + flush warps (column sums) -> producer
    uint64_t* d_full = bars + 2 * NSTAGE;  // accumulator complete
    uint64_t* d_empty = d_full + 1;        // accumulator drained (8 arrivals)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d_empty + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long U = (long long)A.nterms * A.ntiles;
    const long long u0 = U * blockIdx.x / gridDim.x, u1 = U * (blockIdx.x + 1) / gridDim.x;

    if (tid == 0) {
        for (int i = 0; i < NSTAGE; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1 + N_FLUSH_WARPS); }
        mbar_init(d_full, 1);
        mbar_init(d_empty, N_FLUSH_WARPS);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================= producer: two bulk copies per stage (P sub tile, X sub tile) =================
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (long long u = u0; u < u1; ++u) {
                const WgTerm& t = A.terms[u / A.ntiles];
                const long long tile = u % A.ntiles;
                const int pp = t.p_planes == 1 ? 1 : 2, xp = t.x_planes == 1 ? 1 : 2;
                const uint32_t xplane = (uint32_t)t.xchunks * (uint32_t)planes::SUB_CHUNK;
                const uint32_t xsub = xplane * (uint32_t)xp;                                       // one X sub tile: hi (+ lo)
                const uint32_t psub = (uint32_t)planes::BIG_PLANE * (uint32_t)pp;                  // one P sub tile: hi (+ lo)
                const uint8_t* Pb = t.P + (size_t)tile * 4 * psub;
                const uint8_t* Xb = t.X + (size_t)tile * 4 * xsub;
                for (int sub = 0; sub < planes::TM / SUB_ROWS; ++sub) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    mbar_arrive_expect_tx(&full[stage], psub + xsub);
                    uint8_t* dst = ring + stage * STAGE_BYTES;
                    bulk_g2s(dst, Pb + (size_t)sub * psub, psub, &full[stage]);                                      // P_hi (| P_lo)
                    bulk_g2s(dst + 2 * PLANE_STAGE, Xb + (size_t)sub * xsub, xplane, &full[stage]);                  // X_hi
                    if (xp == 2) bulk_g2s(dst + 3 * PLANE_STAGE, Xb + (size_t)sub * xsub + xplane, xplane, &full[stage]);   // X_lo
                    if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            const uint32_t ring_s = smem_u32(ring);
            // MN-major operands: between 8-feature groups = chunk stride (SBO), between 8-point groups = 128 B (LBO)
            const uint32_t lbo = A.variant ? SUB_CHUNK : 128u, sbo = A.variant ? 128u : SUB_CHUNK;
            uint32_t stage = 0, phase = 0, ephase = 0;
            bool fresh = true;
            for (long long u = u0; u < u1; ++u) {
                const WgTerm& t = A.terms[u / A.ntiles];
                const uint32_t idesc = instr_desc_bf16(128, t.xchunks * 8) | (1u << 15) | (1u << 16);
                const bool p2 = t.p_planes != 1, x2 = t.x_planes != 1;       // an operand with its HI plane only: its lo product is skipped
                for (int sub = 0; sub < planes::TM / SUB_ROWS; ++sub) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint32_t base = ring_s + stage * STAGE_BYTES;
#pragma unroll
                    for (int ks = 0; ks < SUB_ROWS / 16; ++ks) {
                        const uint32_t koff = (uint32_t)ks * 256u;                 // 16 points further along K
                        const uint64_t b_hi = smem_desc(base + 2 * PLANE_STAGE + koff, lbo, sbo);
                        const uint64_t b_lo = smem_desc(base + 3 * PLANE_STAGE + koff, lbo, sbo);
#pragma unroll
                        for (int hf = 0; hf < 2; ++hf) {
                            const uint32_t aoff = (uint32_t)hf * 16u * SUB_CHUNK + koff;   // out features 128 hf ..
                            const uint64_t a_hi = smem_desc(base + aoff, lbo, sbo);
                            const uint64_t a_lo = smem_desc(base + PLANE_STAGE + aoff, lbo, sbo);
                            const uint32_t d = tmem_base + (uint32_t)hf * 256u;
                            mma_bf16_ss(d, a_hi, b_hi, idesc, fresh ? 0u : 1u);
                            if (p2) mma_bf16_ss(d, a_lo, b_hi, idesc, 1u);
                            if (x2) mma_bf16_ss(d, a_hi, b_lo, idesc, 1u);
                        }
                        fresh = false;
                    }
                    mma_commit(&empty[stage]);
                    if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
                }
                if (flush_after(A, u, u1)) {
                    mma_commit(d_full);
                    mbar_wait(d_empty, ephase);
                    ephase ^= 1;
                    tc_fence_after();
                    fresh = true;
                }
            }
        }
    } else {
        // ================= flush warps: bias-gradient column sums while the tensor core works, then drain TMEM ==========
        const int q = warp & 3, hf = (warp - 2) >> 2, fw = warp - 2;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        uint32_t stage = 0, phase = 0, dphase = 0;
        float cs[4][8];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int e = 0; e < 8; ++e) cs[a][e] = 0.f;
        for (long long u = u0; u < u1; ++u) {
            const int term = (int)(u / A.ntiles);
            const WgTerm& t = A.terms[term];
            for (int sub = 0; sub < planes::TM / SUB_ROWS; ++sub) {
                mbar_wait(&full[stage], phase);      // always: keeps the stage's arrival count in step with the ring
                if (t.colsum) {
                    // warp fw owns chunks fw, fw+8, fw+16, fw+24 ; lane = row of the 32-point sub tile
                    const uint8_t* sp = ring + stage * STAGE_BYTES + lane * 16;
#pragma unroll
                    for (int a = 0; a < 4; ++a) {
                        const uint4 hi = *reinterpret_cast<const uint4*>(sp + (fw + 8 * a) * SUB_CHUNK);
                        const uint4 lo = (t.p_planes != 1) ? *reinterpret_cast<const uint4*>(sp + PLANE_STAGE + (fw + 8 * a) * SUB_CHUNK) : make_uint4(0, 0, 0, 0);
                        const uint32_t hw[4] = {hi.x, hi.y, hi.z, hi.w}, lw[4] = {lo.x, lo.y, lo.z, lo.w};
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            cs[a][2 * i] += __uint_as_float(hw[i] << 16) + __uint_as_float(lw[i] << 16);
                            cs[a][2 * i + 1] += __uint_as_float(hw[i] & 0xffff0000u) + __uint_as_float(lw[i] & 0xffff0000u);
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[stage]);
                if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
            }
            const bool term_ends = (u + 1 >= u1) || ((u + 1) / A.ntiles != term);
            if (t.colsum && term_ends) {
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        float v = cs[a][e];
                        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                        const int col = (fw + 8 * a) * 8 + e;
                        if (lane == 0 && col < t.colsum_n) atomicAdd(t.colsum + col, v);
                        cs[a][e] = 0.f;
                    }
            }
            if (flush_after(A, u, u1)) {
                const WgJob& J = A.jobs[t.job];
                mbar_wait(d_full, dphase);
                dphase ^= 1;
                tc_fence_after();
                // 32 x 32 blocks: TMEM (lane = dW row) -> SMEM transpose -> one 128-byte coalesced RED per dW row
                float* xp = xpose + fw * XPOSE_FLOATS;
                const int o0 = hf * 128 + q * 32;
                for (int c0 = 0; c0 < J.cols; c0 += 32) {
                    uint32_t v[32];
                    tmem_ld32(tmem_base + lane_base + (uint32_t)hf * 256u + (uint32_t)c0, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) xp[lane * 33 + j] = __uint_as_float(v[j]);
                    __syncwarp();
                    if (c0 + lane < J.cols) {
                        float* dst = J.dW + (size_t)o0 * J.ld + c0 + lane;
                        const int nr = min(32, J.rows - o0);
                        for (int rr = 0; rr < nr; ++rr) atomicAdd(dst + (size_t)rr * J.ld, xp[rr * 33 + lane]);
                    }
                    __syncwarp();
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(d_empty);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// ---- fp32 [M][ld] <-> slot (tests, diagnostics, and the few inputs that arrive as plain arrays) -------------------
__global__ void to_planes_kernel(const float* __restrict__ X, int ld, int width, long long M, uint8_t* __restrict__ slot, int chunks) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // (tile, chunk, row)
    const long long total = planes::ntiles(M) * chunks * planes::TM;
    if (i >= total) return;
    const int r = (int)(i % planes::TM), kc = (int)((i / planes::TM) % chunks);
    const long long tile = i / ((long long)planes::TM * chunks), m = tile * planes::TM + r;
    uint32_t h[4] = {0, 0, 0, 0}, l[4] = {0, 0, 0, 0};
    if (m < M) {
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) { const int c = kc * 8 + e; v[e] = c < width ? X[(size_t)m * ld + c] : 0.f; }
#pragma unroll
        for (int e = 0; e < 4; ++e) split_bf16x2(v[2 * e], v[2 * e + 1], h[e], l[e]);
    }
    const size_t plane = (size_t)chunks * planes::SUB_CHUNK;
    uint8_t* dst = slot + planes::seg(tile * planes::TM + r, kc, chunks);
    *reinterpret_cast<uint4*>(dst) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(dst + plane) = make_uint4(l[0], l[1], l[2], l[3]);
}

__global__ void from_planes_kernel(const uint8_t* __restrict__ slot, int chunks, long long M, float* __restrict__ X, int ld, int width) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = planes::ntiles(M) * chunks * planes::TM;
    if (i >= total) return;
    const int r = (int)(i % planes::TM), kc = (int)((i / planes::TM) % chunks);
    const long long tile = i / ((long long)planes::TM * chunks), m = tile * planes::TM + r;
    if (m >= M) return;
    const size_t plane = (size_t)chunks * planes::SUB_CHUNK;
    const uint8_t* src = slot + planes::seg(m, kc, chunks);
    const uint4 hi = *reinterpret_cast<const uint4*>(src), lo = *reinterpret_cast<const uint4*>(src + plane);
    const uint32_t hw[4] = {hi.x, hi.y, hi.z, hi.w}, lw[4] = {lo.x, lo.y, lo.z, lo.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int c = kc * 8 + 2 * e;
        if (c < width) X[(size_t)m * ld + c] = __uint_as_float(hw[e] << 16) + __uint_as_float(lw[e] << 16);
        if (c + 1 < width) X[(size_t)m * ld + c + 1] = __uint_as_float(hw[e] & 0xffff0000u) + __uint_as_float(lw[e] & 0xffff0000u);
    }
}

// out[k][j] += sum_m w_k[m] * X[m][j]   for a 256-column slot (w null: plain column sums), up to 3 weight vectors per pass
// over the slot (the three rows of the rgb head's gradient read the last radiance activation once); grid = (tile groups, jobs)
// 16 warps x 2 chunks each (round 2: 8 warps x 4 chunks kept 96 accumulators + 32 x 16 B of loads per thread = 205 registers, one 8-warp CTA
// per SM and 2.7 TB/s; half the accumulators per thread doubles the warps in flight at the same register file use)
constexpr int kCsWarps = 16, kCsPer = 2;
__global__ void __launch_bounds__(kCsWarps * 32) planes_colsum_kernel(const CsArgs A) {
    const CsJob& J = A.jobs[blockIdx.y];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nw = J.w ? J.nw : 1;
    const long long t0 = A.ntiles * blockIdx.x / gridDim.x, t1 = A.ntiles * (blockIdx.x + 1) / gridDim.x;
    float cs[3][kCsPer][8];
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int a = 0; a < kCsPer; ++a)
#pragma unroll
            for (int e = 0; e < 8; ++e) cs[k][a][e] = 0.f;
    for (long long tile = t0; tile < t1; ++tile) {
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
            const long long m = tile * planes::TM + rr * 32 + lane;
            float w[3] = {1.f, 0.f, 0.f};
            if (J.w) {
                const long long wlim = (J.wrows > 0 && J.wrows < A.M) ? J.wrows : A.M;
#pragma unroll
                for (int k = 0; k < 3; ++k) w[k] = (k < nw && m < wlim) ? J.w[(size_t)m * J.wstride + k] : 0.f;
            }
#pragma unroll
            for (int a = 0; a < kCsPer; ++a) {
                const int np = J.nplanes == 1 ? 1 : 2;
                const uint8_t* sp = J.slot + planes::segp(m, warp + kCsWarps * a, planes::BIG_CHUNKS, np);
                const uint4 hi = *reinterpret_cast<const uint4*>(sp), lo = (np == 2) ? *reinterpret_cast<const uint4*>(sp + planes::BIG_PLANE) : make_uint4(0, 0, 0, 0);
                const uint32_t hw[4] = {hi.x, hi.y, hi.z, hi.w}, lw[4] = {lo.x, lo.y, lo.z, lo.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float x0 = __uint_as_float(hw[i] << 16) + __uint_as_float(lw[i] << 16);
                    const float x1 = __uint_as_float(hw[i] & 0xffff0000u) + __uint_as_float(lw[i] & 0xffff0000u);
                    cs[0][a][2 * i] = fmaf(w[0], x0, cs[0][a][2 * i]);
                    cs[0][a][2 * i + 1] = fmaf(w[0], x1, cs[0][a][2 * i + 1]);
                    if (nw > 1) {
                        cs[1][a][2 * i] = fmaf(w[1], x0, cs[1][a][2 * i]);
                        cs[1][a][2 * i + 1] = fmaf(w[1], x1, cs[1][a][2 * i + 1]);
                        cs[2][a][2 * i] = fmaf(w[2], x0, cs[2][a][2 * i]);
                        cs[2][a][2 * i + 1] = fmaf(w[2], x1, cs[2][a][2 * i + 1]);
                    }
                }
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        if (k >= nw) break;
#pragma unroll
        for (int a = 0; a < kCsPer; ++a)
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                float v = cs[k][a][e];
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                const int col = (warp + kCsWarps * a) * 8 + e;
                if (lane == 0 && col < J.n) atomicAdd(J.out + (size_t)k * J.ostride + col, v);
            }
    }
}

}  // namespace wgp

int wgrad_planes_launch(const i2sdf_handle* h, const WgArgs& args, cudaStream_t st) {
    using namespace wgp;
    if (args.nterms <= 0 || args.ntiles <= 0) return I2SDF_OK;
    static PerDeviceOnce once;
    if (once.need()) I2SDF_CUDA_CHECK(cudaFuncSetAttribute(wgrad_planes_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    WgArgs a = args;
    const char* v = getenv("I2SDF_WG_VARIANT");
    a.variant = (v && v[0] == '1') ? 1 : 0;
    const long long U = (long long)a.nterms * a.ntiles;
    const int grid = (int)(U < (long long)h->num_sms ? U : (long long)h->num_sms);
    wgrad_planes_kernel<<<grid, NTHREADS, kSmemBytes, st>>>(a);
    I2SDF_CUDA_CHECK(cudaGetLastError());
    return I2SDF_OK;
}

int planes_colsum_launch(const i2sdf_handle* h, const CsArgs& args, cudaStream_t st) {
    if (args.njobs <= 0 || args.ntiles <= 0) return I2SDF_OK;
    // (one 16-warp CTA per SM; a 4x larger grid was measured slower in round 1: 150 vs 111 us)
    int gx = (2 * h->num_sms + args.njobs - 1) / args.njobs;
    if ((long long)gx > args.ntiles) gx = (int)args.ntiles;
    wgp::planes_colsum_kernel<<<dim3(gx, args.njobs), wgp::kCsWarps * 32, 0, st>>>(args);
    I2SDF_CUDA_CHECK(cudaGetLastError());
    return I2SDF_OK;
}

int planes_pack_launch(const float* X, int ld, int width, long long M, uint8_t* slot, int chunks, cudaStream_t st) {
    const long long total = planes::ntiles(M) * chunks * planes::TM;
    if (total <= 0) return I2SDF_OK;
    wgp::to_planes_kernel<<<(int)((total + 255) / 256), 256, 0, st>>>(X, ld, width, M, slot, chunks);
    I2SDF_CUDA_CHECK(cudaGetLastError());
    return I2SDF_OK;
}

int planes_unpack_launch(const uint8_t* slot, int chunks, long long M, float* X, int ld, int width, cudaStream_t st) {
    const long long total = planes::ntiles(M) * chunks * planes::TM;
    if (total <= 0) return I2SDF_OK;
    wgp::from_planes_kernel<<<(int)((total + 255) / 256), 256, 0, st>>>(slot, chunks, M, X, ld, width);
    I2SDF_CUDA_CHECK(cudaGetLastError());
    return I2SDF_OK;
}

}  // namespace i2sdf
