// SDF stack on the 5th-gen tensor cores (tcgen05 / TMEM), sm_100a.
//
// One persistent CTA per SM owns a tile of 128 ray points through all hidden layers of the SDF network:
//   * activations live in shared memory as the A operand, split into bf16 hi + bf16 lo (x ~= hi + lo, ~16 mantissa
//     bits) in the canonical K-major layout, and never leave the SM;
//   * weights (pre-split hi/lo, pre-tiled per 16-wide k step) stream from L2 through a ring of cp.async.bulk stages;
//   * each layer is 3 tcgen05.mma products per k step -- A_hi*W_hi + A_lo*W_hi + A_hi*W_lo -- accumulated in fp32 in
//     TMEM (128 lanes x 256 columns; two accumulators ping-pong across layers);
//   * sixteen epilogue warps (4 per scheduler; warp (q, sub) owns TMEM lane quarter q and 32-column chunks sub, sub+4)
//     read the accumulator with tcgen05.ld, add bias, apply softplus_100 (and the skip concat), re-split to bf16
//     hi/lo and write the next layer's A operand chunk by chunk; the MMA warp trails the epilogue by chunks, so layer
//     l+1's tensor work overlaps layer l's epilogue on a single tile;
//   * the sdf head (one 256-wide dot) is folded into the last epilogue in fp32; only sdf[M] goes back to HBM.
//
// Warp roles: warp 0 = weight producer (one lane), warp 1 = MMA issuer (one lane), warps 2..17 = epilogue.
// Replaces (reference): ImplicitNetwork.get_sdf_vals mlp.py:145-151 (-> forward :84-105, Embedder embedder.py:28-38),
// as called by the sampler at ray_sampler.py:84-89.
#include "common.cuh"
#include "tc_common.cuh"

namespace i2sdf {
namespace tcsdf {

using namespace tc;

constexpr int TM = 128;                 // points per tile = MMA M
constexpr int NCOL = 256;               // MMA N = layer width
constexpr int NSTAGE = 4;
constexpr int STAGE_BYTES = 16384;      // one k step: W_hi [2 chunks][256][8] bf16 (8 KB) + W_lo (8 KB)
constexpr int A_PART_BYTES = TM * 256 * 2;      // 64 KB per bf16 part
constexpr int N_EPI_WARPS = 16;         // 4 per TMEM lane quarter; warp (q, sub) owns 32-column chunks sub and sub+4
constexpr int NTHREADS = (2 + N_EPI_WARPS) * 32;
constexpr int K0_STEPS = 3;             // layer 0: 39 -> 48 columns
constexpr int STASH_LD = 40;            // embedding / sqrt2 per row, for the skip concat
constexpr uint32_t LBO_A = TM * 16, LBO_B = NCOL * 16, SBO = 128;
constexpr size_t kSmemBytes = 1024 + 2 * A_PART_BYTES + NSTAGE * STAGE_BYTES + TM * STASH_LD * 4 + 4 * TM * 4 + 256;

struct TcNet {
    const uint8_t* wpack;       // [layer][kstep][16 KB]
    int layer_off[kMaxLayers];  // in k-step units
    int ksteps[kMaxLayers];
    int NL;                     // MMA layers = L-1
};

__device__ __forceinline__ bool round_active(const MlpParams& P) {
    if (P.round_idx <= 0) return true;
    float b0 = fabsf(*P.beta_param) + P.beta_min;
    for (int j = 0; j < P.round_idx; ++j)
        if (!(P.beta_max[j] > b0)) return false;
    return true;
}

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(N_EPI_WARPS * 32) : "memory"); }

// write 32 consecutive columns (k chunks kc0..kc0+3) of one row of the next layer's A operand, split hi / lo
__device__ __forceinline__ void store_a_chunk(uint8_t* A_hi, uint8_t* A_lo, int row, int kc0, const float (&hv)[32]) {
#pragma unroll
    for (int s4 = 0; s4 < 4; ++s4) {
        uint32_t h[4], lo[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) split_bf16x2(hv[s4 * 8 + 2 * i], hv[s4 * 8 + 2 * i + 1], h[i], lo[i]);
        const uint32_t off = seg_off<TM>(row, kc0 + s4);
        *reinterpret_cast<uint4*>(A_hi + off) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4*>(A_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
}

__global__ void __launch_bounds__(NTHREADS, 1) sdf_tc_kernel(const MlpParams P, const TcNet T) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* A_hi = smem;
    uint8_t* A_lo = smem + A_PART_BYTES;
    uint8_t* ring = smem + 2 * A_PART_BYTES;
    float* stash = reinterpret_cast<float*>(ring + NSTAGE * STAGE_BYTES);    // [TM][STASH_LD] embedding / sqrt2
    float* hpart = stash + TM * STASH_LD;                                      // [4][TM] partial sdf heads
    uint64_t* bars = reinterpret_cast<uint64_t*>(hpart + 4 * TM);
    uint64_t* full = bars;                 // [NSTAGE]
    uint64_t* empty = bars + NSTAGE;       // [NSTAGE]
    uint64_t* a_ready = bars + 2 * NSTAGE; // [8]  (4 arrivals: one per contributing warp)
    uint64_t* d_full = a_ready + 8;        // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d_full + 2);

    if (!round_active(P)) return;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const NetDev& net = P.net;
    const int NL = T.NL;
    const long long ntiles = (P.M + TM - 1) / TM;

    if (tid == 0) {
        for (int i = 0; i < NSTAGE; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 8; ++i) mbar_init(&a_ready[i], 4);
        mbar_init(&d_full[0], 1);
        mbar_init(&d_full[1], 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================= weight producer =================
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                for (int l = 0; l < NL; ++l) {
                    const uint8_t* src = T.wpack + (size_t)T.layer_off[l] * STAGE_BYTES;
                    for (int ks = 0; ks < T.ksteps[l]; ++ks) {
                        mbar_wait(&empty[stage], phase ^ 1);
                        mbar_arrive_expect_tx(&full[stage], STAGE_BYTES);
                        bulk_g2s(ring + stage * STAGE_BYTES, src + (size_t)ks * STAGE_BYTES, STAGE_BYTES, &full[stage]);
                        if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            const uint32_t idesc = instr_desc_bf16(TM, NCOL);
            const uint32_t a_hi_s = smem_u32(A_hi), a_lo_s = smem_u32(A_lo), ring_s = smem_u32(ring);
            uint32_t stage = 0, phase = 0, aphase = 0;
            for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                for (int l = 0; l < NL; ++l) {
                    const uint32_t d_tmem = tmem_base + (uint32_t)(l & 1) * 256u;
                    const int nks = T.ksteps[l];
                    for (int ks = 0; ks < nks; ++ks) {
                        if ((ks & 1) == 0) {
                            const int c = ks >> 1;
                            mbar_wait(&a_ready[c], (aphase >> c) & 1u);
                            aphase ^= (1u << c);
                        }
                        mbar_wait(&full[stage], phase);
                        tc_fence_after();
                        const uint32_t a_off = (uint32_t)ks * 2u * LBO_A;
                        const uint32_t b_s = ring_s + stage * STAGE_BYTES;
                        const uint64_t da_hi = smem_desc(a_hi_s + a_off, LBO_A, SBO);
                        const uint64_t da_lo = smem_desc(a_lo_s + a_off, LBO_A, SBO);
                        const uint64_t db_hi = smem_desc(b_s, LBO_B, SBO);
                        const uint64_t db_lo = smem_desc(b_s + STAGE_BYTES / 2, LBO_B, SBO);
                        mma_bf16_ss(d_tmem, da_hi, db_hi, idesc, ks > 0 ? 1u : 0u);
                        mma_bf16_ss(d_tmem, da_lo, db_hi, idesc, 1u);
                        mma_bf16_ss(d_tmem, da_hi, db_lo, idesc, 1u);
                        mma_commit(&empty[stage]);          // stage reusable once these MMAs have read it
                        if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
                    }
                    mma_commit(&d_full[l & 1]);             // accumulator of layer l complete
                }
            }
        }
    } else {
        // ================= epilogue warps =================
        const int q = warp & 3;                              // TMEM lane quarter this warp may access
        const int sub = (warp - 2) >> 2;                     // 0..3: owns 32-column chunks sub and sub + 4
        const int row = q * 32 + lane;                       // point within the tile
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        const int nsplit = 256 - net.ex;
        uint32_t dphase = 0;                                 // bit b: parity to wait on d_full[b]
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const long long m = tile * TM + row;
            // ---- prologue: positional encoding -> A_0 (48 columns) + stash of embed/sqrt2.  sub 0: columns 0..31,
            //      sub 1: columns 32..47; subs 2,3 idle.
            if (sub < 2) {
                float x[3] = {0.f, 0.f, 0.f};
                if (m < P.M) {
                    if (P.pts) { x[0] = P.pts[m * 3]; x[1] = P.pts[m * 3 + 1]; x[2] = P.pts[m * 3 + 2]; }
                    else {
                        long long r = m / P.ns;
                        int j = (int)(m - r * P.ns);
                        float t = P.zarr[r * P.zstride + j];
#pragma unroll
                        for (int c = 0; c < 3; ++c) x[c] = __fadd_rn(P.ray_o[r * 3 + c], __fmul_rn(t, P.ray_d[r * 3 + c]));
                    }
                }
                float hv[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int i = sub * 32 + j;
                    hv[j] = (i < net.ex) ? embed_col(x, i, net.mx) : 0.f;
                    if (i < STASH_LD) stash[row * STASH_LD + i] = hv[j] * 0.70710678118654752f;
                }
                if (sub == 0) store_a_chunk(A_hi, A_lo, row, 0, hv);
                else {
#pragma unroll
                    for (int s4 = 0; s4 < 2; ++s4) {
                        uint32_t h[4], lo[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) split_bf16x2(hv[s4 * 8 + 2 * i], hv[s4 * 8 + 2 * i + 1], h[i], lo[i]);
                        const uint32_t off = seg_off<TM>(row, 4 + s4);
                        *reinterpret_cast<uint4*>(A_hi + off) = make_uint4(h[0], h[1], h[2], h[3]);
                        *reinterpret_cast<uint4*>(A_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                    }
                }
                fence_proxy_async();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&a_ready[sub]);
            }
            epi_bar_sync();                                  // stash visible to the warps that fill the skip columns

            float head = 0.f;
            for (int l = 0; l < NL; ++l) {
                const int b = l & 1;
                mbar_wait(&d_full[b], (dphase >> b) & 1u);
                dphase ^= (1u << b);
                tc_fence_after();
                const bool last = (l == NL - 1);
                const bool feeds_skip = (l + 1 == net.skip);
                const float* __restrict__ bias = net.sdf_b[l];
#pragma unroll 1
                for (int cc = 0; cc < 2; ++cc) {
                    const int c = sub + 4 * cc;
                    uint32_t v[32];
                    tmem_ld32(tmem_base + lane_base + (uint32_t)b * 256u + (uint32_t)c * 32u, v);
                    tmem_ld_wait();
                    float hv[32];
#pragma unroll
                    for (int j4 = 0; j4 < 8; ++j4) {
                        const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + c * 32) + j4);
                        hv[j4 * 4 + 0] = softplus100_fast(__uint_as_float(v[j4 * 4 + 0]) + bb.x);
                        hv[j4 * 4 + 1] = softplus100_fast(__uint_as_float(v[j4 * 4 + 1]) + bb.y);
                        hv[j4 * 4 + 2] = softplus100_fast(__uint_as_float(v[j4 * 4 + 2]) + bb.z);
                        hv[j4 * 4 + 3] = softplus100_fast(__uint_as_float(v[j4 * 4 + 3]) + bb.w);
                    }
                    if (last) {
#pragma unroll
                        for (int j4 = 0; j4 < 8; ++j4) {
                            const float4 w = __ldg(reinterpret_cast<const float4*>(net.sdf_head + c * 32) + j4);
                            head = fmaf(hv[j4 * 4 + 0], w.x, head);
                            head = fmaf(hv[j4 * 4 + 1], w.y, head);
                            head = fmaf(hv[j4 * 4 + 2], w.z, head);
                            head = fmaf(hv[j4 * 4 + 3], w.w, head);
                        }
                        continue;
                    }
                    if (feeds_skip) {     // cat([h, embed]) / sqrt(2)   (mlp.py:94-95)
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const int f = c * 32 + j;
                            hv[j] = (f >= nsplit) ? stash[row * STASH_LD + (f - nsplit)] : hv[j] * 0.70710678118654752f;
                        }
                    }
                    store_a_chunk(A_hi, A_lo, row, c * 4, hv);
                    fence_proxy_async();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&a_ready[c]);
                }
            }
            // ---- sdf head: combine the 4 column-partials of every row
            hpart[sub * TM + row] = head;
            epi_bar_sync();
            if (sub == 0 && m < P.M)
                P.out_sdf[m] = ((hpart[row] + hpart[TM + row]) + (hpart[2 * TM + row] + hpart[3 * TM + row])) + __ldg(net.sdf_head + 256);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// ---- weight packing: fp32 [out,in] -> per-kstep blocks of bf16 hi / lo in the canonical layout ----------
__global__ void pack_tc_kernel(uint8_t* __restrict__ dst, const float* __restrict__ W, int outd, int in, int ksteps) {
    // one thread per (kstep, chunk, row): writes 8 hi + 8 lo values
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int total = ksteps * 2 * 256;
    if (i >= total) return;
    int n = i % 256, chunk = (i / 256) % 2, ks = i / 512;
    uint16_t hi[8], lo[8];
    for (int e = 0; e < 8; ++e) {
        int k = ks * 16 + chunk * 8 + e;
        float w = (n < outd && k < in) ? W[(size_t)n * in + k] : 0.f;
        __nv_bfloat16 h = __float2bfloat16_rn(w);
        __nv_bfloat16 l = __float2bfloat16_rn(w - __bfloat162float(h));
        hi[e] = *reinterpret_cast<uint16_t*>(&h);
        lo[e] = *reinterpret_cast<uint16_t*>(&l);
    }
    uint8_t* base = dst + (size_t)ks * STAGE_BYTES + (size_t)chunk * 4096 + (size_t)n * 16;
    *reinterpret_cast<uint4*>(base) = *reinterpret_cast<uint4*>(hi);
    *reinterpret_cast<uint4*>(base + STAGE_BYTES / 2) = *reinterpret_cast<uint4*>(lo);
}

struct TcState {
    uint8_t* wpack;
    TcNet net;
};

}  // namespace tcsdf

int tc_create(i2sdf_handle* h) {
    using namespace tcsdf;
    TcState* s = new TcState();
    const NetDev& n = h->net;
    s->net.NL = n.L - 1;
    int off = 0;
    for (int l = 0; l < n.L - 1; ++l) {
        s->net.ksteps[l] = (l == 0) ? K0_STEPS : 16;
        s->net.layer_off[l] = off;
        off += s->net.ksteps[l];
    }
    if (cudaMalloc(&s->wpack, (size_t)off * STAGE_BYTES) != cudaSuccess) {
        delete s;
        set_error("tc_create: cudaMalloc failed");
        return I2SDF_E_CUDA;
    }
    s->net.wpack = s->wpack;
    h->tc = s;
    cudaError_t e = cudaFuncSetAttribute(sdf_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
    if (e != cudaSuccess) { set_error("tc_create: smem attribute: %s", cudaGetErrorString(e)); return I2SDF_E_CUDA; }
    return I2SDF_OK;
}

void tc_destroy(i2sdf_handle* h) {
    tcsdf::TcState* s = (tcsdf::TcState*)h->tc;
    if (!s) return;
    cudaFree(s->wpack);
    delete s;
    h->tc = nullptr;
}

int tc_pack(i2sdf_handle* h, const float* const* W, const float* const* b, cudaStream_t st) {
    using namespace tcsdf;
    (void)b;   // biases / head reuse the fp32 arrays packed for the SIMT path
    TcState* s = (TcState*)h->tc;
    for (int l = 0; l < s->net.NL; ++l) {
        int total = s->net.ksteps[l] * 512;
        pack_tc_kernel<<<(total + 255) / 256, 256, 0, st>>>(s->wpack + (size_t)s->net.layer_off[l] * STAGE_BYTES, W[l], h->lay_out[l],
                                                            h->lay_in[l], s->net.ksteps[l]);
    }
    I2SDF_CUDA_CHECK(cudaGetLastError());
    return I2SDF_OK;
}

int tc_launch_sdf(const i2sdf_handle* h, const MlpParams& p, cudaStream_t st) {
    using namespace tcsdf;
    if (p.M <= 0) return I2SDF_OK;
    const TcState* s = (const TcState*)h->tc;
    long long ntiles = (p.M + TM - 1) / TM;
    int grid = (int)(ntiles < (long long)h->num_sms ? ntiles : (long long)h->num_sms);
    sdf_tc_kernel<<<grid, NTHREADS, kSmemBytes, st>>>(p, s->net);
    I2SDF_CUDA_CHECK(cudaGetLastError());
    return I2SDF_OK;
}

}  // namespace i2sdf
