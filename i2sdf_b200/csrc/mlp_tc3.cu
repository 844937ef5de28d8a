// Tensor-core MLP chain (tcgen05 / TMEM), sm_100a — unified kernel, version 3.
//
// One persistent CTA per SM walks a resident tile of 128 ray points through a table of dense "ops" (one op = one
// layer = ksteps x 3 tcgen05.mma):
//     sdf-only (sampler):   F_0 .. F_{NL-1}                                   -> sdf
//     full main pass:       F_0 .. F_{NL-1}, G, C_0 .. C_{Lc-2}, R_{NL-1} .. R_1, R_0   -> sdf, rgb, grad_x sdf
//   F  SDF hidden layers        epilogue: +b, softplus_100 (+skip concat) [full: h~ also to its slot / the per-CTA scratch]
//   G  feature rows of last layer           +b ; appends PE(view dir) as k chunk 8
//   C  radiance hidden layers               +b, ReLU ; last: rgb head (fp32 dots) + reverse prologue w_sdf * softplus'
//   R  reverse sweep with W^T               (skip split) * softplus'_{l-1} = 1 - exp(-100 h_{l-1}) ;  R_0: J(x)^T r -> grad_x sdf
// Activations never leave the SM: they are the SMEM A operand (fp16 hi + fp16 lo, canonical K-major layout); weights
// stream from L2 by cp.async.bulk into a 4-stage ring; accumulators are fp32 in TMEM, two buffers ping-ponged by op;
// every MAC is 3 fp16 products  A_hi W_hi + A_lo W_hi + A_hi W_lo  (11 + 11 mantissa bits, weights pre-scaled by 2^8 so their low
// halves stay normal: as accurate as an fp32 FMA chain, tools/probe_fmt.cu; round 1 used bf16 halves = 16 bits, 5x the error, which
// the sampler's discrete decisions amplified - profiles/parity_r02.json).  Training slots stay bf16 hi/lo (the backward chain's format).
//
// Warp roles: warp 0 = weight producer, warp 1 = MMA issuer, warps 2..17 = epilogue.  An epilogue warp (q, sub) owns TMEM
// lane quarter q and, in iteration it = 0..3, the 16 columns  (2 it + sub/2) * 32 + (sub%2) * 16 .. +16 : the eight
// warps working on a 32-column chunk finish it together and chunks complete IN ORDER, so the MMA warp (which waits per
// chunk) trails the epilogue by one chunk pair and layer l+1's tensor work overlaps layer l's epilogue on one tile.
//
// Replaces (reference): ImplicitNetwork.get_sdf_vals mlp.py:145-151 as called by the sampler (ray_sampler.py:84-89);
// model/network/__init__.py:103-116 = ImplicitNetwork.get_outputs mlp.py:123-143 (forward :84-105 + autograd.grad
// :134-140) and RenderingNetwork.forward mlp.py:208-229; Embedder embedder.py:28-38.
#include <stdlib.h>
#include <type_traits>
#include "common.cuh"
#include "tc_common.cuh"
#include "tc_chain.cuh"
#include "tc_bwd.cuh"

namespace i2sdf {
namespace tc3 {

using namespace tc;
using namespace chain;

__device__ __forceinline__ bool round_active(const MlpParams& P) {
    if (P.round_idx <= 0) return true;
    float b0 = fabsf(*P.beta_param) + P.beta_min;
    for (int j = 0; j < P.round_idx; ++j)
        if (!(P.beta_max[j] > b0)) return false;
    return true;
}

// ---- full chain (main pass, eikonal points, sdf + features) ---------------------------------------------------------------------
// Work items are 32 rows x 8 columns as in the sampler's kernel below: in every iteration all 16 epilogue warps work on ONE 32-column
// chunk, so the next op's MMA chain starts after 1/8 of an epilogue (round 1 used 16-column items, 8 warps per chunk: the main pass
// ran at ~16 k clocks per op against ~8.4 k for the sampler's kernel).  Each epilogue kind has its own specialisation of the item loop.
// Shared memory: A_hi | A_lo (36 chunks each) | weight ring | parameters | barriers.
//   * parameters: every bias and both heads, copied once per launch, read as LDS broadcasts (ncu, round 2: with LDG the first FFMA of
//     an item carried 17 % of all stall samples - the bias load missed L1 behind the scratch traffic);
//   * the skip concat and the Jacobians of the reverse sweep re-evaluate sincos, but only in the items that reach into the embedding
//     columns (warp-uniform branch; round 1's if-converted form evaluated it for all 256 columns);
//   * the per-row partial sums (head, rgb, grad_x: [4 column groups][7][128]) alias operand chunks 32..35 (PE(view dir), consumed by
//     C_0 long before the tile ends).
enum { K_F = 0, K_FSKIP, K_FLAST, K_FLASTREV, K_G, K_C, K_CLAST, K_REV, K_REVSKIP, K_GRAD };
constexpr int M8_PARAM_FLOATS = 36 * TM;            // 4608: (NL + 1 + Lc - 1) x 256 biases + sdf head (257 -> 260) + rgb head (771 -> 772)
constexpr size_t kSmemMain = 128 + 2 * (size_t)A_PART_BYTES + NSTAGE * STAGE_MAX + M8_PARAM_FLOATS * 4 + 256;
static_assert(kSmemMain <= 232448, "main pass shared memory");

// TL: development probe (I2SDF_DEBUG_TIMELINE=1, tools/timeline.py main): clock64 stamps of CTA 0's second tile into P.tl
template <bool SAVE, bool TL = false>
// 18 warps -> one scheduler hosts 5 of them: 16 K regs / 5 warps caps the kernel at 96 registers per thread
__global__ void __launch_bounds__(NTHREADS, 1) tc_main8_kernel(const MlpParams P, const OpTable T) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_align128(smem_raw);
    uint8_t* A_hi = smem;
    uint8_t* A_lo = smem + A_PART_BYTES;
    uint8_t* ring = smem + 2 * A_PART_BYTES;
    float* sprm = reinterpret_cast<float*>(ring + NSTAGE * STAGE_MAX);      // [sdf_b NL x 256 | feat_b 256 | col_b (Lc-1) x 256 | sdf_head 260 | col_head 772]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sprm + M8_PARAM_FLOATS);
    uint64_t* full = bars;
    uint64_t* empty = bars + NSTAGE;
    uint64_t* a_ready = bars + 2 * NSTAGE;       // [N_READY], 16 arrivals each
    uint64_t* d_full = a_ready + N_READY;        // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d_full + 2);
    float* part01 = reinterpret_cast<float*>(A_hi + 32 * TM * 16);          // [2][7][TM]
    float* part23 = reinterpret_cast<float*>(A_lo + 32 * TM * 16);          // [2][7][TM]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const NetDev& net = P.net;
    const int NL = net.L - 1;
    const long long ntiles = (P.M + TM - 1) / TM;

    if (tid == 0) {
        for (int i = 0; i < NSTAGE; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < N_READY; ++i) mbar_init(&a_ready[i], N_EPI_WARPS);
        mbar_init(&d_full[0], 1);
        mbar_init(&d_full[1], 1);
        fence_mbar_init();
    }
    const int Lc1 = net.Lc - 1;
    float* s_sdf_b = sprm;
    float* s_feat_b = sprm + NL * 256;
    float* s_col_b = s_feat_b + 256;
    float* s_sdf_head = s_col_b + Lc1 * 256;
    float* s_col_head = s_sdf_head + 260;
    for (int i = tid; i < NL * 256; i += NTHREADS) s_sdf_b[i] = net.sdf_b[i >> 8][i & 255];
    for (int i = tid; i < 256; i += NTHREADS) s_feat_b[i] = net.sdf_b[net.L - 1][i];
    for (int i = tid; i < Lc1 * 256; i += NTHREADS) s_col_b[i] = net.col_b[i >> 8][i & 255];
    for (int i = tid; i < 257; i += NTHREADS) s_sdf_head[i] = net.sdf_head[i];
    for (int i = tid; i < 771; i += NTHREADS) s_col_head[i] = net.col_head[i];
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) chain_producer(T, ntiles, ring, full, empty);
    } else if (warp == 1) {
        chain_mma<TL>(T, ntiles, tmem_base, A_hi, A_lo, ring, full, empty, a_ready, d_full, reinterpret_cast<long long*>(P.tl));
    } else {
        // ================= epilogue warps =================
        const int q = warp & 3;
        const int sub = (warp - 2) >> 2;                     // column group: columns 8 sub .. 8 sub + 7 of every 32-column chunk
        const int row = q * 32 + lane;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        const int nsplit = 256 - net.ex;
        const float RS2 = 0.70710678118654752f, C1 = 144.26950408889634f, C2 = 0.0069314718055994531f;
        const float acc_scale = T.acc_scale;
        // The reverse sweep rebuilds softplus'(a_l) from h~_l, which round-trips through HBM/L2 as hi/lo plane segments: a per-CTA
        // scratch in eval (fp16 halves: the bytes the SMEM operand gets); in training (plane slots, P.sl) every A operand is stored
        // per point (h~_l, q_l, features, radiance activations, encodings) as bf16 halves for the backward chain and the weight
        // gradients.
        constexpr bool save = SAVE;               // separate instantiation: the eval kernel carries none of the slot code
        const planes::Layout& SL = P.sl;
        uint32_t dphase = 0, g = 0;
        float x[3], dv[3];
        // point (and view direction) of this thread's row in a tile
        auto load_point = [&](long long tile) {
            const long long m = tile * TM + row;
            x[0] = x[1] = x[2] = 0.f; dv[0] = 0.f; dv[1] = 0.f; dv[2] = 1.f;
            if (m < P.M) {
                const long long r = m / P.ns;
                const bool from_pts = P.pts && m >= P.m_rays;
                const float* pp = P.pts + (m - P.m_rays) * 3;
                if (P.ray_d && !from_pts) { dv[0] = P.ray_d[r * 3]; dv[1] = P.ray_d[r * 3 + 1]; dv[2] = P.ray_d[r * 3 + 2]; }
                if (from_pts) { x[0] = pp[0]; x[1] = pp[1]; x[2] = pp[2]; }
                else {
                    const int j = (int)(m - r * P.ns);
                    const float t = P.zarr[r * P.zstride + j];
#pragma unroll
                    for (int c = 0; c < 3; ++c) x[c] = __fadd_rn(P.ray_o[r * 3 + c], __fmul_rn(t, dv[c]));
                }
            }
        };
        // prologue: A_0 = embedding, 48 columns = k chunks 0..5: group sub writes chunk sub, groups 0 and 1 also chunks 4 and 5
        auto prologue = [&](long long tile) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int kc = sub + 4 * h;
                if (kc < 6) {
                    float hv[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int i = kc * 8 + j;
                        hv[j] = (i < net.ex) ? embed_col(x, i, net.mx) : 0.f;
                    }
                    uint8_t* gs = save ? SL.base + SL.E() + planes::seg(tile * TM + row, kc, planes::SMALL_CHUNKS) : nullptr;
                    store_a8<true, !save>(A_hi, A_lo, row, kc, hv, gs, true, true, (uint32_t)planes::SMALL_PLANE);
                }
            }
            fence_proxy_async();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { mbar_arrive(&a_ready[0]); mbar_arrive(&a_ready[1]); }
        };
        if ((long long)blockIdx.x < ntiles) { load_point(blockIdx.x); prologue(blockIdx.x); }
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const long long m = tile * TM + row;
            const bool valid = m < P.M;
            const long long next_tile = tile + gridDim.x;
            // h~_l storage the reverse sweep reads back: training = H slots (by point), eval = per-CTA scratch holding ONE tile (by row);
            // null for the sdf + features table (no reverse sweep)
            uint8_t* hbase = nullptr;
            size_t hstride = 0;
            long long hm = 0;
            if (save) { hbase = SL.base + SL.H(0); hstride = SL.big; hm = m; }
            else if (P.scratch) { hbase = reinterpret_cast<uint8_t*>(P.scratch) + (size_t)blockIdx.x * (size_t)NL * planes::BIG_TILE; hstride = planes::BIG_TILE; hm = row; }

            float head = 0.f, rgbp[3] = {0.f, 0.f, 0.f}, gacc[3] = {0.f, 0.f, 0.f};
            for (int op = 0; op < T.nops; ++op, ++g) {
                const uint32_t b = g & 1u;
                const int kind = T.ops[op].kind, l = T.ops[op].layer;
                const bool last_op = (op == T.nops - 1);
                long long* tlw = (TL && P.tl && lane == 0 && blockIdx.x == 0 && tile == (long long)gridDim.x)
                                     ? reinterpret_cast<long long*>(P.tl) + (op * 16 + (warp - 2)) * 4 : nullptr;
                if (TL && tlw) tlw[0] = clock64();
                mbar_wait(&d_full[b], (dphase >> b) & 1u);
                dphase ^= (1u << b);
                tc_fence_after();
                if (TL && tlw) tlw[1] = clock64();
                const uint32_t acc_addr = tmem_base + lane_base + b * 256u;
                if (hbase && !last_op && P.pf_op_ahead) {
                    // the next op's stored h~ (written up to 20 ops ago, possibly evicted to HBM): pull this thread's 16 segments into L2 now
                    const int nk = T.ops[op + 1].kind, nl = T.ops[op + 1].layer;
                    if (nk == EK_REV || nk == EK_COL_LAST) {
                        const uint8_t* ns = hbase + (size_t)(nk == EK_COL_LAST ? NL - 1 : nl - 1) * hstride + planes::seg(hm, sub, planes::BIG_CHUNKS);
#pragma unroll
                        for (int it = 0; it < 8; ++it) { pf_l2(ns + (size_t)it * 4 * planes::SUB_CHUNK); pf_l2(ns + (size_t)it * 4 * planes::SUB_CHUNK + planes::BIG_PLANE); }
                    }
                }

                auto items = [&](auto kind_c) {
                    constexpr int K = decltype(kind_c)::value;
                    constexpr bool is_f = (K == K_F || K == K_FSKIP || K == K_FLAST || K == K_FLASTREV);
                    constexpr bool has_bias = is_f || K == K_G || K == K_C || K == K_CLAST;
                    constexpr bool loads_h = (K == K_CLAST || K == K_REV || K == K_REVSKIP);
                    const float* __restrict__ bias = nullptr;
                    if (is_f) bias = s_sdf_b + l * 256;
                    else if (K == K_G) bias = s_feat_b;
                    else if (K == K_C || K == K_CLAST) bias = s_col_b + l * 256;
                    const uint8_t* hsrc = nullptr;             // stored h~ the item needs: h_{NL-1} (reverse prologue) / h~_{l-1}
                    if (loads_h) hsrc = hbase + (size_t)(K == K_CLAST ? NL - 1 : l - 1) * hstride;
                    const float hs = (K == K_REVSKIP) ? C1 * 1.41421356237309505f : C1;
                    // stored h~ segments travel one item ahead in registers (they come from L2 / HBM: the scratch of 148 CTAs exceeds L2)
                    uint4 ph = make_uint4(0, 0, 0, 0), pl = ph;
                    if (loads_h) {
                        const uint8_t* ps = hsrc + planes::seg(hm, sub, planes::BIG_CHUNKS);
                        ph = ldg_cs(ps);
                        pl = ldg_cs(ps + planes::BIG_PLANE);
                    }
#pragma unroll 2
                    for (int it = 0; it < (K == K_GRAD ? 2 : 8); ++it) {
                        const int col0 = it * 32 + sub * 8, kc = it * 4 + sub;
                        if (K == K_GRAD && col0 >= 48) break;
                        uint32_t v[8];
                        tmem_ld8(acc_addr + (uint32_t)col0, v);
                        uint4 nph = ph, npl = pl;
                        float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
                        if (loads_h && it < 7) {
                            const uint8_t* ps = hsrc + planes::seg(hm, kc + 4, planes::BIG_CHUNKS);
                            nph = ldg_cs(ps);
                            npl = ldg_cs(ps + planes::BIG_PLANE);
                        }
                        if (has_bias) { b0 = *reinterpret_cast<const float4*>(bias + col0); b1 = *reinterpret_cast<const float4*>(bias + col0 + 4); }
                        tmem_ld_wait();
                        const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                        float hv[8];
                        uint32_t ch[4], cl[4];                 // K_CLAST in training: the split last radiance activation, stored behind the publish
                        uint8_t* gs = nullptr;                 // slot / scratch segment this item stores to
                        bool keep = true;
                        if (is_f) {
                            float sp[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const float a = fmaf(__uint_as_float(v[j]), acc_scale, bv[j]);
                                const float e = ex2_approx(-fabsf(a) * C1);          // exp(-|100 a|)
                                hv[j] = fmaf(lg2_approx(1.0f + e), C2, fmaxf(a, 0.0f));
                                if (K == K_FLASTREV) {
                                    const float rr = rcp_approx(1.0f + e);
                                    sp[j] = (a >= 0.f) ? rr : e * rr;                 // softplus'(a) = sigmoid(100 a)
                                }
                            }
                            if (K == K_FSKIP) {                // cat([h, embed]) / sqrt(2)   (mlp.py:94-95)
                                if (col0 + 8 > nsplit) {       // warp-uniform: only the items that reach into the embedding part
#pragma unroll
                                    for (int j = 0; j < 8; ++j) {
                                        const int f = col0 + j;
                                        if (f >= nsplit) hv[j] = embed_col(x, f - nsplit, net.mx);
                                    }
                                }
#pragma unroll
                                for (int j = 0; j < 8; ++j) hv[j] *= RS2;
                            }
                            if (K == K_FLAST || K == K_FLASTREV) {
                                const float4 w0 = *reinterpret_cast<const float4*>(s_sdf_head + col0), w1 = *reinterpret_cast<const float4*>(s_sdf_head + col0 + 4);
                                const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                                for (int j = 0; j < 8; ++j) head = fmaf(hv[j], w[j], head);
                                if (K == K_FLASTREV) {
                                    // sdf + grad only: h_{NL-1} to its slot, then straight into the reverse sweep: q = w_sdf * softplus'
                                    if (save) store_a8<true, false>(A_hi, A_lo, row, kc, hv, SL.base + SL.H(l) + planes::seg(m, kc, planes::BIG_CHUNKS), true, false);
#pragma unroll
                                    for (int j = 0; j < 8; ++j) hv[j] = w[j] * sp[j];
                                }
                            }
                            if (K == K_FLASTREV) { if (save) { gs = SL.base + SL.Q(NL - 1) + planes::seg(m, kc, planes::BIG_CHUNKS); keep = valid; } }
                            else if (save) gs = SL.base + SL.H(l) + planes::seg(m, kc, planes::BIG_CHUNKS);
                            else if (hbase) gs = hbase + (size_t)l * hstride + planes::seg(hm, kc, planes::BIG_CHUNKS);
                        } else if (K == K_G) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) hv[j] = fmaf(__uint_as_float(v[j]), acc_scale, bv[j]);
                            if (last_op) {                     // sdf + features only: nothing follows
                                if (P.out_feat && valid) {
                                    float4* o4 = reinterpret_cast<float4*>(P.out_feat + (size_t)m * 256 + col0);
                                    o4[0] = make_float4(hv[0], hv[1], hv[2], hv[3]);
                                    o4[1] = make_float4(hv[4], hv[5], hv[6], hv[7]);
                                }
                                continue;
                            }
                            if (save) gs = SL.base + SL.CF() + planes::seg(m, kc, planes::BIG_CHUNKS);
                        } else if (K == K_C || K == K_CLAST) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) hv[j] = fmaxf(fmaf(__uint_as_float(v[j]), acc_scale, bv[j]), 0.f);
                            if (K == K_C) { if (save) gs = SL.base + SL.C(l) + planes::seg(m, kc, planes::BIG_CHUNKS); }
                            else {
                                const float* __restrict__ wh = s_col_head + col0;
#pragma unroll
                                for (int c = 0; c < 3; ++c) {
                                    const float4 w0 = *reinterpret_cast<const float4*>(wh + c * 256), w1 = *reinterpret_cast<const float4*>(wh + c * 256 + 4);
                                    rgbp[c] = fmaf(hv[0], w0.x, fmaf(hv[1], w0.y, fmaf(hv[2], w0.z, fmaf(hv[3], w0.w, rgbp[c]))));
                                    rgbp[c] = fmaf(hv[4], w1.x, fmaf(hv[5], w1.y, fmaf(hv[6], w1.z, fmaf(hv[7], w1.w, rgbp[c]))));
                                }
                                // last radiance activation: slot only (stored behind the publish below); then the reverse prologue: adjoint of
                                // a_{NL-1} = w_sdf * softplus'(a_{NL-1})
                                if (save) {
#pragma unroll
                                    for (int i = 0; i < 4; ++i) split_bf16x2(hv[2 * i], hv[2 * i + 1], ch[i], cl[i]);
                                }
                                seg8_values<!save>(ph, pl, hv);                       // h_{NL-1}
                                const float4 w0 = *reinterpret_cast<const float4*>(s_sdf_head + col0), w1 = *reinterpret_cast<const float4*>(s_sdf_head + col0 + 4);
                                const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                                for (int j = 0; j < 8; ++j) hv[j] = w[j] * (1.0f - ex2_approx(-C1 * hv[j]));
                                if (save) { gs = SL.base + SL.Q(NL - 1) + planes::seg(m, kc, planes::BIG_CHUNKS); keep = valid; }
                            }
                        } else if (K == K_REV || K == K_REVSKIP) {
                            // accumulator = adjoint of the input of SDF layer l ; next A = (that) * softplus'(a_{l-1}), softplus' from the
                            // stored h~_{l-1} (a skip concat stored it scaled by 1/sqrt2; its embedding part goes to grad_x through J(x)^T)
                            float sv[8];
                            seg8_values<!save>(ph, pl, sv);
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const float rv = __uint_as_float(v[j]) * (K == K_REVSKIP ? RS2 * acc_scale : acc_scale);
                                hv[j] = (rv != 0.f) ? rv * (1.0f - ex2_approx(-hs * sv[j])) : 0.f;
                            }
                            if (K == K_REVSKIP && col0 + 8 > nsplit) {     // warp-uniform: the items that reach into the embedding part
#pragma unroll
                                for (int j = 0; j < 8; ++j) {
                                    const int f = col0 + j;
                                    if (f >= nsplit) {
                                        int coord;
                                        const float jac = embed_jac(x, f - nsplit, net.mx, coord);
                                        const float rv = __uint_as_float(v[j]) * (RS2 * acc_scale);
                                        gacc[0] += (coord == 0) ? jac * rv : 0.f;
                                        gacc[1] += (coord == 1) ? jac * rv : 0.f;
                                        gacc[2] += (coord == 2) ? jac * rv : 0.f;
                                        hv[j] = 0.f;
                                    }
                                }
                            }
                            if (save) { gs = SL.base + SL.Q(l - 1) + planes::seg(m, kc, planes::BIG_CHUNKS); keep = valid; }
                        } else {                               // K_GRAD: accumulator columns 0..47 = adjoint of the embedding
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const int i = col0 + j;
                                if (i < net.ex) {
                                    int coord;
                                    const float jac = embed_jac(x, i, net.mx, coord);
                                    const float rv = __uint_as_float(v[j]) * acc_scale;
                                    gacc[0] += (coord == 0) ? jac * rv : 0.f;
                                    gacc[1] += (coord == 1) ? jac * rv : 0.f;
                                    gacc[2] += (coord == 2) ? jac * rv : 0.f;
                                }
                            }
                            continue;
                        }
                        {
                            // SMEM operand first, published at once; the slot / scratch copy leaves behind the fence
                            uint32_t hh[4], ll[4];
                            sts_a8<true>(A_hi, A_lo, row, kc, hv, hh, ll);
                            publish_chunk(&a_ready[it], lane);
                            if (gs) stg_a8<!save>(gs, (uint32_t)planes::BIG_PLANE, keep, hv, hh, ll);
                            if (K == K_CLAST && save) {
                                uint8_t* cs = SL.base + SL.C(l) + planes::seg(m, kc, planes::BIG_CHUNKS);
                                *reinterpret_cast<uint4*>(cs) = make_uint4(ch[0], ch[1], ch[2], ch[3]);
                                *reinterpret_cast<uint4*>(cs + planes::BIG_PLANE) = make_uint4(cl[0], cl[1], cl[2], cl[3]);
                            }
                            if (K == K_G && P.out_feat && valid) {
                                float4* o4 = reinterpret_cast<float4*>(P.out_feat + (size_t)m * 256 + col0);
                                o4[0] = make_float4(hv[0], hv[1], hv[2], hv[3]);
                                o4[1] = make_float4(hv[4], hv[5], hv[6], hv[7]);
                            }
                        }
                        ph = nph; pl = npl;
                        if (TL && tlw && it == 0) tlw[2] = clock64();
                    }
                    if (TL && tlw) tlw[3] = clock64();
                };
                // the A operand is free once the tile's last MMAs are done: start the NEXT tile's first layer inside the last epilogue
                // (behind the items when they still need this tile's point: the Jacobians of the last reverse op)
                const bool start_next = last_op && next_tile < ntiles;
                if (start_next && kind != EK_GRAD) { load_point(next_tile); prologue(next_tile); }
                switch (kind) {
                    case EK_SDF_HIDDEN: if (l + 1 == net.skip) items(std::integral_constant<int, K_FSKIP>{}); else items(std::integral_constant<int, K_F>{}); break;
                    case EK_SDF_LAST: items(std::integral_constant<int, K_FLAST>{}); break;
                    case EK_SDF_LAST_REV: items(std::integral_constant<int, K_FLASTREV>{}); break;
                    case EK_FEAT: items(std::integral_constant<int, K_G>{}); break;
                    case EK_COL_HIDDEN: items(std::integral_constant<int, K_C>{}); break;
                    case EK_COL_LAST: items(std::integral_constant<int, K_CLAST>{}); break;
                    case EK_REV: if (l == net.skip) items(std::integral_constant<int, K_REVSKIP>{}); else items(std::integral_constant<int, K_REV>{}); break;
                    default: items(std::integral_constant<int, K_GRAD>{}); break;
                }
                if (kind == EK_FEAT && !last_op) {
                    // k chunk 8 (columns 256..287) = positional encoding of the view direction, zero padded: group sub writes 8 columns
                    float hv[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int i = sub * 8 + j;
                        hv[j] = (i < net.ed) ? embed_col(dv, i, net.md) : 0.f;
                    }
                    uint8_t* gs = save ? SL.base + SL.DV() + planes::seg(m, sub, planes::DV_CHUNKS) : nullptr;
                    store_a8<true, !save>(A_hi, A_lo, row, 32 + sub, hv, gs, true, true, (uint32_t)(planes::DV_CHUNKS * planes::SUB_CHUNK));
                    publish_chunk(&a_ready[8], lane);
                }
                if (start_next && kind == EK_GRAD) { load_point(next_tile); prologue(next_tile); }
            }
            // ---- combine the 4 column-group partials of every row and write the per-sample results
            float* pp = (sub < 2 ? part01 : part23) + (size_t)(sub & 1) * 7 * TM;
            pp[row] = head;
            pp[TM + row] = rgbp[0]; pp[2 * TM + row] = rgbp[1]; pp[3 * TM + row] = rgbp[2];
            pp[4 * TM + row] = gacc[0]; pp[5 * TM + row] = gacc[1]; pp[6 * TM + row] = gacc[2];
            epi_bar_sync();
            if (sub == 0 && m < P.M) {
                float acc[7];
#pragma unroll
                for (int k = 0; k < 7; ++k)
                    acc[k] = (part01[k * TM + row] + part01[(7 + k) * TM + row]) + (part23[k * TM + row] + part23[(7 + k) * TM + row]);
                P.out_sdf[m] = acc[0] + s_sdf_head[256];
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    if (P.out_rgb) {
                        const float s = acc[1 + c] + s_col_head[768 + c];
                        P.out_rgb[m * 3 + c] = __fdiv_rn(1.0f, 1.0f + expf(-s));
                    }
                    if (P.out_grad) P.out_grad[m * 3 + c] = acc[4 + c];
                }
            }
            epi_bar_sync();     // part[] free for the next tile
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// ---- sdf-only chain with 8-column work items (the sampler's kernel) ---------------------------------------------------------
// Same chain as the full kernel's F ops, but an epilogue work item is 32 rows x 8 columns: in every iteration all 16 epilogue warps
// work on ONE 32-column chunk (4 lane quarters x 4 column groups), so the first chunk of the next A operand - and with it the MMA
// chain of the next layer - is ready after 1/8 of the epilogue.
// Own shared-memory layout (K <= 256: 32 operand chunks): A_hi | A_lo | weight ring | PE stash | biases | partial heads | barriers.
//   * PE stash [39][128] fp32: the positional encoding of the tile's points, written once by the tile's prologue and read by the
//     layer that feeds the skip concat.  Round 2's timeline (tools/timeline.py) showed that op taking 14.8 k clocks against 8.4 k
//     for the others: the concat re-evaluated a Cody-Waite sincos per element (if-converted, so for ALL 256 columns).
//   * biases of all hidden layers, copied once per launch: LDS broadcast instead of two LDG.128 per item on the critical path
//     between "accumulator complete" and "first chunk published".
constexpr int S8_A_PART = 32 * TM * 16;                          // 65536
constexpr int S8_PE_FLOATS = 40 * TM;                            // [ex <= 39][128]
constexpr int S8_BIAS_FLOATS = (kMaxLayers - 1) * 256;
constexpr size_t kSmemSdf8 = 1024 + 2 * (size_t)S8_A_PART + NSTAGE * STAGE_MAX + (S8_PE_FLOATS + S8_BIAS_FLOATS + 4 * TM) * 4 + 256;
static_assert(kSmemSdf8 <= 232448, "sdf8 shared memory");

// TL: development probe (I2SDF_DEBUG_TIMELINE=1, tools/timeline.py) - clock64 stamps of CTA 0's second tile into P.scratch:
// epilogue warp w, op: tl[(op * 16 + w) * 4 + {0 waits for the accumulator, 1 sees it, 2 first chunk published, 3 last chunk published}]
template <bool TL>
__global__ void __launch_bounds__(NTHREADS, 1) tc_sdf8_kernel(const MlpParams P, const OpTable T) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_align1024(smem_raw);
    uint8_t* A_hi = smem;
    uint8_t* A_lo = smem + S8_A_PART;
    uint8_t* ring = smem + 2 * S8_A_PART;
    float* pe = reinterpret_cast<float*>(ring + NSTAGE * STAGE_MAX);        // [40][TM]
    float* sbias = pe + S8_PE_FLOATS;                                        // [NL][256]
    float* part = sbias + S8_BIAS_FLOATS;                                    // [4][TM]
    uint64_t* bars = reinterpret_cast<uint64_t*>(part + 4 * TM);
    uint64_t* full = bars;
    uint64_t* empty = bars + NSTAGE;
    uint64_t* a_ready = bars + 2 * NSTAGE;       // [N_READY], 16 arrivals each (every epilogue warp contributes to every chunk)
    uint64_t* d_full = a_ready + N_READY;        // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d_full + 2);

    if (!round_active(P)) return;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const NetDev& net = P.net;
    const int NL = net.L - 1;
    const long long ntiles = (P.M + TM - 1) / TM;

    if (tid == 0) {
        for (int i = 0; i < NSTAGE; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < N_READY; ++i) mbar_init(&a_ready[i], N_EPI_WARPS);
        mbar_init(&d_full[0], 1);
        mbar_init(&d_full[1], 1);
        fence_mbar_init();
    }
    for (int i = tid; i < NL * 256; i += NTHREADS) sbias[i] = net.sdf_b[i >> 8][i & 255];
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) chain_producer(T, ntiles, ring, full, empty);
    } else if (warp == 1) {
        chain_mma<TL>(T, ntiles, tmem_base, A_hi, A_lo, ring, full, empty, a_ready, d_full, reinterpret_cast<long long*>(P.scratch));
    } else {
        const int q = warp & 3;
        const int sub = (warp - 2) >> 2;                     // column group: columns 8 sub .. 8 sub + 7 of every 32-column chunk
        const int row = q * 32 + lane;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        const int nsplit = 256 - net.ex;
        const float RS2 = 0.70710678118654752f;
        const float acc_scale = T.acc_scale;
        uint32_t dphase = 0, g = 0;
        auto load_point = [&](long long tile, float (&x)[3]) {
            const long long m = tile * TM + row;
            x[0] = x[1] = x[2] = 0.f;
            if (m < P.M) {
                if (P.gx) {                       // regular grid, generated here (utils/plots.py:440-451 order; model/eval/recon.py:80-84 affine)
                    const long long ji = m / P.nz;
                    const int k = (int)(m - ji * P.nz), j = (int)(ji / P.nx), i = (int)(ji - (long long)j * P.nx);
                    const float px = __ldg(P.gx + i), py = __ldg(P.gy + j), pz = __ldg(P.gz + k);
                    x[0] = px; x[1] = py; x[2] = pz;
                    if (P.grid_affine) {
                        const float* A = P.grid_affine;
#pragma unroll
                        for (int c = 0; c < 3; ++c)
                            x[c] = __fadd_rn(fmaf(__ldg(A + c * 3 + 2), pz, fmaf(__ldg(A + c * 3 + 1), py, __fmul_rn(__ldg(A + c * 3), px))), __ldg(A + 9 + c));
                    }
                } else if (P.pts && m >= P.m_rays) { const float* pp = P.pts + (m - P.m_rays) * 3; x[0] = pp[0]; x[1] = pp[1]; x[2] = pp[2]; }
                else {
                    const long long r = m / P.ns;
                    const int j = (int)(m - r * P.ns);
                    const float t = P.zarr[r * P.zstride + j];
#pragma unroll
                    for (int c = 0; c < 3; ++c) x[c] = __fadd_rn(P.ray_o[r * 3 + c], __fmul_rn(t, P.ray_d[r * 3 + c]));
                }
            }
        };
        // A_0 = embedding, 48 columns = k chunks 0..5: group sub writes chunk sub, groups 0 and 1 also chunks 4 and 5; every value
        // also goes to the PE stash (the previous tile's skip layer is long done when this runs inside its last op)
        auto prologue = [&](const float (&x)[3]) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int kc = sub + 4 * h;
                if (kc < 6) {
                    float hv[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int i = kc * 8 + j;
                        hv[j] = (i < net.ex) ? embed_col(x, i, net.mx) : 0.f;
                        if (i < 40) pe[i * TM + row] = hv[j];
                    }
                    store_a8<true>(A_hi, A_lo, row, kc, hv, nullptr, true, true);
                }
            }
            fence_proxy_async();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { mbar_arrive(&a_ready[0]); mbar_arrive(&a_ready[1]); }
        };
        float x[3];
        if ((long long)blockIdx.x < ntiles) { load_point(blockIdx.x, x); prologue(x); }
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const long long m = tile * TM + row;
            const long long next_tile = tile + gridDim.x;
            float head = 0.f;
            for (int op = 0; op < T.nops; ++op, ++g) {
                const uint32_t b = g & 1u;
                const int l = T.ops[op].layer;
                const bool last = (T.ops[op].kind == EK_SDF_LAST);
                const float* __restrict__ bias = sbias + l * 256 + sub * 8;
                long long* tlw = (TL && P.scratch && lane == 0 && blockIdx.x == 0 && tile == (long long)gridDim.x)
                                     ? reinterpret_cast<long long*>(P.scratch) + (op * 16 + (warp - 2)) * 4 : nullptr;
                if (TL && tlw) tlw[0] = clock64();
                mbar_wait(&d_full[b], (dphase >> b) & 1u);
                dphase ^= (1u << b);
                tc_fence_after();
                if (TL && tlw) tlw[1] = clock64();
                // one work item = 32 rows x 8 columns; three specialisations of the item loop so that the common one carries neither the
                // head dot (last layer) nor the skip concat (mode: 0 hidden layer, 1 hidden layer feeding the skip concat, 2 last)
                auto items = [&](auto mode_c) {
                    constexpr int MODE = decltype(mode_c)::value;
#pragma unroll 2
                    for (int it = 0; it < 8; ++it) {
                        const int col0 = it * 32 + sub * 8;
                        uint32_t v[8];
                        tmem_ld8(tmem_base + lane_base + b * 256u + (uint32_t)col0, v);
                        const float4 b0 = *reinterpret_cast<const float4*>(bias + it * 32), b1 = *reinterpret_cast<const float4*>(bias + it * 32 + 4);
                        tmem_ld_wait();
                        const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                        float hv[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float a = fmaf(__uint_as_float(v[j]), acc_scale, bv[j]);
                            const float e = ex2_approx(-fabsf(a) * 144.26950408889634f);          // exp(-|100 a|)
                            hv[j] = fmaf(lg2_approx(1.0f + e), 0.0069314718055994531f, fmaxf(a, 0.0f));
                        }
                        if (MODE == 2) {
                            const float4 w0 = __ldg(reinterpret_cast<const float4*>(net.sdf_head + col0)), w1 = __ldg(reinterpret_cast<const float4*>(net.sdf_head + col0 + 4));
                            head = fmaf(hv[0], w0.x, fmaf(hv[1], w0.y, fmaf(hv[2], w0.z, fmaf(hv[3], w0.w, head))));
                            head = fmaf(hv[4], w1.x, fmaf(hv[5], w1.y, fmaf(hv[6], w1.z, fmaf(hv[7], w1.w, head))));
                            continue;
                        }
                        if (MODE == 1) {                           // cat([h, embed]) / sqrt(2)   (mlp.py:94-95)
                            if (col0 + 8 > nsplit) {               // warp-uniform: only the items that reach into the embedding part
#pragma unroll
                                for (int j = 0; j < 8; ++j) {
                                    const int f = col0 + j;
                                    if (f >= nsplit) hv[j] = pe[(f - nsplit) * TM + row];
                                }
                            }
#pragma unroll
                            for (int j = 0; j < 8; ++j) hv[j] *= RS2;
                        }
                        store_a8<true>(A_hi, A_lo, row, col0 >> 3, hv, nullptr, true, true);
                        publish_chunk(&a_ready[it], lane);
                        if (TL && tlw && it == 0) tlw[2] = clock64();
                    }
                    if (TL && tlw) tlw[3] = clock64();
                };
                if (last) {
                    if (next_tile < ntiles) {      // the A operand is free (this tile's last MMAs are done): start the next tile's first layer now
                        load_point(next_tile, x);
                        prologue(x);
                    }
                    items(std::integral_constant<int, 2>{});
                } else if (l + 1 == net.skip) items(std::integral_constant<int, 1>{});
                else items(std::integral_constant<int, 0>{});
            }
            // ---- combine the 4 column-group partials of every row
            part[sub * TM + row] = head;
            epi_bar_sync();
            if (sub == 0 && m < P.M)
                P.out_sdf[m] = (part[row] + part[TM + row]) + (part[2 * TM + row] + part[3 * TM + row]) + __ldg(net.sdf_head + 256);
            epi_bar_sync();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// ---- weight packing -------------------------------------------------------------------------------------
// dst block per k step: [part hi|lo][chunk 0|1][n rows][8 bf16].  Element (n, k):
//   mode 0 (forward):  W[(n + row_off) * in + col(k)]   col(k) = k, or for the radiance input layer
//                      k < feat ? ed + k : k - feat  (our A operand is [feat | PE(dir)], the reference's [PE(dir) | feat])
//   mode 1 (reverse):  W[(k + row_off) * in + n + col_off]   (B = W^T: n = input index, k = output index)
// Two images of every block: fp16 hi/lo of W * kWScale (forward chains) at dst16, bf16 hi/lo of W (backward chain, tc_gemm.cu) at dst.
struct TcPackJob { uint8_t* dst; uint8_t* dst16; const float* W; int outd, in, ksteps, n_rows, mode, row_off, feat_first, pad; };
struct TcPackBatch { int n; int ed; TcPackJob jobs[MAX_OPS]; };
// one launch for all weight blocks (grid.y = block): re-packing follows every optimizer step in training
__global__ void pack_kernel(const TcPackBatch B) {
    const TcPackJob& J = B.jobs[blockIdx.y];
    const int total = J.ksteps * 2 * J.n_rows;
    const float* __restrict__ W = J.W;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int n_rows = J.n_rows, in = J.in, outd = J.outd, row_off = J.row_off, feat_first = J.feat_first, ed = B.ed;
        int n = i % n_rows, chunk = (i / n_rows) % 2, ks = i / (2 * n_rows);
        uint16_t hi[8], lo[8], hi16[8], lo16[8];
        for (int e = 0; e < 8; ++e) {
            int k = ks * 16 + chunk * 8 + e;
            float w = 0.f;
            if (J.mode == 0) {
                int col = k;
                if (feat_first > 0) col = (k < feat_first) ? ed + k : ((k - feat_first < ed) ? k - feat_first : -1);
                if (col >= 0 && col < in && n + row_off < outd) w = W[(size_t)(n + row_off) * in + col];
            } else {
                // reverse: feat_first doubles as a column offset (radiance layer 0: only the feature columns are needed)
                if (k + row_off < outd && n + feat_first < in) w = W[(size_t)(k + row_off) * in + n + feat_first];
            }
            __nv_bfloat16 h = __float2bfloat16_rn(w);
            __nv_bfloat16 l = __float2bfloat16_rn(w - __bfloat162float(h));
            hi[e] = *reinterpret_cast<uint16_t*>(&h);
            lo[e] = *reinterpret_cast<uint16_t*>(&l);
            uint32_t h2, l2;
            split_f16x2(w * kWScale, 0.f, h2, l2);
            hi16[e] = (uint16_t)(h2 & 0xffffu);
            lo16[e] = (uint16_t)(l2 & 0xffffu);
        }
        const size_t sb = (size_t)n_rows * 64;
        const size_t o = (size_t)ks * sb + (size_t)chunk * n_rows * 16 + (size_t)n * 16;
        *reinterpret_cast<uint4*>(J.dst + o) = *reinterpret_cast<uint4*>(hi);
        *reinterpret_cast<uint4*>(J.dst + o + sb / 2) = *reinterpret_cast<uint4*>(lo);
        *reinterpret_cast<uint4*>(J.dst16 + o) = *reinterpret_cast<uint4*>(hi16);
        *reinterpret_cast<uint4*>(J.dst16 + o + sb / 2) = *reinterpret_cast<uint4*>(lo16);
    }
}

// The tensor core's fp32 accumulator rounds TOWARD ZERO at every accumulation step (tools/probe_fmt.cu on the B200: with all-positive
// operands the 48 steps of a K = 256 layer leave every sum 1.8e-6 too small; with zero-mean weights the least-squares shrink is
// 8.3e-7).  Through nine layers that is a systematic sdf error of -2.3e-6 near the surface - 30x the fp32 SIMT kernel's - which the
// sampler's exp(-sdf / beta) with beta = 0.01 turns into flipped bisection decisions.  The expected shrink is removed by scaling
// the accumulator by (1 + kappa) in the epilogue (folded into the constant of the bias FFMA, no extra instruction).
static float acc_scale() {
    double kappa = 8.3e-7;
    if (const char* e = getenv("I2SDF_KAPPA")) kappa = atof(e);      // development: tools/sdf_error.py sweeps it
    return (float)((double)kWInv * (1.0 + kappa));
}

struct State {
    uint8_t* wpack;           // bf16 hi/lo blocks (backward chain, tc_gemm.cu)
    uint8_t* wpack16;         // fp16 hi/lo blocks of W * kWScale at the same offsets (forward chains)
    OpTable sdf;              // ops of the sdf-only chain (a prefix of the full table)
    OpTable full;             // nops == 0 if the full main pass is unavailable for this network
    OpTable sg;               // SDF + grad_x only (eikonal points): F_0..F_{NL-1}, R_{NL-1}..R_0
    OpTable sf;               // SDF + features (ImplicitNetwork.forward: meshing / plots): F_0..F_{NL-1}, G
    OpTable bwd_full;         // training backward chain (mlp_tc_bwd.cu): T_0..T_{NL-1}, CR.., CR_0, PF, P_{NL-1}..P_1
    OpTable bwd_sdf;          // same without the radiance stack (eikonal points): T.., PF, P..
    int src_layer[MAX_OPS];   // packing recipe per op of the full table
    int mode[MAX_OPS], row_off[MAX_OPS], feat_first[MAX_OPS];
    int n_pack;               // number of ops to pack
    // weight blocks by role, for the backward GEMMs (tc_gemm.cu): index into full.ops, -1 = absent
    int blk_fwd_sdf[kMaxLayers], blk_fwd_feat, blk_fwd_col[kMaxLayers], blk_rev_sdf[kMaxLayers], blk_rev_feat, blk_rev_col[kMaxLayers];
    int blk_fwd_light;        // light head layer 0 (N = light_hidden), consumed by tc_gemm_pw (the head runs as its own small pass)
};

}  // namespace tc3

int tc_create(i2sdf_handle* h) {
    using namespace tc3;
    State* s = new State();
    const NetDev& n = h->net;
    const int L = n.L, NL = L - 1, Lc = n.Lc;
    // (the light-mask head is not an op of the chain: it reads the features the main pass writes, see light_forward in backward.cu)
    // (the full chain keeps every bias and both heads in shared memory: (NL + Lc) x 256 + 1032 floats must fit M8_PARAM_FLOATS)
    bool want_full = ((NL + Lc) * 256 + 1032 <= M8_PARAM_FLOATS);
#ifdef I2SDF_CHECK_BUILD
    if (getenv("I2SDF_SIMT_MAIN") && getenv("I2SDF_SIMT_MAIN")[0] == '1') want_full = false;      // cross-check: main pass on the fp32 kernel
#endif
    OpTable& T = s->full;
    int nops = 0;
    size_t off = 0;
    auto add = [&](int ksteps, int nn, int kind, int layer, int src, int mode, int row_off, int feat_first) {
        Op& o = T.ops[nops];
        o.w_off = (int)off; o.ksteps = (short)ksteps; o.n = (short)nn; o.kind = (short)kind; o.layer = (short)layer;
        s->src_layer[nops] = src; s->mode[nops] = mode; s->row_off[nops] = row_off; s->feat_first[nops] = feat_first;
        off += (size_t)ksteps * nn * 64;
        ++nops;
    };
    for (int i = 0; i < kMaxLayers; ++i) s->blk_fwd_sdf[i] = s->blk_fwd_col[i] = s->blk_rev_sdf[i] = s->blk_rev_col[i] = -1;
    s->blk_fwd_feat = s->blk_rev_feat = s->blk_fwd_light = -1;
    for (int l = 0; l < NL; ++l) { s->blk_fwd_sdf[l] = nops; add(l == 0 ? 3 : 16, 256, l == NL - 1 ? EK_SDF_LAST : EK_SDF_HIDDEN, l, l, 0, 0, 0); }
    s->blk_fwd_feat = nops;
    add(16, 256, EK_FEAT, L - 1, L - 1, 0, 1, 0);
    for (int l = 0; l < Lc - 1; ++l) { s->blk_fwd_col[l] = nops; add(l == 0 ? 18 : 16, 256, l == Lc - 2 ? EK_COL_LAST : EK_COL_HIDDEN, l, L + l, 0, 0, l == 0 ? 256 : 0); }
    for (int l = NL - 1; l >= 1; --l) { s->blk_rev_sdf[l] = nops; add(16, 256, EK_REV, l, l, 1, 0, 0); }
    add(16, 48, EK_GRAD, 0, 0, 1, 0, 0);
    const int n_kernel_ops = nops;
    // pack-only blocks used by the training backward
    s->blk_rev_feat = nops;
    add(16, 256, -1, L - 1, L - 1, 1, 1, 0);                                   // B[j][i] = W_last[1 + i][j]
    for (int l = Lc - 2; l >= 1; --l) { s->blk_rev_col[l] = nops; add(16, 256, -1, l, L + l, 1, 0, 0); }
    s->blk_rev_col[0] = nops;
    add(16, 256, -1, 0, L, 1, 0, n.ed);                                          // feature columns of radiance layer 0
    if (n.Ll == 2) { s->blk_fwd_light = nops; add(16, n.lh, -1, 0, L + Lc, 0, 0, 0); }   // light layer 0: [lh out][256 in]
    s->n_pack = nops;
    T.nops = want_full ? n_kernel_ops : 0;
    if (cudaMalloc(&s->wpack, 2 * off) != cudaSuccess) { delete s; set_error("tc_create: cudaMalloc failed"); return I2SDF_E_CUDA; }
    s->wpack16 = s->wpack + off;
    T.wpack = s->wpack16;
    T.f16 = 1;
    T.acc_scale = acc_scale();
    s->sdf = T;
    s->sdf.nops = NL;
    s->sg.wpack = s->wpack16;
    s->sg.f16 = 1;
    s->sg.acc_scale = T.acc_scale;
    s->sg.nops = 0;
    for (int l = 0; l < NL; ++l) { s->sg.ops[s->sg.nops] = T.ops[s->blk_fwd_sdf[l]]; if (l == NL - 1) s->sg.ops[s->sg.nops].kind = EK_SDF_LAST_REV; ++s->sg.nops; }
    for (int l = NL - 1; l >= 1; --l) s->sg.ops[s->sg.nops++] = T.ops[s->blk_rev_sdf[l]];
    s->sg.ops[s->sg.nops++] = T.ops[n_kernel_ops - 1];          // R_0 (EK_GRAD)
    s->sf.wpack = s->wpack16;
    s->sf.f16 = 1;
    s->sf.acc_scale = T.acc_scale;
    s->sf.nops = 0;
    for (int l = 0; l < NL; ++l) s->sf.ops[s->sf.nops++] = T.ops[s->blk_fwd_sdf[l]];
    s->sf.ops[s->sf.nops++] = T.ops[s->blk_fwd_feat];
    // backward chains reuse the forward / reverse blocks: tangent = forward SDF blocks, adjoints = W^T blocks
    for (int variant = 0; variant < 2; ++variant) {
        OpTable& B = variant == 0 ? s->bwd_full : s->bwd_sdf;
        B.wpack = s->wpack;
        B.f16 = 0;
        B.nops = 0;
        auto push = [&](int idx, int kind, int layer) { B.ops[B.nops] = T.ops[idx]; B.ops[B.nops].kind = (short)kind; B.ops[B.nops].layer = (short)layer; ++B.nops; };
        for (int l = 0; l < NL; ++l) push(s->blk_fwd_sdf[l], tcb::BK_TAN, l);
        if (variant == 0) {
            for (int l = Lc - 2; l >= 1; --l) push(s->blk_rev_col[l], tcb::BK_COL_REV, l);
            push(s->blk_rev_col[0], tcb::BK_FEAT_ADJ, 0);
        }
        push(s->blk_rev_feat, tcb::BK_P, NL - 1);
        for (int l = NL - 1; l >= 1; --l) push(s->blk_rev_sdf[l], tcb::BK_P, l - 1);
        if (variant == 0 && !want_full) B.nops = 0;
    }
    cudaError_t e = cudaFuncSetAttribute(tc_sdf8_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemSdf8);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tc_sdf8_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemSdf8);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tc_main8_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMain);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tc_main8_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMain);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tc_main8_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMain);
    if (e != cudaSuccess) { cudaFree(s->wpack); delete s; set_error("tc_create: smem attribute: %s", cudaGetErrorString(e)); return I2SDF_E_CUDA; }
    h->tc = s;
    h->tcmain = (void*)s;     // full main pass only if s->full.nops > 0 (tcmain_has_full)
    return I2SDF_OK;
}

void tc_destroy(i2sdf_handle* h) {
    tc3::State* s = (tc3::State*)h->tc;
    if (!s) return;
    cudaFree(s->wpack);
    delete s;
    h->tc = nullptr;
    h->tcmain = nullptr;
}

int tc_pack(i2sdf_handle* h, const float* const* W, const float* const* b, cudaStream_t st) {
    using namespace tc3;
    (void)b;   // biases / heads reuse the fp32 arrays packed for the fp32 path
    State* s = (State*)h->tc;
    TcPackBatch B{};
    B.n = s->n_pack; B.ed = h->net.ed;
    for (int op = 0; op < s->n_pack; ++op) {
        const Op& o = s->full.ops[op];
        const int li = s->src_layer[op];
        B.jobs[op] = TcPackJob{s->wpack + o.w_off, s->wpack16 + o.w_off, W[li], h->lay_out[li], h->lay_in[li], o.ksteps, o.n, s->mode[op], s->row_off[op], s->feat_first[op], 0};
    }
    pack_kernel<<<dim3(36, B.n), 256, 0, st>>>(B);
    I2SDF_CUDA_CHECK(cudaGetLastError());
    return I2SDF_OK;
}

static inline int tc_grid(const i2sdf_handle* h, long long M) {
    long long ntiles = (M + tc3::TM - 1) / tc3::TM;
    return balanced_grid(h->num_sms, ntiles);
}

const chain::OpTable* tc_bwd_table(const i2sdf_handle* h, bool with_color) {
    const tc3::State* s = (const tc3::State*)h->tc;
    if (!s) return nullptr;
    return with_color ? &s->bwd_full : &s->bwd_sdf;
}

TcBlock tc_block(const i2sdf_handle* h, int role, int layer) {
    using namespace tc3;
    TcBlock b{nullptr, 0, 0};
    const State* s = (const State*)h->tc;
    if (!s) return b;
    int idx = -1;
    switch (role) {
        case TCB_FWD_SDF: idx = s->blk_fwd_sdf[layer]; break;
        case TCB_FWD_FEAT: idx = s->blk_fwd_feat; break;
        case TCB_FWD_COL: idx = s->blk_fwd_col[layer]; break;
        case TCB_REV_SDF: idx = s->blk_rev_sdf[layer]; break;
        case TCB_REV_FEAT: idx = s->blk_rev_feat; break;
        case TCB_REV_COL: idx = s->blk_rev_col[layer]; break;
        case TCB_FWD_LIGHT: idx = s->blk_fwd_light; break;
    }
    if (idx < 0) return b;
    const Op& o = s->full.ops[idx];
    b.ptr = s->wpack + o.w_off; b.ksteps = o.ksteps; b.n = o.n;
    return b;
}

int tc_launch_sdf(const i2sdf_handle* h, const MlpParams& p, cudaStream_t st) {
    using namespace tc3;
    if (p.M <= 0) return I2SDF_OK;
    const State* s = (const State*)h->tc;
    static int tl = -1;
    if (tl < 0) tl = getenv("I2SDF_DEBUG_TIMELINE") ? 1 : 0;
    if (tl && p.scratch) tc_sdf8_kernel<true><<<tc_grid(h, p.M), NTHREADS, kSmemSdf8, st>>>(p, s->sdf);
    else tc_sdf8_kernel<false><<<tc_grid(h, p.M), NTHREADS, kSmemSdf8, st>>>(p, s->sdf);
    I2SDF_CUDA_CHECK(cudaGetLastError());
    return I2SDF_OK;
}

// full main pass (kept under the tcmain_* names used by c_abi.cu)
int tcmain_create(i2sdf_handle*, void** out_state) { *out_state = nullptr; return I2SDF_OK; }
void tcmain_destroy(void*) {}
int tcmain_pack(i2sdf_handle*, void*, const float* const*, cudaStream_t) { return I2SDF_OK; }
int tcmain_has_full(const void* state) { return state && ((const tc3::State*)state)->full.nops > 0; }
int tcmain_launch(const i2sdf_handle* h, void* state, const MlpParams& p, cudaStream_t st) {
    using namespace tc3;
    if (p.M <= 0) return I2SDF_OK;
    const State* s = (const State*)state;
    const OpTable& tab = p.want_color ? s->full : (p.out_grad ? s->sg : s->sf);
    // op-ahead L2 prefetch of the stored h~ (round 2, first sessions: on) measured against off on the 1024-ray C2 step
    // (profiles/r02b_main_prefetch.txt): training main pass 1.110 vs 1.024 ms - as in the backward chain, a whole op of 134 CTAs ahead
    // evicts more than it saves; the one-item-ahead register prefetch of the items loop stays.  I2SDF_MAIN_PREFETCH=1 turns it back on.
    static const int pf = [] { const char* e = getenv("I2SDF_MAIN_PREFETCH"); return (e && e[0] == '1') ? 1 : 0; }();
    MlpParams q = p;
    q.pf_op_ahead = pf;
    if (p.sl.base) tc_main8_kernel<true><<<tc_grid(h, p.M), NTHREADS, kSmemMain, st>>>(q, tab);
    else if (p.tl) tc_main8_kernel<false, true><<<tc_grid(h, p.M), NTHREADS, kSmemMain, st>>>(q, tab);
    else tc_main8_kernel<false><<<tc_grid(h, p.M), NTHREADS, kSmemMain, st>>>(q, tab);
    I2SDF_CUDA_CHECK(cudaGetLastError());
    return I2SDF_OK;
}

}  // namespace i2sdf
