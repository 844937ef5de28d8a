// Training backward of the main pass as ONE tensor-core chain kernel (tcgen05 / TMEM), sm_100a.
//
// Same skeleton as the forward (tc_chain.cuh): a resident tile of 128 points walks through a table of dense ops, the A
// operand never leaves the SM between ops.  Per tile:
//     T_0 .. T_{NL-1}          tangent forward of the SDF stack along gbar (upstream of grad_x sdf):
//                              adot_l = W_l hdot~_{l-1},  hdot_l = softplus'(a_l) * adot_l            -> slots HD(l)
//     [ CR_{Lc-2} .. CR_1      radiance stack reverse: pc_{l-1} = (W_l^T pc_l) * [c_{l-1} > 0]         -> slots PC(l)
//       CR_0 ]                 fbar = W_0[:, feat]^T pc_0                                              -> slot FB
//     PF, P_{NL-1} .. P_1      SDF stack reverse carrying first- and second-order terms:
//                              u = W_{l+1}^T p_{l+1}  (PF: W_feat^T fbar + sbar w_sdf)
//                              p_l = s'(a_l) u + s''(a_l) adot_l v_l ,  with  q_l = s'(a_l) v_l  saved by the forward      -> slots P(l)
// Everything softplus-related is rebuilt from the stored activation h_l = softplus(a_l):  s' = 1 - exp(-100 h),
// s'' adot v = 100 (1 - s') hdot_l q_l / s'.  Operands that the forward saved (H, Q, C slots) and the tangents written
// earlier by the same thread (HD) are read as 16-byte plane segments; the adjoints leave as plane slots that the
// weight-gradient kernel (wgrad_planes.cu) consumes without conversion.  No dense activation array in fp32 exists.
//
// Replaces (reference): what torch.autograd runs for loss.backward() through ImplicitNetwork.forward / .gradient
// (mlp.py:84-143, incl. the create_graph second-order graph) and RenderingNetwork.forward (mlp.py:208-229) — every
// per-point matrix product except the weight gradients themselves.
#include <stdlib.h>
#include "common.cuh"
#include "tc_common.cuh"
#include "tc_chain.cuh"
#include "tc_bwd.cuh"

namespace i2sdf {
namespace tcb {

using namespace tc;
using namespace chain;

// the four 16-byte segments (hi / lo planes x two chunks) of a 16-column item
__device__ __forceinline__ void pf_seg16(const uint8_t* seg) {
    pf_l2(seg); pf_l2(seg + planes::SUB_CHUNK); pf_l2(seg + planes::BIG_PLANE); pf_l2(seg + planes::BIG_PLANE + planes::SUB_CHUNK);
}

template <bool PF_NEXT>
__global__ void __launch_bounds__(NTHREADS, 1) tc_bwd_kernel(const BwdParams P, const OpTable T) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_align1024(smem_raw);
    uint8_t* A_hi = smem;
    uint8_t* A_lo = smem + A_PART_BYTES;
    uint8_t* ring = smem + 2 * A_PART_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(ring + NSTAGE * STAGE_MAX + PART_FLOATS * 4);
    uint64_t* full = bars;
    uint64_t* empty = bars + NSTAGE;
    uint64_t* a_ready = bars + 2 * NSTAGE;       // [N_READY], 8 arrivals each
    uint64_t* d_full = a_ready + N_READY;        // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d_full + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const NetDev& net = P.net;
    const int NL = net.L - 1;
    const long long ntiles = (P.M + TM - 1) / TM;

    if (tid == 0) {
        for (int i = 0; i < NSTAGE; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < N_READY; ++i) mbar_init(&a_ready[i], 8);
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
        if (lane == 0) chain_producer(T, ntiles, ring, full, empty);
    } else if (warp == 1) {
        chain_mma(T, ntiles, tmem_base, A_hi, A_lo, ring, full, empty, a_ready, d_full);
    } else {
        // ================= epilogue warps =================
        const int q = warp & 3;
        const int sub = (warp - 2) >> 2;
        const int row = q * 32 + lane;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        const int nsplit = 256 - net.ex;
        const float RS2 = 0.70710678118654752f, S2 = 1.41421356237309505f;
        const planes::Layout& SL = P.sl;
        const bool color = P.with_color != 0;
        uint32_t dphase = 0, g = 0;
        // point of this thread's row and the upstream of grad_x sdf there
        auto load_point = [&](long long tile, float (&x)[3], float (&gb)[3]) {
            const long long m = tile * TM + row;
            x[0] = x[1] = x[2] = 0.f; gb[0] = gb[1] = gb[2] = 0.f;
            if (m < P.M) {
                if (P.pts && m >= P.m_rays) { const float* pp = P.pts + (m - P.m_rays) * 3; x[0] = pp[0]; x[1] = pp[1]; x[2] = pp[2]; }
                else {
                    const long long r = m / P.ns;
                    const int j = (int)(m - r * P.ns);
                    const float t = P.zarr[r * P.zstride + j];
#pragma unroll
                    for (int c = 0; c < 3; ++c) x[c] = __fadd_rn(P.ray_o[r * 3 + c], __fmul_rn(t, P.ray_d[r * 3 + c]));
                }
                if (P.g_grad) { gb[0] = P.g_grad[m * 3]; gb[1] = P.g_grad[m * 3 + 1]; gb[2] = P.g_grad[m * 3 + 2]; }
            }
        };
        // prologue: A_0 = tangent of the embedding, J(x) gbar, 48 columns (sub s: columns 16 s ..), also slot ED
        auto prologue = [&](const float (&x)[3], const float (&gb)[3], long long tile) {
            if (sub < 3) {
                float hv[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int i = sub * 16 + j;
                    int coord = 0;
                    const float jac = (i < net.ex) ? embed_jac(x, i, net.mx, coord) : 0.f;
                    hv[j] = jac * (coord == 0 ? gb[0] : (coord == 1 ? gb[1] : gb[2]));
                }
                store_a16<false>(A_hi, A_lo, row, sub * 2, hv);
                publish_chunk(&a_ready[sub >> 1], lane);
                // slot copies leave BEHIND the publish: a fence.proxy.async behind a global store waits for that store (MEMBAR.ALL.CTA)
                store_a16<false>(A_hi, A_lo, row, sub * 2, hv, SL.wbase + SL.ED() + planes::seg(tile * TM + row, sub * 2, planes::SMALL_CHUNKS),
                                 (uint32_t)planes::SMALL_PLANE, true, false);
            } else {
                publish_chunk(&a_ready[sub >> 1], lane);
            }
        };
        float x[3], gb[3];
        if ((long long)blockIdx.x < ntiles) { load_point(blockIdx.x, x, gb); prologue(x, gb, blockIdx.x); }
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const long long m = tile * TM + row;
            const bool valid = m < P.M;
            const long long next_tile = tile + gridDim.x;
            float xn[3] = {0.f, 0.f, 0.f}, gbn[3] = {0.f, 0.f, 0.f};
            // per-point upstream scalars
            const float sbar = (valid && P.g_sdf) ? P.g_sdf[m] : 0.f;
            float delta[3] = {0.f, 0.f, 0.f};
            if (color && valid) {
#pragma unroll
                for (int c = 0; c < 3; ++c) { const float y = P.s_rgb[m * 3 + c]; delta[c] = P.g_rgb[m * 3 + c] * y * (1.f - y); }
            }
            for (int op = 0; op < T.nops; ++op, ++g) {
                const uint32_t b = g & 1u;
                const int kind = T.ops[op].kind, l = T.ops[op].layer;
                mbar_wait(&d_full[b], (dphase >> b) & 1u);
                dphase ^= (1u << b);
                tc_fence_after();
                const bool last_op = (op == T.nops - 1);
                if (last_op && next_tile < ntiles) {
                    // the A operand is free (this tile's last MMAs are done): start the next tile's first op now
                    load_point(next_tile, xn, gbn);
                    prologue(xn, gbn, next_tile);
                }
#pragma unroll 1
                for (int it = 0; it < 4; ++it) {
                    const int c = 2 * it + (sub >> 1);                 // 32-column chunk
                    const int col0 = c * 32 + (sub & 1) * 16;          // first of this warp's 16 columns
                    const int kc0 = col0 >> 3;
                    const size_t sg = planes::seg(m, kc0, planes::BIG_CHUNKS);
                    if (PF_NEXT) {
                        // this warp's NEXT item (same op, or the first item of the next op) reads 2..6 slot segments per 8 columns
                        // straight from HBM with nothing else to hide the latency behind: pull them into L2 one item ahead
                        int pk = kind, pl = l, pit = it + P.pf_dist;
                        if (pit >= 4) { pit -= 4; if (op + 1 < T.nops) { pk = T.ops[op + 1].kind; pl = T.ops[op + 1].layer; } else pk = -1; }
                        const size_t psg = planes::seg(m, ((2 * pit + (sub >> 1)) * 32 + (sub & 1) * 16) >> 3, planes::BIG_CHUNKS);
                        if (pk == BK_P) {
                            pf_seg16(SL.base + SL.H(pl) + psg); pf_seg16(SL.base + SL.Q(pl) + psg); pf_seg16(SL.wbase + SL.HD(pl) + psg);
                        } else if (pk == BK_TAN) {
                            pf_seg16(SL.base + SL.H(pl) + psg);
                        } else if (pk == BK_COL_REV) {
                            pf_l2(SL.base + SL.C(pl - 1) + psg); pf_l2(SL.base + SL.C(pl - 1) + psg + planes::SUB_CHUNK);
                        }
                    }
                    if (kind == BK_TAN) {
                        // hdot_l = softplus'(a_l) * adot_l   (skip concat: [hdot | J gbar] / sqrt2)
                        const bool feeds_skip = (l + 1 == net.skip);
                        uint4 raw[4];
                        load_slot16(SL.base + SL.H(l) + sg, raw);
                        uint32_t v[16];
                        tmem_ld16(tmem_base + lane_base + b * 256u + (uint32_t)col0, v);
                        tmem_ld_wait();
                        float hv[16];
                        slot16_values(raw, hv);
                        const float hs = feeds_skip ? 144.26950408889634f * S2 : 144.26950408889634f;
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            float t = (1.0f - ex2_approx(-hs * hv[j])) * __uint_as_float(v[j]);
                            if (feeds_skip) {
                                const int f = col0 + j;
                                if (f >= nsplit) {
                                    int coord;
                                    const float jac = embed_jac(x, f - nsplit, net.mx, coord);
                                    t = jac * (coord == 0 ? gb[0] : (coord == 1 ? gb[1] : gb[2]));
                                }
                                t *= RS2;
                            }
                            hv[j] = t;
                        }
                        const bool more = (l < NL - 1);
                        if (more) {
                            store_a16<false>(A_hi, A_lo, row, kc0, hv);
                            publish_chunk(&a_ready[c], lane);
                        }
                        store_a16<false>(A_hi, A_lo, row, kc0, hv, SL.wbase + SL.HD(l) + sg, (uint32_t)planes::BIG_PLANE, true, false);
                        if (!more) {
                            // the tangent pass is over and its last MMAs are done: build the A operand of the reverse pass
                            if (color) {
                                // pc_{Lc-2} = [c_{Lc-2} > 0] * (W_head^T delta)
                                const uint8_t* cs = SL.base + SL.C(net.Lc - 2) + sg;
                                const uint4 c0 = *reinterpret_cast<const uint4*>(cs), c1 = *reinterpret_cast<const uint4*>(cs + planes::SUB_CHUNK);
                                const uint32_t cw[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
                                const float* __restrict__ wh = net.col_head + col0;
#pragma unroll
                                for (int j = 0; j < 16; ++j) {
                                    const float cv = (j & 1) ? __uint_as_float(cw[j >> 1] & 0xffff0000u) : __uint_as_float(cw[j >> 1] << 16);
                                    const float u = fmaf(delta[0], __ldg(wh + j), fmaf(delta[1], __ldg(wh + 256 + j), delta[2] * __ldg(wh + 512 + j)));
                                    hv[j] = cv > 0.f ? u : 0.f;
                                }
                                store_a16<false>(A_hi, A_lo, row, kc0, hv);
                                publish_chunk(&a_ready[c], lane);
                                store_a16<false>(A_hi, A_lo, row, kc0, hv, SL.wbase + SL.PC(net.Lc - 2) + sg, (uint32_t)planes::BIG_PLANE, valid, false);
                            } else {
#pragma unroll
                                for (int j = 0; j < 16; ++j) hv[j] = 0.f;     // no radiance stack: fbar = 0
                                store_a16<false>(A_hi, A_lo, row, kc0, hv);
                                publish_chunk(&a_ready[c], lane);
                            }
                        }
                    } else if (kind == BK_COL_REV) {
                        // accumulator = W_l^T pc_l ; pc_{l-1} = that * [c_{l-1} > 0]
                        const uint8_t* cs = SL.base + SL.C(l - 1) + sg;
                        const uint4 c0 = *reinterpret_cast<const uint4*>(cs), c1 = *reinterpret_cast<const uint4*>(cs + planes::SUB_CHUNK);
                        const uint32_t cw[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
                        uint32_t v[16];
                        tmem_ld16(tmem_base + lane_base + b * 256u + (uint32_t)col0, v);
                        tmem_ld_wait();
                        float hv[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const float cv = (j & 1) ? __uint_as_float(cw[j >> 1] & 0xffff0000u) : __uint_as_float(cw[j >> 1] << 16);
                            hv[j] = cv > 0.f ? __uint_as_float(v[j]) : 0.f;
                        }
                        store_a16<false>(A_hi, A_lo, row, kc0, hv);
                        publish_chunk(&a_ready[c], lane);
                        store_a16<false>(A_hi, A_lo, row, kc0, hv, SL.wbase + SL.PC(l - 1) + sg, (uint32_t)planes::BIG_PLANE, valid, false);
                    } else if (kind == BK_FEAT_ADJ) {
                        // accumulator = adjoint of the features
                        uint32_t v[16];
                        tmem_ld16(tmem_base + lane_base + b * 256u + (uint32_t)col0, v);
                        tmem_ld_wait();
                        float hv[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) hv[j] = __uint_as_float(v[j]);
                        store_a16<false>(A_hi, A_lo, row, kc0, hv);
                        publish_chunk(&a_ready[c], lane);
                        store_a16<false>(A_hi, A_lo, row, kc0, hv, SL.wbase + SL.FB() + sg, (uint32_t)planes::BIG_PLANE, valid, false);
                    } else {
                        // BK_P: accumulator = W_{l+1}^T p_{l+1} ; produce p_l (l = op.layer), 8 columns at a time
                        const bool feeds_skip = (l + 1 == net.skip);
                        const bool top = (l == NL - 1);
                        const float hs = feeds_skip ? 144.26950408889634f * S2 : 144.26950408889634f;
                        float pv16[16];
#pragma unroll
                        for (int s = 0; s < 2; ++s) {
                            const size_t sg8 = sg + (size_t)s * planes::SUB_CHUNK;
                            const uint8_t* ph = SL.base + SL.H(l) + sg8;
                            const uint8_t* pq = SL.base + SL.Q(l) + sg8;
                            const uint8_t* pd = SL.wbase + SL.HD(l) + sg8;
                            const uint4 h_hi = *reinterpret_cast<const uint4*>(ph), h_lo = *reinterpret_cast<const uint4*>(ph + planes::BIG_PLANE);
                            const uint4 q_hi = *reinterpret_cast<const uint4*>(pq), q_lo = *reinterpret_cast<const uint4*>(pq + planes::BIG_PLANE);
                            const uint4 d_hi = *reinterpret_cast<const uint4*>(pd), d_lo = *reinterpret_cast<const uint4*>(pd + planes::BIG_PLANE);
                            uint32_t v[8];
                            tmem_ld8(tmem_base + lane_base + b * 256u + (uint32_t)(col0 + 8 * s), v);
                            tmem_ld_wait();
                            float hh[8], qq[8], dd[8], pv[8];
                            seg8_values(h_hi, h_lo, hh);
                            seg8_values(q_hi, q_lo, qq);
                            seg8_values(d_hi, d_lo, dd);
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const int f = col0 + 8 * s + j;
                                const float e = ex2_approx(-hs * hh[j]);          // 1 - softplus'
                                const float sp = 1.0f - e;
                                float u = __uint_as_float(v[j]);
                                if (top) u = fmaf(sbar, __ldg(net.sdf_head + f), u);
                                float hd = dd[j];
                                if (feeds_skip) { u *= RS2; hd *= S2; }
                                const float t2 = sp > 0.f ? 100.f * e * hd * qq[j] * rcp_approx(sp) : 0.f;
                                float p = fmaf(sp, u, t2);
                                if (feeds_skip && f >= nsplit) p = 0.f;
                                pv[j] = p;
                            }
                            if (!last_op) store_a8<false>(A_hi, A_lo, row, kc0 + s, pv, nullptr, true, true);
#pragma unroll
                            for (int j = 0; j < 8; ++j) pv16[s * 8 + j] = pv[j];
                        }
                        if (!last_op) publish_chunk(&a_ready[c], lane);
                        store_a16<false>(A_hi, A_lo, row, kc0, pv16, SL.wbase + SL.P(l) + sg, (uint32_t)planes::BIG_PLANE, valid, false);
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) { x[c] = xn[c]; gb[c] = gbn[c]; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<512>(tmem_base);
}

}  // namespace tcb

int tc_bwd_launch(const i2sdf_handle* h, const BwdParams& p, cudaStream_t st) {
    using namespace tcb;
    if (p.M <= 0) return I2SDF_OK;
    const chain::OpTable* tab = tc_bwd_table(h, p.with_color != 0);
    if (!tab || tab->nops == 0) { set_error("tc_bwd_launch: no backward op table for this network"); return I2SDF_E_INVALID; }
    static bool attr_done = false;
    static int pf = 1;
    if (!attr_done) {
        I2SDF_CUDA_CHECK(cudaFuncSetAttribute(tc_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)chain::kSmemBytes));
        I2SDF_CUDA_CHECK(cudaFuncSetAttribute(tc_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)chain::kSmemBytes));
        const char* e = getenv("I2SDF_BWD_PREFETCH");          // items ahead (1..4), 0 = off
        pf = e ? atoi(e) : 1;
        if (pf < 0 || pf > 4) pf = 1;
        attr_done = true;
    }
    const long long ntiles = (p.M + chain::TM - 1) / chain::TM;
    const int grid = (int)(ntiles < (long long)h->num_sms ? ntiles : (long long)h->num_sms);
    BwdParams q = p;
    q.pf_dist = pf;
    if (pf) tc_bwd_kernel<true><<<grid, chain::NTHREADS, chain::kSmemBytes, st>>>(q, *tab);
    else tc_bwd_kernel<false><<<grid, chain::NTHREADS, chain::kSmemBytes, st>>>(q, *tab);
    I2SDF_CUDA_CHECK(cudaGetLastError());
    return I2SDF_OK;
}

}  // namespace i2sdf
