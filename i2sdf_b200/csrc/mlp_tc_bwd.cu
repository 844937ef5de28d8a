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
#include <type_traits>
#include "common.cuh"
#include "tc_common.cuh"
#include "tc_chain.cuh"
#include "tc_bwd.cuh"

namespace i2sdf {
namespace tcb {

using namespace tc;
using namespace chain;

// Work items are 32 rows x 8 columns (round 2): in every iteration all 16 epilogue warps work on ONE 32-column chunk, so the next op's
// MMA chain starts after 1/8 of an epilogue; the slot segments an op reads (H / Q / HD / C: written up to 20 ops - or a whole kernel -
// earlier, i.e. in HBM) are pulled into L2 ONE ITEM ahead (the warp's next item of the same op, or the first item of the next op) and
// requested in front of the item's TMEM load; the slot copies of an item leave behind its publish.
// Prefetch distance, measured on the 1024-ray C2 step (profiles/r02y_bwd_prefetch_sweep.txt, I2SDF_BWD_PREFETCH): off 1.55 ms, 1 item
// 1.355, 2 items 1.37, 3 items 1.41, 4 items 1.43, one whole OP ahead 1.61 ms - an op ahead puts 148 CTAs x 384 KB = 57 MB of P-op
// operands (plus the dirty adjoint lines) into the 2 x 63 MB L2 at once and evicts them again before use.
// Shared memory: A_hi | A_lo (32 chunks each: K <= 256) | weight ring | heads (sdf head 257, rgb head 771 floats) | barriers.
constexpr int B8_A_PART = 32 * TM * 16;
constexpr int B8_PARAM_FLOATS = 260 + 772;
constexpr size_t kSmemBwd8 = 128 + 2 * (size_t)B8_A_PART + NSTAGE * STAGE_MAX + B8_PARAM_FLOATS * 4 + 256;
constexpr int kBwdPrefetchDefault = 1;
enum { KB_TAN = 0, KB_TAN_SKIP, KB_TAN_LAST, KB_COL_REV, KB_FEAT_ADJ, KB_P, KB_P_SKIP, KB_P_TOP };

// TL: development probe (I2SDF_DEBUG_TIMELINE=1, tools/timeline.py bwd): clock64 stamps of CTA 0's second tile into P.tl - per op and epilogue
// warp [wait starts, accumulator seen, first item published, last item done] at tl[(op * 16 + warp) * 4], the MMA warp's stamps as in
// chain_mma, and for epilogue warp 0 every item's [start, TMEM + operands there, values computed, published, slot stores issued] at
// tl[7168 + (op * 8 + it) * 5]
__device__ __forceinline__ long long tl_clock() { long long t; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t)::"memory"); return t; }
template <bool TL = false>
__global__ void __launch_bounds__(NTHREADS, 1) tc_bwd8_kernel(const BwdParams P, const OpTable T) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_align128(smem_raw);
    uint8_t* A_hi = smem;
    uint8_t* A_lo = smem + B8_A_PART;
    uint8_t* ring = smem + 2 * B8_A_PART;
    float* s_sdf_head = reinterpret_cast<float*>(ring + NSTAGE * STAGE_MAX);
    float* s_col_head = s_sdf_head + 260;
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_sdf_head + B8_PARAM_FLOATS);
    uint64_t* full = bars;
    uint64_t* empty = bars + NSTAGE;
    uint64_t* a_ready = bars + 2 * NSTAGE;       // [N_READY], 16 arrivals each
    uint64_t* d_full = a_ready + N_READY;        // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d_full + 2);

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
        chain_mma<TL>(T, ntiles, tmem_base, A_hi, A_lo, ring, full, empty, a_ready, d_full, P.tl);
    } else {
        // ================= epilogue warps =================
        const int q = warp & 3;
        const int sub = (warp - 2) >> 2;                     // column group: columns 8 sub .. 8 sub + 7 of every 32-column chunk
        const int row = q * 32 + lane;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        const int nsplit = 256 - net.ex;
        const float RS2 = 0.70710678118654752f, S2 = 1.41421356237309505f, C1 = 144.26950408889634f;
        const planes::Layout& SL = P.sl;
        const bool color = P.with_color != 0;
        uint32_t dphase = 0, g = 0;
        float x[3], gb[3];
        // point of this thread's row and the upstream of grad_x sdf there
        auto load_point = [&](long long tile) {
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
                if (m < P.m_up) { if (P.g_grad) { gb[0] = P.g_grad[m * 3]; gb[1] = P.g_grad[m * 3 + 1]; gb[2] = P.g_grad[m * 3 + 2]; } }
                else if (P.g_grad_tail) { const float* gt = P.g_grad_tail + (m - P.m_up) * 3; gb[0] = gt[0]; gb[1] = gt[1]; gb[2] = gt[2]; }
            }
        };
        auto gb_of = [&](int coord) -> float { return coord == 0 ? gb[0] : (coord == 1 ? gb[1] : gb[2]); };
        // prologue: A_0 = tangent of the embedding, J(x) gbar, 48 columns = k chunks 0..5 (group sub: chunk sub, groups 0 / 1 also 4 / 5), also slot ED
        auto prologue = [&](long long tile) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int kc = sub + 4 * h;
                if (kc < 6) {
                    float hv[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int i = kc * 8 + j;
                        int coord = 0;
                        const float jac = (i < net.ex) ? embed_jac(x, i, net.mx, coord) : 0.f;
                        hv[j] = jac * gb_of(coord);
                    }
                    uint32_t hh[4], ll[4];
                    sts_a8<false>(A_hi, A_lo, row, kc, hv, hh, ll);
                    stg_a8<true>(SL.wbase + SL.ED() + planes::seg(tile * TM + row, kc, planes::SMALL_CHUNKS), (uint32_t)planes::SMALL_PLANE, true, hv, hh, ll);
                }
            }
            fence_proxy_async();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { mbar_arrive(&a_ready[0]); mbar_arrive(&a_ready[1]); }
        };
        // L2 prefetch of the 8 segments (one per item) this thread reads from a slot in one op
        auto pf_slot = [&](const uint8_t* base, long long m, bool both_planes, int slot_planes = 2) {
            const uint8_t* s0 = base + planes::segp(m, sub, planes::BIG_CHUNKS, slot_planes);
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                pf_l2(s0 + (size_t)it * 4 * planes::SUB_CHUNK);
                if (both_planes) pf_l2(s0 + (size_t)it * 4 * planes::SUB_CHUNK + planes::BIG_PLANE);
            }
        };
        // the same for ONE item (chunk 4 itx + sub) of an op of kind bk (BK_*), layer pl: the item-ahead mode (P.pf_dist = 1..7 items)
        auto pf_item = [&](int bk, int pl, long long pm, int itx) {
            const int pkc = itx * 4 + sub;
            const size_t psg = planes::seg(pm, pkc, planes::BIG_CHUNKS);
            if (bk == BK_TAN || bk == BK_P) {
                const uint8_t* ph = SL.base + SL.H(pl) + psg;
                pf_l2(ph); pf_l2(ph + planes::BIG_PLANE);
                if (bk == BK_P) {
                    const uint8_t* pq = SL.base + SL.Q(pl) + psg;
                    const uint8_t* pd = SL.wbase + SL.HD(pl) + planes::segp(pm, pkc, planes::BIG_CHUNKS, planes::kPlanesHD);
                    pf_l2(pq); pf_l2(pq + planes::BIG_PLANE);
                    pf_l2(pd);
                    if (planes::kPlanesHD == 2) pf_l2(pd + planes::BIG_PLANE);
                } else if (pl == NL - 1 && color) pf_l2(SL.base + SL.C(net.Lc - 2) + psg);
            } else if (bk == BK_COL_REV) pf_l2(SL.base + SL.C(pl - 1) + psg);
        };
        const int pfd = P.pf_dist;
        if ((long long)blockIdx.x < ntiles) { load_point(blockIdx.x); prologue(blockIdx.x); }
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const long long m = tile * TM + row;
            const bool valid = m < P.M;
            const long long next_tile = tile + gridDim.x;
            // per-point upstream scalars
            const bool has_up = valid && m < P.m_up;
            const float sbar = (has_up && P.g_sdf) ? P.g_sdf[m] : 0.f;
            float delta[3] = {0.f, 0.f, 0.f};
            if (color && has_up) {
#pragma unroll
                for (int c = 0; c < 3; ++c) { const float y = P.s_rgb[m * 3 + c]; delta[c] = P.g_rgb[m * 3 + c] * y * (1.f - y); }
            }
            if (tile == (long long)blockIdx.x && pfd >= 8) pf_slot(SL.base + SL.H(0), m, true);       // first op of the first tile
            for (int op = 0; op < T.nops; ++op, ++g) {
                const uint32_t b = g & 1u;
                const int kind = T.ops[op].kind, l = T.ops[op].layer;
                const bool last_op = (op == T.nops - 1);
                long long* tlw = (TL && P.tl && lane == 0 && blockIdx.x == 0 && tile == (long long)gridDim.x) ? P.tl + (op * 16 + (warp - 2)) * 4 : nullptr;
                long long* tli = (TL && tlw && warp == 2) ? P.tl + 7168 + op * 40 : nullptr;
                const uint32_t acc_addr = tmem_base + lane_base + b * 256u;
                // what the NEXT op reads (for the last op: the next tile's first op)
                const int nk = last_op ? (next_tile < ntiles ? (int)BK_TAN : -1) : T.ops[op + 1].kind, nl = last_op ? 0 : T.ops[op + 1].layer;
                const long long nm = last_op ? next_tile * TM + row : m;
                // Slot operands travel ONE ITEM AHEAD in registers (round 2, profiles/r02e_timeline_bwd.txt: with the loads issued at the top of
                // their own item every item exposed 400-1300 clocks of L2 / HBM latency, all four warps of a scheduler waiting together): item 0's
                // segments are requested BEFORE the wait for the accumulator (the warp idles there anyway), item it + 1's as soon as item it's
                // values are computed - the operand registers are dead from there on - i.e. in front of the publish fence and the slot stores,
                // which cover most of the latency.  Same register peak as before.
                uint4 h_hi = make_uint4(0, 0, 0, 0), h_lo = h_hi, c_hi = h_hi;
                auto request = [&](auto kind_c, int it2) {
                    constexpr int K = decltype(kind_c)::value;
                    constexpr bool is_tan = (K == KB_TAN || K == KB_TAN_SKIP || K == KB_TAN_LAST);
                    constexpr bool is_p = (K == KB_P || K == KB_P_SKIP || K == KB_P_TOP);
                    if (is_p) return;                          // P ops load inside their own item, into item-local registers (see items)
                    const int kc2 = it2 * 4 + sub;
                    const size_t sg2 = planes::seg(m, kc2, planes::BIG_CHUNKS);
                    if (is_tan) {
                        const uint8_t* ph = SL.base + SL.H(l) + sg2;
                        h_hi = ldg_cs(ph); h_lo = ldg_cs(ph + planes::BIG_PLANE);
                    }
                    if (K == KB_COL_REV) c_hi = ldg_cs(SL.base + SL.C(l - 1) + sg2);
                    if (K == KB_TAN_LAST && color) c_hi = ldg_cs(SL.base + SL.C(net.Lc - 2) + sg2);
                };
                // kind of this op as a compile-time constant for the two templated pieces (request / items)
                auto dispatch = [&](auto&& fn) {
                    switch (kind) {
                        case BK_TAN:
                            if (l == NL - 1) fn(std::integral_constant<int, KB_TAN_LAST>{});
                            else if (l + 1 == net.skip) fn(std::integral_constant<int, KB_TAN_SKIP>{});
                            else fn(std::integral_constant<int, KB_TAN>{});
                            break;
                        case BK_COL_REV: fn(std::integral_constant<int, KB_COL_REV>{}); break;
                        case BK_FEAT_ADJ: fn(std::integral_constant<int, KB_FEAT_ADJ>{}); break;
                        default:
                            if (l == NL - 1) fn(std::integral_constant<int, KB_P_TOP>{});
                            else if (l + 1 == net.skip) fn(std::integral_constant<int, KB_P_SKIP>{});
                            else fn(std::integral_constant<int, KB_P>{});
                            break;
                    }
                };
                dispatch([&](auto kind_c) { request(kind_c, 0); });        // (does nothing for P ops)
                // accumulator of this op complete; then the work that had to wait for it
                if (TL && tlw) tlw[0] = tl_clock();
                mbar_wait(&d_full[b], (dphase >> b) & 1u);
                dphase ^= (1u << b);
                tc_fence_after();
                if (TL && tlw) tlw[1] = tl_clock();
                if (pfd >= 8) {   // op-ahead mode: pull all of it into L2 now
                    if (nk >= 0) {
                        if (nk == BK_TAN) {
                            pf_slot(SL.base + SL.H(nl), nm, true);
                            if (nl == NL - 1 && color) pf_slot(SL.base + SL.C(net.Lc - 2), nm, false);
                        } else if (nk == BK_COL_REV) pf_slot(SL.base + SL.C(nl - 1), nm, false);
                        else if (nk == BK_P) { pf_slot(SL.base + SL.H(nl), nm, true); pf_slot(SL.base + SL.Q(nl), nm, true); pf_slot(SL.wbase + SL.HD(nl), nm, planes::kPlanesHD == 2, planes::kPlanesHD); }
                    }
                }
                if (last_op && next_tile < ntiles) {
                    // the A operand is free (this tile's last MMAs are done): start the next tile's first op now
                    load_point(next_tile);
                    prologue(next_tile);
                }
                auto items = [&](auto kind_c) {
                    constexpr int K = decltype(kind_c)::value;
                    constexpr bool is_tan = (K == KB_TAN || K == KB_TAN_SKIP || K == KB_TAN_LAST);
                    constexpr bool is_p = (K == KB_P || K == KB_P_SKIP || K == KB_P_TOP);
                    const float hs = (K == KB_TAN_SKIP || K == KB_P_SKIP) ? C1 * S2 : C1;
#pragma unroll 1
                    for (int it = 0; it < 8; ++it) {
                        const int col0 = it * 32 + sub * 8, kc = it * 4 + sub;
                        const size_t sg_adj = planes::segp(m, kc, planes::BIG_CHUNKS, planes::kPlanesAdj);      // adjoint slots: HI plane only
                        const size_t sg_hd = planes::segp(m, kc, planes::BIG_CHUNKS, planes::kPlanesHD);
                        if (TL && tli) tli[it * 5] = tl_clock();
                        if (pfd > 0 && pfd < 8) {
                            // L2 prefetch: the item that will be REQUESTED pfd items from now (this warp's item it + 1 + pfd of the same op, or the head of
                            // the next op)
                            const int tt = it + pfd + (is_p ? 0 : 1);
                            if (tt < 8) pf_item(is_p ? (int)BK_P : (is_tan ? (int)BK_TAN : (K == KB_COL_REV ? (int)BK_COL_REV : (int)BK_FEAT_ADJ)), l, m, tt);
                            else if (nk >= 0 && tt - 8 < 8) pf_item(nk, nl, nm, tt - 8);
                        }
                        // P ops: the six segments of THIS item, into item-local registers, in front of the item's own TMEM load (see below)
                        uint4 p_h_hi = make_uint4(0, 0, 0, 0), p_h_lo = p_h_hi, q_hi = p_h_hi, q_lo = p_h_hi, d_hi = p_h_hi, d_lo = p_h_hi;
                        if (is_p) {
                            const size_t sg = planes::seg(m, kc, planes::BIG_CHUNKS);
                            const uint8_t* ph = SL.base + SL.H(l) + sg;
                            const uint8_t* pq = SL.base + SL.Q(l) + sg;
                            const uint8_t* pd = SL.wbase + SL.HD(l) + sg_hd;
                            p_h_hi = ldg_cs(ph); p_h_lo = ldg_cs(ph + planes::BIG_PLANE);
                            q_hi = ldg_cs(pq); q_lo = ldg_cs(pq + planes::BIG_PLANE);
                            d_hi = ldg_cs(pd);
                            if (planes::kPlanesHD == 2) d_lo = ldg_cs(pd + planes::BIG_PLANE);
                        }
                        uint32_t v[8];
                        tmem_ld8(acc_addr + (uint32_t)col0, v);
                        tmem_ld_wait();
                        if (TL && tli) {
                            // operands there: a dependent use of every requested segment in front of the stamp
                            const uint32_t dep = h_hi.x ^ h_lo.x ^ p_h_hi.x ^ p_h_lo.x ^ q_hi.x ^ q_lo.x ^ d_hi.x ^ d_lo.x ^ c_hi.x;
                            asm volatile("" ::"r"(dep) : "memory");
                            tli[it * 5 + 1] = tl_clock();
                        }
                        float hv[8];
                        uint8_t* gs = nullptr;
                        bool keep = true, to_smem = true, adj_slot = true;       // adj_slot: gs is an adjoint slot (HI plane only), else a tangent slot
                        if (is_tan) {
                            // hdot_l = softplus'(a_l) * adot_l   (skip concat: [hdot | J gbar] / sqrt2)
                            float hh[8];
                            seg8_values<false>(h_hi, h_lo, hh);
#pragma unroll
                            for (int j = 0; j < 8; ++j) hv[j] = (1.0f - ex2_approx(-hs * hh[j])) * __uint_as_float(v[j]);
                            if (K == KB_TAN_SKIP) {
                                if (col0 + 8 > nsplit) {
#pragma unroll
                                    for (int j = 0; j < 8; ++j) {
                                        const int f = col0 + j;
                                        if (f >= nsplit) {
                                            int coord;
                                            const float jac = embed_jac(x, f - nsplit, net.mx, coord);
                                            hv[j] = jac * gb_of(coord);
                                        }
                                    }
                                }
#pragma unroll
                                for (int j = 0; j < 8; ++j) hv[j] *= RS2;
                            }
                            gs = SL.wbase + SL.HD(l) + sg_hd;
                            adj_slot = false;
                            if (K == KB_TAN_LAST) {
                                // the tangent pass is over and its last MMAs are done: HD(l) to its slot, then build the A operand of the reverse pass
                                uint32_t th[4], tl[4];
#pragma unroll
                                for (int i = 0; i < 4; ++i) split_bf16x2(hv[2 * i], hv[2 * i + 1], th[i], tl[i]);
                                stg_cs(gs, make_uint4(th[0], th[1], th[2], th[3]));
                                if (planes::kPlanesHD == 2) stg_cs(gs + planes::BIG_PLANE, make_uint4(tl[0], tl[1], tl[2], tl[3]));
                                gs = nullptr;
                                adj_slot = true;
                                if (color) {
                                    // pc_{Lc-2} = [c_{Lc-2} > 0] * (W_head^T delta)
                                    const uint32_t cw[4] = {c_hi.x, c_hi.y, c_hi.z, c_hi.w};
                                    const float* __restrict__ wh = s_col_head + col0;
#pragma unroll
                                    for (int j = 0; j < 8; ++j) {
                                        const float cv = (j & 1) ? __uint_as_float(cw[j >> 1] & 0xffff0000u) : __uint_as_float(cw[j >> 1] << 16);
                                        const float u = fmaf(delta[0], wh[j], fmaf(delta[1], wh[256 + j], delta[2] * wh[512 + j]));
                                        hv[j] = cv > 0.f ? u : 0.f;
                                    }
                                    gs = SL.wbase + SL.PC(net.Lc - 2) + sg_adj;
                                    keep = valid;
                                } else {
#pragma unroll
                                    for (int j = 0; j < 8; ++j) hv[j] = 0.f;     // no radiance stack: fbar = 0
                                }
                            }
                        } else if (K == KB_COL_REV) {
                            // accumulator = W_l^T pc_l ; pc_{l-1} = that * [c_{l-1} > 0]
                            const uint32_t cw[4] = {c_hi.x, c_hi.y, c_hi.z, c_hi.w};
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const float cv = (j & 1) ? __uint_as_float(cw[j >> 1] & 0xffff0000u) : __uint_as_float(cw[j >> 1] << 16);
                                hv[j] = cv > 0.f ? __uint_as_float(v[j]) : 0.f;
                            }
                            gs = SL.wbase + SL.PC(l - 1) + sg_adj;
                            keep = valid;
                        } else if (K == KB_FEAT_ADJ) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) hv[j] = __uint_as_float(v[j]);
                            gs = SL.wbase + SL.FB() + sg_adj;
                            keep = valid;
                        } else {
                            // BK_P: accumulator = W_{l+1}^T p_{l+1} ; p_l = s'(a_l) u + s''(a_l) adot_l v_l with q_l = s'(a_l) v_l from the forward
                            float hh[8], qq[8], dd[8];
                            seg8_values<false>(p_h_hi, p_h_lo, hh);
                            seg8_values<false>(q_hi, q_lo, qq);
                            seg8_values<false>(d_hi, d_lo, dd);
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const float e = ex2_approx(-hs * hh[j]);          // 1 - softplus'
                                const float sp = 1.0f - e;
                                float u = __uint_as_float(v[j]);
                                if (K == KB_P_TOP) u = fmaf(sbar, s_sdf_head[col0 + j], u);
                                float hd = dd[j];
                                if (K == KB_P_SKIP) { u *= RS2; hd *= S2; }
                                const float t2 = sp > 0.f ? 100.f * e * hd * qq[j] * rcp_approx(sp) : 0.f;
                                float p = fmaf(sp, u, t2);
                                if (K == KB_P_SKIP && col0 + j >= nsplit) p = 0.f;
                                hv[j] = p;
                            }
                            gs = SL.wbase + SL.P(l) + sg_adj;
                            keep = valid;
                            to_smem = !last_op;
                        }
                        uint32_t hh2[4], ll2[4];
                        if (TL && tli) { asm volatile("" ::"f"(hv[0]), "f"(hv[7]) : "memory"); tli[it * 5 + 2] = tl_clock(); }
                        // the operand registers are free: T / CR ops (1-2 segments per item) request the next item's segments here, in front of the
                        // publish: its fence (MEMBAR.ALL.CTA) also waits for loads in flight, but two L2-resident segments cost it little and their
                        // latency hides behind the publish + store phase (T ops 17.2 k -> 15.5 k clocks per op, CR ops 14.6 k -> 14.1 k).  P ops (6
                        // segments = 48 KB per 32-column step per CTA) run at their share of HBM bandwidth: requested one item ahead - in front of the
                        // publish or behind it - the segments arrive no earlier and only delay the publish and the slot stores (measured: 23 k ->
                        // 26.5 k / 27.5 k clocks per op, profiles/r02f_ / r02g_timeline_bwd.txt), so P ops request at the top of their own item
                        if (!is_p && it < 7) request(kind_c, it + 1);
                        if (to_smem) {
                            sts_a8<false>(A_hi, A_lo, row, kc, hv, hh2, ll2);
                            publish_chunk(&a_ready[it], lane);
                            if (TL && tlw && it == 0) tlw[2] = tl_clock();
                        } else {
#pragma unroll
                            for (int i = 0; i < 4; ++i) split_bf16x2(hv[2 * i], hv[2 * i + 1], hh2[i], ll2[i]);
                        }
                        if (TL && tli) tli[it * 5 + 3] = tl_clock();
                        if (gs) {
                            if (adj_slot) stg_a8<true, planes::kPlanesAdj>(gs, (uint32_t)planes::BIG_PLANE, keep, hv, hh2, ll2);
                            else stg_a8<true, planes::kPlanesHD>(gs, (uint32_t)planes::BIG_PLANE, keep, hv, hh2, ll2);
                        }
                        if (TL && tli) tli[it * 5 + 4] = tl_clock();
                    }
                    if (TL && tlw) tlw[3] = tl_clock();
                };
                dispatch(items);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<512>(tmem_base);
}

}  // namespace tcb

static long long* g_bwd_tl = nullptr;       // development probe buffer (I2SDF_DEBUG_TIMELINE)

// copies the probe's stamps to the host (after a synchronize); returns the number of int64 copied, 0 if the probe never ran
long long tc_bwd_timeline_read(long long* out, long long n) {
    if (!g_bwd_tl || !out || n <= 0) return 0;
    if (n > 8192) n = 8192;
    if (cudaMemcpy(out, g_bwd_tl, (size_t)n * 8, cudaMemcpyDeviceToHost) != cudaSuccess) return 0;
    return n;
}

int tc_bwd_launch(const i2sdf_handle* h, const BwdParams& p, cudaStream_t st) {
    using namespace tcb;
    if (p.M <= 0) return I2SDF_OK;
    const chain::OpTable* tab = tc_bwd_table(h, p.with_color != 0);
    if (!tab || tab->nops == 0) { set_error("tc_bwd_launch: no backward op table for this network"); return I2SDF_E_INVALID; }
    // (per device: cudaFuncSetAttribute is a per-device setting)
    static PerDeviceOnce once;
    if (once.need()) {
        I2SDF_CUDA_CHECK(cudaFuncSetAttribute(tc_bwd8_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBwd8));
        I2SDF_CUDA_CHECK(cudaFuncSetAttribute(tc_bwd8_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBwd8));
    }
    const long long ntiles = (p.M + chain::TM - 1) / chain::TM;
    const int grid = balanced_grid(h->num_sms, ntiles);
    // L2 prefetch distance of the slot segments: 1..7 = that many 8-column items ahead, 8 = one whole op ahead, 0 = off
    static const int pf = [] { const char* e = getenv("I2SDF_BWD_PREFETCH"); const int v = e ? atoi(e) : kBwdPrefetchDefault; return (v < 0 || v > 8) ? kBwdPrefetchDefault : v; }();
    BwdParams q = p;
    q.pf_dist = pf;
    static const bool tl_on = getenv("I2SDF_DEBUG_TIMELINE") != nullptr;
    if (tl_on) {
        if (!g_bwd_tl) { I2SDF_CUDA_CHECK(cudaMalloc(&g_bwd_tl, 65536)); I2SDF_CUDA_CHECK(cudaMemset(g_bwd_tl, 0, 65536)); }
        q.tl = g_bwd_tl;
        tc_bwd8_kernel<true><<<grid, chain::NTHREADS, kSmemBwd8, st>>>(q, *tab);
    } else {
        q.tl = nullptr;
        tc_bwd8_kernel<false><<<grid, chain::NTHREADS, kSmemBwd8, st>>>(q, *tab);
    }
    I2SDF_CUDA_CHECK(cudaGetLastError());
    return I2SDF_OK;
}

}  // namespace i2sdf
