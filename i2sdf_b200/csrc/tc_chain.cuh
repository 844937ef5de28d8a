// Shared skeleton of the tensor-core "chain" kernels (forward mlp_tc3.cu, backward mlp_tc_bwd.cu): a resident tile of 128
// points walks through a table of dense ops; warp 0 streams the packed weight blocks of each op through a 4-stage SMEM
// ring, warp 1 issues 3 tcgen05.mma per k step (A_hi W_hi + A_lo W_hi + A_hi W_lo; halves = fp16 in the forward chains, bf16 in
// the backward chain) into one of two TMEM accumulators, 16
// epilogue warps turn accumulator g into the A operand of op g+1 in 16-column work items and publish 32-column chunks.
#pragma once
#include "common.cuh"
#include "tc_common.cuh"

namespace i2sdf {
namespace chain {

using namespace tc;

constexpr int TM = 128;
constexpr int NSTAGE = 4;
constexpr int STAGE_MAX = 16384;
constexpr int A_CHUNKS = 36;                         // 288 columns
constexpr int A_PART_BYTES = A_CHUNKS * TM * 16;     // 73728
constexpr int N_EPI_WARPS = 16;
constexpr int NTHREADS = (2 + N_EPI_WARPS) * 32;
constexpr int MAX_OPS = 40;
constexpr int N_READY = 9;
constexpr uint32_t LBO_A = TM * 16, SBO = 128;
constexpr int PART_FLOATS = 4 * 7 * TM;              // [sub][sdf, rgb x3, grad x3][row]
constexpr size_t kSmemBytes = 1024 + 2 * (size_t)A_PART_BYTES + NSTAGE * STAGE_MAX + PART_FLOATS * 4 + 256;

enum { EK_SDF_HIDDEN = 0, EK_SDF_LAST, EK_FEAT, EK_COL_HIDDEN, EK_COL_LAST, EK_REV, EK_GRAD, EK_SDF_LAST_REV };

struct Op {
    int w_off;          // byte offset into wpack
    short ksteps;
    short n;            // MMA N (256 or 48)
    short kind;         // epilogue kind applied to this op's accumulator
    short layer;        // layer index within its stack
};
struct OpTable {
    const uint8_t* wpack;
    int nops;
    int f16;            // 1: blocks and A operand are fp16 hi/lo, weights scaled by tc::kWScale (forward chains); 0: bf16 hi/lo (backward chain)
    float acc_scale;    // fp16 chains: what the epilogues multiply the accumulator by = kWInv * (1 + kappa), see tc3::acc_scale()
    Op ops[MAX_OPS];
};

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(N_EPI_WARPS * 32) : "memory"); }

template <bool F16>
__device__ __forceinline__ void split2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    if (F16) split_f16x2(x0, x1, hi, lo);
    else split_bf16x2(x0, x1, hi, lo);
}
// 16 consecutive columns (k chunks kc0, kc0+1) of one row of the next A operand, split hi / lo in the chain's format (F16S).
// g (optional): HI segment of chunk kc0 of this thread's point in a plane slot (planes.cuh), g_lo = byte distance to the LO
// plane: SLOT_SAME -> the same split values are stored there (the eval scratch of the forward, every slot of the bf16 backward);
// !SLOT_SAME -> the forward's fp16 chain saving a training slot: slots are bf16 hi/lo (what the backward chain and the weight
// gradients multiply - tcgen05 cannot mix the two formats in one MMA), so the values are split a second time.
// Zeros if !keep: adjoint slots of rows beyond M.
template <bool F16S, bool SLOT_SAME = true>
__device__ __forceinline__ void store_a16(uint8_t* A_hi, uint8_t* A_lo, int row, int kc0, const float (&hv)[16], uint8_t* g = nullptr,
                                          uint32_t g_lo = 0, bool keep = true, bool to_smem = true) {
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        uint32_t h[4], lo[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) split2<F16S>(hv[s * 8 + 2 * i], hv[s * 8 + 2 * i + 1], h[i], lo[i]);
        if (to_smem) {
            const uint32_t off = seg_off<TM>(row, kc0 + s);
            *reinterpret_cast<uint4*>(A_hi + off) = make_uint4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<uint4*>(A_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
        if (g) {
            if (!SLOT_SAME) {
#pragma unroll
                for (int i = 0; i < 4; ++i) split_bf16x2(hv[s * 8 + 2 * i], hv[s * 8 + 2 * i + 1], h[i], lo[i]);
            }
            uint8_t* gs = g + s * planes::SUB_CHUNK;
            *reinterpret_cast<uint4*>(gs) = keep ? make_uint4(h[0], h[1], h[2], h[3]) : make_uint4(0, 0, 0, 0);
            *reinterpret_cast<uint4*>(gs + g_lo) = keep ? make_uint4(lo[0], lo[1], lo[2], lo[3]) : make_uint4(0, 0, 0, 0);
        }
    }
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr)
                 : "memory");
}
// one chunk (8 columns) of a row of the next A operand and / or of a slot; g = HI segment, g_lo = byte distance to the LO plane
// (SLOT_SAME as in store_a16)
template <bool F16S, bool SLOT_SAME = true>
__device__ __forceinline__ void store_a8(uint8_t* A_hi, uint8_t* A_lo, int row, int kc, const float (&hv)[8], uint8_t* g, bool keep, bool to_smem,
                                         uint32_t g_lo = (uint32_t)planes::BIG_PLANE) {
    uint32_t h[4], lo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) split2<F16S>(hv[2 * i], hv[2 * i + 1], h[i], lo[i]);
    if (to_smem) {
        const uint32_t off = seg_off<TM>(row, kc);
        *reinterpret_cast<uint4*>(A_hi + off) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4*>(A_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
    if (g) {
        if (!SLOT_SAME) {
#pragma unroll
            for (int i = 0; i < 4; ++i) split_bf16x2(hv[2 * i], hv[2 * i + 1], h[i], lo[i]);
        }
        *reinterpret_cast<uint4*>(g) = keep ? make_uint4(h[0], h[1], h[2], h[3]) : make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4*>(g + g_lo) = keep ? make_uint4(lo[0], lo[1], lo[2], lo[3]) : make_uint4(0, 0, 0, 0);
    }
}
// the same in two steps, for epilogues that publish the SMEM operand BEFORE they store the slot copy (a fence.proxy.async behind a
// global store compiles to MEMBAR.ALL.CTA and waits for that store: measured on the main pass, F ops 10-12 k clocks per epilogue with
// the slot / scratch store in front of the fence against 6 k for ops without one)
template <bool F16S>
__device__ __forceinline__ void sts_a8(uint8_t* A_hi, uint8_t* A_lo, int row, int kc, const float (&hv)[8], uint32_t (&h)[4], uint32_t (&lo)[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) split2<F16S>(hv[2 * i], hv[2 * i + 1], h[i], lo[i]);
    const uint32_t off = seg_off<TM>(row, kc);
    *reinterpret_cast<uint4*>(A_hi + off) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(A_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}
template <bool SLOT_SAME, int NPLANES = 2>
__device__ __forceinline__ void stg_a8(uint8_t* g, uint32_t g_lo, bool keep, const float (&hv)[8], uint32_t (&h)[4], uint32_t (&lo)[4]) {
    if (!SLOT_SAME) {
#pragma unroll
        for (int i = 0; i < 4; ++i) split_bf16x2(hv[2 * i], hv[2 * i + 1], h[i], lo[i]);
    }
    // streaming stores: the eval scratch is read back once by the same CTA, the training slots once by the backward kernels, and
    // neither should push the weight blocks out of L2
    stg_cs(g, keep ? make_uint4(h[0], h[1], h[2], h[3]) : make_uint4(0, 0, 0, 0));
    if (NPLANES == 2) stg_cs(g + g_lo, keep ? make_uint4(lo[0], lo[1], lo[2], lo[3]) : make_uint4(0, 0, 0, 0));
}
__device__ __forceinline__ void pf_l2(const uint8_t* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// 8 columns (one chunk) of a slot: hi / lo segments -> fp32
template <bool F16 = false>
__device__ __forceinline__ void seg8_values(const uint4& hi, const uint4& lo, float (&out)[8]) {
    const uint32_t hw[4] = {hi.x, hi.y, hi.z, hi.w}, lw[4] = {lo.x, lo.y, lo.z, lo.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 v = F16 ? join_f16x2(hw[i], lw[i]) : join_bf16x2(hw[i], lw[i]);
        out[2 * i] = v.x;
        out[2 * i + 1] = v.y;
    }
}

// 16 columns of a 256-column slot back as fp32: seg = HI segment of the first chunk
__device__ __forceinline__ void load_slot16(const uint8_t* seg, uint4 (&raw)[4]) {
    raw[0] = *reinterpret_cast<const uint4*>(seg);
    raw[1] = *reinterpret_cast<const uint4*>(seg + planes::SUB_CHUNK);
    raw[2] = *reinterpret_cast<const uint4*>(seg + planes::BIG_PLANE);
    raw[3] = *reinterpret_cast<const uint4*>(seg + planes::BIG_PLANE + planes::SUB_CHUNK);
}
template <bool F16 = false>
__device__ __forceinline__ void slot16_values(const uint4 (&raw)[4], float (&out)[16]) {
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        const uint32_t hw[4] = {raw[s].x, raw[s].y, raw[s].z, raw[s].w}, lw[4] = {raw[2 + s].x, raw[2 + s].y, raw[2 + s].z, raw[2 + s].w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 v = F16 ? join_f16x2(hw[i], lw[i]) : join_bf16x2(hw[i], lw[i]);
            out[s * 8 + 2 * i] = v.x;
            out[s * 8 + 2 * i + 1] = v.y;
        }
    }
}
// softplus_100'(a) = 1 - exp(-100 softplus_100(a)), from a stored activation h = softplus_100(a)
__device__ __forceinline__ float dsoftplus_from_h(float h) { return 1.0f - ex2_approx(-144.26950408889634f * h); }
__device__ __forceinline__ void publish_chunk(uint64_t* bar, int lane) {
    fence_proxy_async();          // generic-proxy smem writes -> async proxy (tensor core operand reads)
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar);
}

// d embed_i / d x_c(i) and the coordinate c(i) it belongs to
__device__ __forceinline__ float embed_jac(const float (&x)[3], int i, int mx, int& coord) {
    if (i < 3) { coord = i; return 1.f; }
    const int qq = i - 3, k = qq / 6, cc = qq % 3;
    coord = cc;
    if (k >= mx) return 0.f;
    const float f = (float)(1 << k);
    float s, c;
    sincos_cw(__fmul_rn(x[cc], f), s, c);
    return ((qq % 6) < 3) ? f * c : -f * s;
}


// ---- warp 0, lane 0: weight producer --------------------------------------------------------------------------------
// Every op occupies a whole number of passes over the ring (k steps padded to a multiple of NSTAGE with "null" steps: the stage
// is signalled without a copy), so an op always starts at stage 0 and the MMA issuer's loop over the stages can be unrolled.
__device__ __forceinline__ void chain_producer(const OpTable& T, long long ntiles, uint8_t* ring, uint64_t* full, uint64_t* empty) {
    uint32_t phase = 0;
    const uint64_t pol = l2_policy_evict_last();
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        for (int op = 0; op < T.nops; ++op) {
            const uint32_t sb = (uint32_t)T.ops[op].n * 64u;
            const uint8_t* src = T.wpack + T.ops[op].w_off;
            const int nks = T.ops[op].ksteps;
            for (int ks0 = 0; ks0 < nks; ks0 += NSTAGE) {
#pragma unroll
                for (int st = 0; st < NSTAGE; ++st) {
                    const int ks = ks0 + st;
                    mbar_wait(&empty[st], phase ^ 1);
                    if (ks < nks) {
                        mbar_arrive_expect_tx(&full[st], sb);
                        bulk_g2s_hint(ring + st * STAGE_MAX, src + (size_t)ks * sb, sb, &full[st], pol);
                    } else {
                        mbar_arrive(&full[st]);
                    }
                }
                phase ^= 1;
            }
        }
    }
}

// ---- warp 1: MMA issuer (waits per 32-column chunk of the A operand, commits per op) ----------------------------------------
// Executed by ALL 32 lanes of the warp in converged control flow; one elected lane issues.  Under `if (lane == 0)` the
// compiler cannot prove the descriptors / TMEM addresses warp-uniform and wraps every UTCHMMA / UTCBAR in an
// ELECT + R2UR.BROADCAST + BRA.U.ANY "waterfall" (round 1: ~550 clocks per k step against 384 of tensor-pipe time).
// Round 2 (tools/probe_kstep.cu, profiles/r02_probes.txt): with a run-time ring stage the loop still needed 410-417 clocks per k
// step next to busy epilogue warps in the stand-alone probe and 480-770 in the kernels (clock64 timeline); unrolled over the ring
// stages (barrier addresses and B descriptors become constants, the A descriptors advance by a constant) and without the
// tcgen05 fence behind the weight barrier (the weights arrive through the async proxy; only the a_ready waits order generic-proxy
// stores of the epilogue warps) it paces at the tensor pipe's 384.
// TL (development probe, tools/timeline.py): clock64 stamps of CTA 0's second tile: tl[2048 + op * 32 + ks] = k step issued,
// tl[4096 + op * 32 + ks] = A chunk of an even k step seen, tl[6144 + op] = accumulator committed.
template <bool TL = false>
__device__ __forceinline__ void chain_mma(const OpTable& T, long long ntiles, uint32_t tmem_base, uint8_t* A_hi, uint8_t* A_lo, uint8_t* ring,
                                          uint64_t* full, uint64_t* empty, uint64_t* a_ready, uint64_t* d_full, long long* tl = nullptr) {
    const uint32_t a_hi_s = smem_u32(A_hi), a_lo_s = smem_u32(A_lo), ring_s = smem_u32(ring);
    const uint64_t dA_hi0 = smem_desc(a_hi_s, LBO_A, SBO), dA_lo0 = smem_desc(a_lo_s, LBO_A, SBO);
    constexpr uint64_t kStepA = (uint64_t)((2u * LBO_A) >> 4);
    uint32_t phase = 0, aphase = 0, g = 0;      // g: global op counter -> TMEM buffer g & 1
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        for (int op = 0; op < T.nops; ++op, ++g) {
            const int n = T.ops[op].n;
            const uint32_t idesc = T.f16 ? instr_desc_f16(TM, n) : instr_desc_bf16(TM, n);
            const uint32_t lbo_b = (uint32_t)n * 16u, lo_off = (uint32_t)n * 32u;
            const uint64_t dB0 = smem_desc(ring_s, lbo_b, SBO);          // stage 0, hi part; stage / lo part = address-field adds
            const uint32_t d_tmem = tmem_base + (g & 1u) * 256u;
            const int nks = T.ops[op].ksteps;
            uint64_t da_hi = dA_hi0, da_lo = dA_lo0;
            const bool rec = TL && tl && blockIdx.x == 0 && tile == (long long)gridDim.x && (threadIdx.x & 31) == 0;
#pragma unroll 1
            for (int ks0 = 0; ks0 < nks; ks0 += NSTAGE) {
#pragma unroll
                for (int st = 0; st < NSTAGE; ++st) {
                    const int ks = ks0 + st;
                    const bool real = ks < nks;
                    if ((st & 1) == 0 && real) {
                        const int c = ks >> 1;
                        mbar_wait(&a_ready[c], (aphase >> c) & 1u);
                        aphase ^= (1u << c);
                        tc_fence_after();
                        if (TL && rec) tl[4096 + op * 32 + ks] = clock64();
                    }
                    mbar_wait(&full[st], phase);
                    const uint64_t db_hi = dB0 + (uint64_t)(((uint32_t)st * (uint32_t)STAGE_MAX) >> 4);
                    const uint64_t db_lo = db_hi + (uint64_t)(lo_off >> 4);
                    if (elect_one_sync()) {
                        if (real) {
                            mma_bf16_ss(d_tmem, da_hi, db_hi, idesc, ks > 0 ? 1u : 0u);
                            mma_bf16_ss(d_tmem, da_lo, db_hi, idesc, 1u);
                            mma_bf16_ss(d_tmem, da_hi, db_lo, idesc, 1u);
                        }
                        mma_commit(&empty[st]);
                    }
                    __syncwarp();
                    if (TL && rec) tl[2048 + op * 32 + ks] = clock64();
                    da_hi += kStepA;
                    da_lo += kStepA;
                }
                phase ^= 1;
            }
            if (elect_one_sync()) mma_commit(&d_full[g & 1u]);
            __syncwarp();
            if (TL && rec) tl[6144 + op] = clock64();
        }
    }
}

}  // namespace chain
}  // namespace i2sdf
