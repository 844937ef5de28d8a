// Hand-written sm_100a primitives: mbarrier, bulk async copy, TMEM allocation, tcgen05.mma / ld / commit, and the
// shared-memory operand layout used by the tensor-core MLP kernels.
//
// Operand layout ("K-major, no swizzle" canonical UMMA layout, bf16):
//   a 16-byte row segment = 8 consecutive k for one row; a core matrix = 8 rows x 16 B stored contiguously (128 B);
//   core matrices of consecutive 8-row groups are contiguous (SBO = 128 B); consecutive 8-wide k chunks are
//   ROWS*16 B apart (LBO).  Element (row r, k) lives at  (k/8)*ROWS*16 + (r/8)*128 + (r%8)*16 + (k%8)*2  bytes.
//   One MMA (K=16) consumes two k chunks: its descriptor starts at chunk 2*kstep.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace i2sdf {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// First 1024-byte aligned address of the dynamic shared memory, derived by pointer arithmetic on the `extern __shared__` array
// so that the compiler still knows every pointer built from it is shared memory.  (Rounding the pointer through uintptr_t
// makes it generic: the operand stores of the epilogues then compile to ST.E.128 with 64-bit address arithmetic and a
// per-item S2UR SR_CgaCtaId / SR_SWINHI window rebuild instead of STS.128 - seen in the SASS of every chain kernel.)
__device__ __forceinline__ uint8_t* smem_align1024(uint8_t* raw) { return raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u); }
// (the operand layouts used here are un-swizzled: descriptors and bulk copies need 16-byte alignment only)
__device__ __forceinline__ uint8_t* smem_align128(uint8_t* raw) { return raw + ((128u - (smem_u32(raw) & 127u)) & 127u); }

// ---- mbarrier ------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}

// one lane of a CONVERGED warp (elect.sync): code under this predicate may use warp-uniform operands without a waterfall
__device__ __forceinline__ bool elect_one_sync() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
        "elect.sync rx|px, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, px;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

// ---- bulk async copy global -> shared, completion on an mbarrier (SASS: UBLKCP) ----------------------
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// the same with an L2 eviction-priority hint (createpolicy): the 3-6 MB of packed weights are re-read by every CTA for every tile and
// must survive the streaming scratch / slot traffic of the chain kernels in L2
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void bulk_g2s_hint(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar, uint64_t pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
                 : "memory");
}
// streaming (evict-first) 16-byte global accesses for data that is written once and read once (per-CTA scratch)
__device__ __forceinline__ void stg_cs(void* p, const uint4& v) {
    asm volatile("st.global.cs.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 ldg_cs(const void* p) {
    uint4 v;
    asm volatile("ld.global.cs.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
// generic-proxy smem writes -> visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- TMEM ----------------------------------------------------------------------------------------
template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result) {   // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "n"(COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {       // same warp that allocated
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 32 lanes x 32 consecutive fp32 columns: thread i of the warp gets lane (taddr.lane + i)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- tcgen05.mma ---------------------------------------------------------------------------------
// shared-memory matrix descriptor, K-major / no swizzle (cute::UMMA::SmemDescriptor, version 1)
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);            // start address   bits [0,14)
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;   // leading byte offset (between k chunks)  bits [16,30)
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;   // stride byte offset (between 8-row groups) bits [32,46)
    d |= (uint64_t)1 << 46;                              // descriptor version 1 (Blackwell)
    return d;                                            // base_offset 0, lbo_mode 0, layout SWIZZLE_NONE (0)
}
// instruction descriptor: kind::f16, A/B = bf16 K-major, D = fp32, M x N   (cute::UMMA::InstrDescriptor)
__host__ __device__ constexpr uint32_t instr_desc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// same with A/B = fp16 (a_format = b_format = 0).  The two 16-bit formats cannot be mixed in one instruction: A bf16 x B fp16 raises
// "illegal instruction" on the B200 (tools/probe_fmt.cu), although the descriptor has independent format fields.
__host__ __device__ constexpr uint32_t instr_desc_f16(int M, int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// D[tmem] (+)= A[smem] * B[smem]^T ; issued by ONE thread
__device__ __forceinline__ void mma_bf16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// all previously issued MMAs of this thread complete -> one arrival on the mbarrier
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- operand layout helpers -------------------------------------------------------------------------
// byte offset of the 16-byte segment holding k chunk `kc` (8 k's) of row r, for a tile with ROWS rows
template <int ROWS>
__device__ __host__ __forceinline__ constexpr uint32_t seg_off(int r, int kc) {
    return (uint32_t)kc * (ROWS * 16) + (uint32_t)(r >> 3) * 128 + (uint32_t)(r & 7) * 16;
}

// split fp32 into bf16 hi (round-to-nearest) + bf16 lo (rounded residual): x ~= hi + lo with ~16 mantissa bits
__device__ __forceinline__ void split_bf16x2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
    float2 hf = __bfloat1622float2(h);
    __nv_bfloat162 l = __floats2bfloat162_rn(x0 - hf.x, x1 - hf.y);
    hi = *reinterpret_cast<uint32_t*>(&h);
    lo = *reinterpret_cast<uint32_t*>(&l);
}

// split fp32 into fp16 hi + fp16 lo (11 + 11 mantissa bits; the low half of a small value is an fp16 subnormal, absolute error
// <= 2^-25): with 3 products and fp32 accumulation in TMEM a K = 256 layer is as accurate as an fp32 FMA chain (measured on the
// B200, tools/probe_fmt.cu: max error 1.1e-6 of max|D| against 7.0e-7 for the FMA chain and 5.9e-6 for the bf16 split).
// satfinite: a value beyond fp16's range (|x| > 65504) saturates instead of turning the whole tile into NaNs.
__device__ __forceinline__ void split_f16x2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    uint32_t h, l;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(x1), "f"(x0));
    const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&h));
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(l) : "f"(x1 - hf.y), "f"(x0 - hf.x));
    hi = h;
    lo = l;
}
// value pair of one 32-bit word of a hi plane + the same word of the lo plane
__device__ __forceinline__ float2 join_f16x2(uint32_t hi, uint32_t lo) {
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&hi)), b = __half22float2(*reinterpret_cast<const __half2*>(&lo));
    return make_float2(a.x + b.x, a.y + b.y);
}
__device__ __forceinline__ float2 join_bf16x2(uint32_t hi, uint32_t lo) {
    return make_float2(__uint_as_float(hi << 16) + __uint_as_float(lo << 16), __uint_as_float(hi & 0xffff0000u) + __uint_as_float(lo & 0xffff0000u));
}
// The fp16 weight blocks hold W * kWScale (a power of two: exact), so that the low halves of typical weights (|W| ~ 0.05) are
// normal fp16 numbers; the accumulator is multiplied by kWInv in the epilogue (folded into the bias add: one FFMA).
constexpr float kWScale = 256.0f, kWInv = 1.0f / 256.0f;

// softplus_100 in the overflow-free form max(a,0) + log1p(exp(-|100 a|))/100; equals nn.Softplus(beta=100,
// threshold=20) to < 3e-11 absolute (the thresholded branch differs from the exact value by log1p(e^-20)/100).
__device__ __forceinline__ float softplus100_fast(float a) {
    float e = exp2f(-fabsf(a) * 144.26950408889634f);            // exp(-|100 a|)   (MUFU.EX2)
    return fmaf(__log2f(1.0f + e), 0.0069314718055994531f, fmaxf(a, 0.0f));   // + ln2/100 * log2(1+e)   (MUFU.LG2)
}

// sin & cos for |a| < ~1e4: 3-constant Cody-Waite reduction to [-pi/4, pi/4] + minimax polynomials (~1 ulp)
__device__ __forceinline__ void sincos_cw(float a, float& s, float& c) {
    float q = rintf(a * 0.63661977236758134f);
    int n = (int)q;
    float r = fmaf(q, -1.5707962512969971f, a);
    r = fmaf(q, -7.5497894158615964e-08f, r);
    r = fmaf(q, -5.3903029534742384e-15f, r);
    float r2 = r * r;
    float sp = fmaf(fmaf(fmaf(-1.9515295891e-4f, r2, 8.3321608736e-3f), r2, -1.6666654611e-1f), r2 * r, r);
    float cp = fmaf(fmaf(fmaf(2.443315711809948e-5f, r2, -1.388731625493765e-3f), r2, 4.166664568298827e-2f), r2 * r2,
                    fmaf(-0.5f, r2, 1.0f));
    float ss = (n & 1) ? cp : sp;
    float cc = (n & 1) ? sp : cp;
    s = (n & 2) ? -ss : ss;
    c = ((n + 1) & 2) ? -cc : cc;
}

// value of embedding column i (0..38) of point x:  [x, sin(2^k x), cos(2^k x)]   (embedder.py:28-38)
__device__ __forceinline__ float embed_col(const float (&x)[3], int i, int mx) {
    if (i < 3) return x[i];
    const int qq = i - 3, k = qq / 6, cc = qq % 3;
    if (k >= mx) return 0.f;
    float s, c;
    sincos_cw(__fmul_rn(x[cc], (float)(1 << k)), s, c);
    return ((qq % 6) < 3) ? s : c;
}


}  // namespace tc
}  // namespace i2sdf

namespace i2sdf {
namespace tc {
// 32 lanes x 16 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2_approx(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
}  // namespace tc
}  // namespace i2sdf
