// Stand-alone numerics probe (development tool, NOT part of libi2sdf_b200.so).
// Questions it answers on the B200:
//   1. does tcgen05.mma kind::f16 accept A and B in DIFFERENT 16-bit formats (A bf16 x B fp16 and the reverse)?  The instruction
//      descriptor has independent a_format / b_format fields; the backward chain wants bf16 adjoints (range) against fp16 weights.
//   2. how accurate is a 3-product split  D = A_hi B_hi + A_lo B_hi + A_hi B_lo  with fp32 accumulation in TMEM when the halves are
//      bf16 (8 + 8 mantissa bits, what round 1 shipped) vs fp16 (11 + 11 bits, operands pre-scaled by powers of two so the low
//      halves stay normal), against a float64 product of the same fp32 inputs?  K = 256 (one layer of the SDF stack).
// Build + run:  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I i2sdf_b200/csrc tools/probe_fmt.cu -o tools/probe_fmt && ./tools/probe_fmt
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_fp16.h>

#include "tc_common.cuh"

using namespace i2sdf::tc;

constexpr int TM = 128, TN = 64, TK = 256;
constexpr int A_BYTES = TM * TK * 2, B_BYTES = TN * TK * 2;
constexpr size_t kSmem = 1024 + 2 * (size_t)A_BYTES + 2 * (size_t)B_BYTES + 64;

// fmt: 0 = fp16, 1 = bf16
__host__ __device__ constexpr uint32_t idesc(int M, int N, int fa, int fb) {
    return (1u << 4) | ((uint32_t)fa << 7) | ((uint32_t)fb << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ uint16_t cvt16(float x, int fmt) {
    if (fmt) { __nv_bfloat16 h = __float2bfloat16_rn(x); return *reinterpret_cast<uint16_t*>(&h); }
    __half h = __float2half_rn(x);
    return *reinterpret_cast<uint16_t*>(&h);
}
__device__ __forceinline__ float back16(uint16_t v, int fmt) {
    if (fmt) return __uint_as_float((uint32_t)v << 16);
    return __half2float(*reinterpret_cast<__half*>(&v));
}

struct Args { const float* A; const float* B; float* D; int fa, fb, nprod; float sa, sb; int order; };   // order 1: all cross terms first, then the hi*hi products

__global__ void __launch_bounds__(128, 1) probe(const Args a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_align1024(smem_raw);
    uint8_t* A_hi = smem;
    uint8_t* A_lo = A_hi + A_BYTES;
    uint8_t* B_hi = A_lo + A_BYTES;
    uint8_t* B_lo = B_hi + B_BYTES;
    uint64_t* done = reinterpret_cast<uint64_t*>(B_lo + B_BYTES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < TM * TK; i += 128) {
        const int r = i / TK, k = i % TK;
        const float x = a.A[i] * a.sa;
        const uint16_t h = cvt16(x, a.fa), l = cvt16(x - back16(h, a.fa), a.fa);
        const uint32_t off = seg_off<TM>(r, k >> 3) + (k & 7) * 2;
        *reinterpret_cast<uint16_t*>(A_hi + off) = h;
        *reinterpret_cast<uint16_t*>(A_lo + off) = l;
    }
    for (int i = tid; i < TN * TK; i += 128) {
        const int n = i / TK, k = i % TK;
        const float x = a.B[i] * a.sb;
        const uint16_t h = cvt16(x, a.fb), l = cvt16(x - back16(h, a.fb), a.fb);
        const uint32_t off = seg_off<TN>(n, k >> 3) + (k & 7) * 2;
        *reinterpret_cast<uint16_t*>(B_hi + off) = h;
        *reinterpret_cast<uint16_t*>(B_lo + off) = l;
    }
    if (tid == 0) { mbar_init(done, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc<64>(tmem_slot);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (warp == 0) {
        const uint32_t id = idesc(TM, TN, a.fa, a.fb);
        if (a.order == 1) {
            for (int ph = 0; ph < 2; ++ph)
                for (int ks = 0; ks < TK / 16; ++ks) {
                    const uint64_t dah = smem_desc(smem_u32(A_hi) + ks * 2 * TM * 16, TM * 16, 128), dal = smem_desc(smem_u32(A_lo) + ks * 2 * TM * 16, TM * 16, 128);
                    const uint64_t dbh = smem_desc(smem_u32(B_hi) + ks * 2 * TN * 16, TN * 16, 128), dbl = smem_desc(smem_u32(B_lo) + ks * 2 * TN * 16, TN * 16, 128);
                    if (elect_one_sync()) {
                        if (ph == 0) { mma_bf16_ss(tmem_base, dal, dbh, id, ks > 0 ? 1u : 0u); mma_bf16_ss(tmem_base, dah, dbl, id, 1u); }
                        else mma_bf16_ss(tmem_base, dah, dbh, id, 1u);
                    }
                    __syncwarp();
                }
        } else
        for (int ks = 0; ks < TK / 16; ++ks) {
            const uint64_t dah = smem_desc(smem_u32(A_hi) + ks * 2 * TM * 16, TM * 16, 128), dal = smem_desc(smem_u32(A_lo) + ks * 2 * TM * 16, TM * 16, 128);
            const uint64_t dbh = smem_desc(smem_u32(B_hi) + ks * 2 * TN * 16, TN * 16, 128), dbl = smem_desc(smem_u32(B_lo) + ks * 2 * TN * 16, TN * 16, 128);
            if (elect_one_sync()) {
                mma_bf16_ss(tmem_base, dah, dbh, id, ks > 0 ? 1u : 0u);
                if (a.nprod > 1) { mma_bf16_ss(tmem_base, dal, dbh, id, 1u); mma_bf16_ss(tmem_base, dah, dbl, id, 1u); }
                if (a.nprod > 3) mma_bf16_ss(tmem_base, dal, dbl, id, 1u);
            }
            __syncwarp();
        }
        if (elect_one_sync()) mma_commit(done);
        __syncwarp();
    }
    mbar_wait(done, 0);
    tc_fence_after();
    const float inv = 1.0f / (a.sa * a.sb);
    for (int c0 = 0; c0 < TN; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, v);
        tmem_ld_wait();
        for (int j = 0; j < 32; ++j) a.D[(size_t)tid * TN + c0 + j] = __uint_as_float(v[j]) * inv;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<64>(tmem_base);
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

static double frand() { return (double)rand() / RAND_MAX; }
static double nrand() { return sqrt(-2.0 * log(frand() + 1e-300)) * cos(6.283185307179586 * frand()); }

int main() {
    srand(1);
    std::vector<float> A(TM * TK), B(TN * TK), D(TM * TN);
    // activations like softplus_100 outputs of the SDF stack (non-negative, many near zero), weights ~ N(0, sqrt(2/256))
    for (auto& x : A) { double t = 0.3 * nrand(); x = (float)(t > 0 ? t : 1e-3 * exp(10 * t)); }
    for (auto& x : B) x = (float)(0.0884 * nrand());
    std::vector<double> ref(TM * TN);
    double refmax = 0;
    for (int r = 0; r < TM; ++r)
        for (int n = 0; n < TN; ++n) {
            double s = 0;
            for (int k = 0; k < TK; ++k) s += (double)A[r * TK + k] * (double)B[n * TK + k];
            ref[r * TN + n] = s;
            refmax = fmax(refmax, fabs(s));
        }
    // fp32 sequential sum, as a yardstick
    double e32 = 0;
    for (int r = 0; r < TM; ++r)
        for (int n = 0; n < TN; ++n) {
            float s = 0;
            for (int k = 0; k < TK; ++k) s = fmaf(A[r * TK + k], B[n * TK + k], s);
            e32 = fmax(e32, fabs((double)s - ref[r * TN + n]));
        }
    printf("yardstick: fp32 FMA chain  max|err|/max|D| = %.3e\n", e32 / refmax);
    float *dA, *dB, *dD;
    CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem));
    struct V { const char* name; int fa, fb, nprod; float sa, sb; int order; };
    const V vs[] = {
        {"bf16 x bf16, 1 product              ", 1, 1, 1, 1.f, 1.f, 0},
        {"bf16 x bf16, 3 products (round 1)   ", 1, 1, 3, 1.f, 1.f, 0},
        {"bf16 x bf16, 4 products             ", 1, 1, 4, 1.f, 1.f, 0},
        {"fp16 x fp16, 1 product              ", 0, 0, 1, 1.f, 1.f, 0},
        {"fp16 x fp16, 3 products, no scaling ", 0, 0, 3, 1.f, 1.f, 0},
        {"fp16 x fp16, 3 products, 16 / 4096  ", 0, 0, 3, 16.f, 4096.f, 0},
        {"fp16 x fp16, 3 products, 64 / 16384 ", 0, 0, 3, 64.f, 16384.f, 0},
        {"fp16 x fp16, 4 products, 64 / 16384 ", 0, 0, 4, 64.f, 16384.f, 0},
        {"fp16 x fp16, 3 products, 1 / 256    ", 0, 0, 3, 1.f, 256.f, 0},
        {"same, cross terms first             ", 0, 0, 3, 1.f, 256.f, 1},
    };
    for (int pass = 0; pass < 2; ++pass) {
    if (pass == 1) {
        printf("---- all-positive weights (every product and every partial sum positive): a round-toward-zero accumulator shows as a negative mean error\n");
        for (auto& x : B) x = fabsf(x);
        CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
        refmax = 0;
        for (int r = 0; r < TM; ++r)
            for (int n = 0; n < TN; ++n) {
                double s = 0;
                for (int k = 0; k < TK; ++k) s += (double)A[r * TK + k] * (double)B[n * TK + k];
                ref[r * TN + n] = s;
                refmax = fmax(refmax, fabs(s));
            }
    }
    for (const V& v : vs) {
        Args a{dA, dB, dD, v.fa, v.fb, v.nprod, v.sa, v.sb, v.order};
        CK(cudaMemset(dD, 0, D.size() * 4));
        probe<<<1, 128, kSmem>>>(a);
        CK(cudaGetLastError());
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s: LAUNCH FAILED: %s\n", v.name, cudaGetErrorString(e)); return 1; }
        CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
        double emax = 0, erms = 0, sed = 0, sdd = 0;
        for (size_t i = 0; i < D.size(); ++i) {
            double d = (double)D[i] - ref[i];
            emax = fmax(emax, fabs(d)); erms += d * d;
            sed += d * ref[i]; sdd += ref[i] * ref[i];
        }
        // least-squares fit  err = -kappa * D  (a round-toward-zero accumulator shrinks |D|) and what remains after scaling D by (1 + kappa)
        const double kappa = -sed / sdd;
        double rres = 0, rmax = 0;
        for (size_t i = 0; i < D.size(); ++i) { double d = (double)D[i] * (1.0 + kappa) - ref[i]; rres += d * d; rmax = fmax(rmax, fabs(d)); }
        printf("%s max|err|/max|D| = %.3e   rms/max = %.3e   shrink kappa = %+.3e   after (1+kappa): max %.3e rms %.3e\n", v.name, emax / refmax,
               sqrt(erms / D.size()) / refmax, kappa, rmax / refmax, sqrt(rres / D.size()) / refmax);
    }
    }
    return 0;
}
