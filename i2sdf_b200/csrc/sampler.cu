// Per-ray kernels: camera rays, VolSDF error-bounded sampler rounds, final sample set, alpha compositing.
// One warp owns one ray; its z / sdf / d* / cdf arrays (<= 640 entries) live in shared memory and every scan is a
// lane-local sequential pass + a warp shuffle scan.  Compiled with -fmad=false: elementwise arithmetic is the
// same sequence of IEEE fp32 operations the reference's ATen CPU kernels perform, and prefix sums accumulate in
// double and round once, as torch's CPU cumsum does for fp32 inputs.
//
// Replaces (reference file:line): utils/rend_util.py:92-147 (rays); model/network/ray_sampler.py:22-43 (uniform
// init + jitter), :67-251 (ErrorBoundSampler.get_z_vals, get_error_bound); model/network/density.py:21-30;
// model/network/__init__.py:118-125, 204-219, 223-240 (volume_rendering and the per-ray reductions).
#include <float.h>
#include "common.cuh"

// exp of the up-sampling pdf.  pdf = (min(exp(E), 1e6) - 1) T + 1e-6 amplifies a last-bit difference of exp(E) at small E (rays that miss
// the surface: E ~ 1e-7, exp(E) - 1 is 0 or 1 ulp) into a different inverse-CDF sample set.  torch's CPU exp (Sleef, 1 ulp) agrees with
// the correctly rounded value for ~99 % of arguments, CUDA's expf (2 ulp) with torch's for ~90 %: that ONE exp is evaluated in double and
// rounded once.  Measured at 1024 W-sharp rays (profiles/parity_r02.json): z's within 1e-3 of the CPU reference 0.947 -> 0.985 of all samples
// (0.776 -> 0.855 within 1e-5).  Doing the same for every exp / expm1 of the beta search costs +0.13 ms per step and changes no output
// (I2SDF_SAMPLER_EXP64=1 at compile time).
#ifndef I2SDF_SAMPLER_EXP64
#define I2SDF_SAMPLER_EXP64 0
#endif
#if I2SDF_SAMPLER_EXP64
#define EXPF(x) ((float)exp((double)(x)))
#define EXPM1F(x) ((float)expm1((double)(x)))
#else
#define EXPF(x) expf(x)
#define EXPM1F(x) expm1f(x)
#endif
#define EXPF_CR(x) ((float)exp((double)(x)))

namespace i2sdf {

constexpr int kWarpsPerCta = 4;
constexpr int kZMax = 640 + 1;

__device__ __forceinline__ float beta0_of(const float* beta_param, float beta_min) {
    return fabsf(*beta_param) + beta_min;            // density.py:28-30
}

// LaplaceDensity.density_func (density.py:25-26): alpha * (0.5 + 0.5*sign(s)*expm1(-|s|/beta)), alpha = 1/beta
__device__ __forceinline__ float laplace_density(float s, float beta, float alpha) {
    float sg = (s > 0.f) ? 1.f : ((s < 0.f) ? -1.f : 0.f);
    return alpha * (0.5f + (0.5f * sg) * EXPM1F(-fabsf(s) / beta));
}

__device__ __forceinline__ double warp_excl_scan(double v, int lane, double* total) {
    double x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        double y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    *total = __shfl_sync(0xffffffffu, x, 31);
    return x - v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_sumf(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sumd(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ bool sampler_round_active(const float* beta_max, int k, float b0) {
    for (int j = 0; j < k; ++j)
        if (!(beta_max[j] > b0)) return false;
    return true;
}

// ------------------------------------------------------------------------------------------------
// rays
// ------------------------------------------------------------------------------------------------
__global__ void rays_kernel(const float* __restrict__ uv, const float* __restrict__ pose,
                            const float* __restrict__ intr, int B, int Pn, float* __restrict__ o,
                            float* __restrict__ d, float* __restrict__ dnorm) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)B * Pn) return;
    int b = (int)(i / Pn);
    const float* K = intr + b * 16;
    const float* Pm = pose + b * 16;
    float fx = K[0], fy = K[5], cx = K[2], cy = K[6], sk = K[1];
    float u = uv[i * 2], v = uv[i * 2 + 1];
    float xl = (u - cx + cy * sk / fy - sk * v / fy) / fx * 1.0f;      // rend_util.py:143
    float yl = (v - cy) / fy * 1.0f;
    float pc[4] = {xl, yl, 1.0f, 1.0f};
    float w[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        float acc = Pm[r * 4] * pc[0];
        acc = fmaf(Pm[r * 4 + 1], pc[1], acc);
        acc = fmaf(Pm[r * 4 + 2], pc[2], acc);
        acc = fmaf(Pm[r * 4 + 3], pc[3], acc);
        w[r] = acc - Pm[r * 4 + 3];                                    // world - cam_loc
    }
    float n = sqrtf(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
    float den = fmaxf(n, 1e-12f);                                      // F.normalize eps
    dnorm[i] = n;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        o[i * 3 + r] = Pm[r * 4 + 3];
        d[i * 3 + r] = w[r] / den;
    }
}

// ------------------------------------------------------------------------------------------------
// sampler: init
// ------------------------------------------------------------------------------------------------
__global__ void sampler_init_kernel(SamplerDev S, SamplerWs W, long long R, const float* __restrict__ jitter,
                                    float lemma2_coeff) {
    const int lane = threadIdx.x & 31;
    const long long r = (long long)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
    if (r == 0 && lane < S.max_iters) W.beta_max[lane] = 0.f;
    if (r >= R) return;
    const int n = S.n_eval;
    float* z = W.z[0] + r * W.zmax;
    float* smp = W.samples + r * S.n_eval;
    auto zlin = [&](int j) { float t = S.t_init[j]; return S.near_ * (1.0f - t) + S.far_ * t; };   // ray_sampler.py:30-31
    double ss = 0.0;
    for (int j = lane; j < n; j += 32) {
        float zj = zlin(j);
        if (jitter) {                                                                              // :33-41
            float lower = (j == 0) ? zj : 0.5f * (zj + zlin(j - 1));
            float upper = (j == n - 1) ? zj : 0.5f * (zlin(j + 1) + zj);
            zj = lower + (upper - lower) * jitter[r * n + j];
        }
        z[j] = zj;
        smp[j] = zj;
    }
    __syncwarp();
    for (int j = lane; j < n - 1; j += 32) {
        float dd = z[j + 1] - z[j];
        ss += (double)(dd * dd);
    }
    ss = warp_sumd(ss);
    if (lane == 0) W.beta[r] = sqrtf(lemma2_coeff * (float)ss);                                    // :75-77
}

// ------------------------------------------------------------------------------------------------
// sampler rounds: one CTA (128 threads) per ray.  z / sdf / d* / cdf live in shared memory (<= 640 entries), every
// thread owns <= 5 contiguous sections in registers; prefix sums are lane-local + warp shuffle + 4 warp totals,
// accumulated in double and rounded once (torch's CPU cumsum semantics).
// ------------------------------------------------------------------------------------------------
constexpr int kRayThreads = 128;
constexpr int kPer = 5;                     // ceil(640 / 128)

struct RayArrays {
    float* z;       // [n]
    float* s;       // [n] sdf
    float* ds;      // [n-1] d*
    float* cdf;     // [n]
    double* scan;   // [8] warp totals of the two running sums
    float* red;     // [4] warp maxima
};

__device__ __forceinline__ RayArrays carve_ray(float* smem) {
    RayArrays A;
    A.z = smem; A.s = smem + kZMax; A.ds = smem + 2 * kZMax; A.cdf = smem + 3 * kZMax;
    A.scan = reinterpret_cast<double*>(smem + 4 * kZMax + 4);   // 8-byte aligned: 4*641+4 = 2568 floats
    A.red = smem + 4 * kZMax + 4 + 16;
    return A;
}
constexpr size_t kSamplerSmem = (size_t)(4 * kZMax + 4 + 16 + 8) * sizeof(float);

// exclusive prefix (over threads, in thread order) of two per-thread sums
__device__ __forceinline__ void block_scan2(double& a, double& b, double* sh, int warp, int lane) {
    double ia = a, ib = b;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        double ya = __shfl_up_sync(0xffffffffu, ia, o), yb = __shfl_up_sync(0xffffffffu, ib, o);
        if (lane >= o) { ia += ya; ib += yb; }
    }
    if (lane == 31) { sh[warp] = ia; sh[4 + warp] = ib; }
    __syncthreads();
    double pa = 0.0, pb = 0.0;
    for (int w = 0; w < warp; ++w) { pa += sh[w]; pb += sh[4 + w]; }
    a = ia - a + pa;
    b = ib - b + pb;
}
__device__ __forceinline__ float block_max(float v, float* sh, int warp, int lane) {
    v = warp_max(v);
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    return fmaxf(fmaxf(sh[0], sh[1]), fmaxf(sh[2], sh[3]));
}
__device__ __forceinline__ double block_sum(double v, double* sh, int warp, int lane) {
    v = warp_sumd(v);
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    return (sh[0] + sh[1]) + (sh[2] + sh[3]);
}

// Theorem-1 bound per section   (ray_sampler.py:98-114)
__device__ __forceinline__ void compute_dstar(const RayArrays& A, int n) {
    for (int i = threadIdx.x; i < n - 1; i += kRayThreads) {
        float a = A.z[i + 1] - A.z[i];
        float s0 = A.s[i], s1 = A.s[i + 1];
        float b = fabsf(s0), c = fabsf(s1);
        float a2 = a * a, b2 = b * b, c2 = c * c;
        bool c1 = (a2 + b2) <= c2;
        bool c2_ = (a2 + c2) <= b2;
        float sp = (a + b + c) / 2.0f;
        float area = sp * (sp - a) * (sp - b) * (sp - c);
        bool tri = !c1 && !c2_ && ((b + c - a) > 0.f);
        c1 = c1 && !c2_;
        float h = (2.0f * sqrtf(area)) / a;
        if (isnan(h)) h = 0.f;                       // nan_to_num
        else if (isinf(h)) h = h > 0.f ? FLT_MAX : -FLT_MAX;
        float dstar = ((c1 ? b : 0.f) + (c2_ ? c : 0.f)) + h * (tri ? 1.f : 0.f);
        float sg0 = (s0 > 0.f) ? 1.f : ((s0 < 0.f) ? -1.f : 0.f);
        float sg1 = (s1 > 0.f) ? 1.f : ((s1 < 0.f) ? -1.f : 0.f);
        A.ds[i] = (sg1 * sg0 == 1.f) ? dstar : 0.f * dstar;
    }
    __syncthreads();
}

// get_error_bound (ray_sampler.py:243-251) for one ray; all threads return the max.
__device__ float error_bound(const RayArrays& A, int n, float beta) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cnt = n - 1;
    const int per = (cnt + kRayThreads - 1) / kRayThreads;
    const int i0 = threadIdx.x * per, i1 = min(i0 + per, cnt);
    const float alpha = 1.0f / beta;
    const float four_b2 = 4.0f * (beta * beta);
    float fe[kPer], ee[kPer];
    double sumI = 0.0, sumE = 0.0;
#pragma unroll
    for (int q = 0; q < kPer; ++q) {
        const int i = i0 + q;
        fe[q] = 0.f; ee[q] = 0.f;
        if (i < i1) {
            float dist = A.z[i + 1] - A.z[i];
            fe[q] = dist * laplace_density(A.s[i], beta, alpha);
            ee[q] = EXPF(-A.ds[i] / beta) * (dist * dist) / four_b2;
            sumI += (double)fe[q];
            sumE += (double)ee[q];
        }
    }
    double accI = sumI, accE = sumE;
    block_scan2(accI, accE, A.scan, warp, lane);
    float best = -FLT_MAX;
#pragma unroll
    for (int q = 0; q < kPer; ++q) {
        if (i0 + q < i1) {
            float I = (float)accI;                       // exclusive: integral_estimation[:, :-1]
            accE += (double)ee[q];
            float E = (float)accE;                       // inclusive
            float bo = (fminf(EXPF(E), 1.0e6f) - 1.0f) * EXPF(-I);
            best = fmaxf(best, bo);
            accI += (double)fe[q];
        }
    }
    return block_max(best, A.red, warp, lane);
}

// beta line search (ray_sampler.py:118-132)
__device__ float beta_search(const RayArrays& A, int n, float beta0, float beta_in, const SamplerDev& S) {
    float err = error_bound(A, n, beta0);
    float hi = (err <= S.eps) ? beta0 : beta_in;
    float lo = beta0;
    for (int it = 0; it < S.beta_iters; ++it) {
        float mid = (lo + hi) / 2.0f;
        err = error_bound(A, n, mid);
        bool ok = err <= S.eps;
        hi = ok ? mid : hi;
        lo = ok ? lo : mid;
    }
    return hi;
}

// pdf -> cdf -> inverse CDF (ray_sampler.py:139-207).  Leaves cdf in A.cdf[0..n-1]; writes ns samples (+ inds).
__device__ void resample(const RayArrays& A, int n, float beta, bool upsample, const SamplerDev& S, const float* __restrict__ u,
                         int ns, float* __restrict__ out_samples, int* __restrict__ out_inds) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cnt = n - 1;
    const int per = (n + kRayThreads - 1) / kRayThreads;          // covers i = 0..n-1 (last: dist = 1e10)
    const int i0 = threadIdx.x * per, i1 = min(i0 + per, n);
    const float alpha = 1.0f / beta;
    const float four_b2 = 4.0f * (beta * beta);
    float fe[kPer + 1], ee[kPer + 1];
    double sumF = 0.0, sumE = 0.0;
#pragma unroll
    for (int q = 0; q < kPer + 1; ++q) {
        const int i = i0 + q;
        fe[q] = 0.f; ee[q] = 0.f;
        if (i < i1) {
            float dist = (i < cnt) ? (A.z[i + 1] - A.z[i]) : 1.0e10f;
            fe[q] = dist * laplace_density(A.s[i], beta, alpha);
            sumF += (double)fe[q];
            if (upsample && i < cnt) {
                ee[q] = EXPF(-A.ds[i] / beta) * (dist * dist) / four_b2;
                sumE += (double)ee[q];
            }
        }
    }
    double accF = sumF, accE = sumE;
    block_scan2(accF, accE, A.scan, warp, lane);
    float pdf[kPer + 1];
    double psum = 0.0;
#pragma unroll
    for (int q = 0; q < kPer + 1; ++q) {
        const int i = i0 + q;
        pdf[q] = 0.f;
        if (i < i1) {
            float T = EXPF(-(float)accF);                         // transmittance (exclusive cumsum)
            accF += (double)fe[q];
            if (i < cnt) {
                if (upsample) {
                    accE += (double)ee[q];
                    pdf[q] = (fminf(EXPF_CR((float)accE), 1.0e6f) - 1.0f) * T + S.add_tiny;
                } else {
                    pdf[q] = (1.0f - EXPF(-fe[q])) * T + 1e-5f;
                }
                psum += (double)pdf[q];
            }
        }
    }
    __syncthreads();                                              // A.scan reuse
    const float total = (float)block_sum(psum, A.scan, warp, lane);
    double csum = 0.0, dummy = 0.0;
#pragma unroll
    for (int q = 0; q < kPer + 1; ++q) {
        const int i = i0 + q;
        if (i < i1 && i < cnt) { pdf[q] = pdf[q] / total; csum += (double)pdf[q]; }
    }
    __syncthreads();
    double acc = csum;
    block_scan2(acc, dummy, A.scan, warp, lane);
#pragma unroll
    for (int q = 0; q < kPer + 1; ++q) {
        const int i = i0 + q;
        if (i < i1 && i < cnt) { acc += (double)pdf[q]; A.cdf[i + 1] = (float)acc; }
    }
    if (threadIdx.x == 0) A.cdf[0] = 0.f;
    __syncthreads();
    const float* cdf = A.cdf;
    for (int j = threadIdx.x; j < ns; j += kRayThreads) {
        float uj = u[j];
        int lo = 0, hi = n;                                    // searchsorted(right=True): first idx with cdf > u
        while (lo < hi) {
            int mid = (lo + hi) >> 1;
            if (cdf[mid] <= uj) lo = mid + 1; else hi = mid;
        }
        int inds = lo;
        int below = max(inds - 1, 0), above = min(n - 1, inds);
        float cb = cdf[below], ca = cdf[above];
        float zb = A.z[below], za = A.z[above];
        float denom = ca - cb;
        if (denom < 1e-5f) denom = 1.0f;
        float t = (uj - cb) / denom;
        out_samples[j] = zb + t * (za - zb);
        if (out_inds) out_inds[j] = inds;
    }
    __syncthreads();
}

// stable merge of sorted z[0..n) (first on ties) with sorted smp[0..ns): values + source index (:211-212)
__device__ void merge_sorted(const float* __restrict__ z, int n, const float* __restrict__ smp, int ns,
                             float* __restrict__ out_z, int* __restrict__ out_src) {
    for (int i = threadIdx.x; i < n; i += kRayThreads) {
        float v = z[i];
        int lo = 0, hi = ns;                                   // #samples strictly less than v
        while (lo < hi) { int mid = (lo + hi) >> 1; if (smp[mid] < v) lo = mid + 1; else hi = mid; }
        out_z[i + lo] = v;
        out_src[i + lo] = i;
    }
    for (int j = threadIdx.x; j < ns; j += kRayThreads) {
        float v = smp[j];
        int lo = 0, hi = n;                                    // #z less than or equal to v
        while (lo < hi) { int mid = (lo + hi) >> 1; if (z[mid] <= v) lo = mid + 1; else hi = mid; }
        out_z[j + lo] = v;
        out_src[j + lo] = n + j;
    }
}

// ------------------------------------------------------------------------------------------------
// round k: merge sdf, d*, beta search, batch-global max(beta) - and, in the same launch, this ray's up-sampling for round k + 1.
// Whether round k + 1 happens is a batch-global decision (max over ALL rays of beta, ray_sampler.py:151) that no CTA can know before
// the launch ends, but when it does happen every ray is up-sampled with its own beta whatever its own state (:153-176), so the
// up-sampling is done speculatively here (the ray's z / sdf / d* are already in shared memory); if the batch turns out converged,
// round k + 1's launches predicate themselves off and the speculative samples / merged z (the OTHER z buffer) are never read.
// One launch per round instead of two: 12 -> 8 per-ray launches per forward, and d* / the z, sdf reload happen once.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kRayThreads) sampler_beta_kernel(SamplerDev S, SamplerWs W, long long R, int k, const float* __restrict__ beta_param) {
    extern __shared__ __align__(16) float smem[];
    __shared__ float s_smp[128];
    const long long r = blockIdx.x;
    const float b0 = beta0_of(beta_param, S.beta_min);
    if (!sampler_round_active(W.beta_max, k, b0)) return;
    const int n = S.n_eval * (k + 1), n_old = S.n_eval * k, cur = k & 1;
    RayArrays A = carve_ray(smem);
    const float* zg = W.z[cur] + r * W.zmax;
    float* sg = W.sdf[cur] + r * W.zmax;
    const float* sold = W.sdf[cur ^ 1] + r * W.zmax;
    const float* snew = W.sdf_new + r * S.n_eval;
    const int* src = W.src + r * W.zmax;
    for (int i = threadIdx.x; i < n; i += kRayThreads) {
        float v;
        if (k == 0) v = snew[i];
        else { int s = src[i]; v = (s < n_old) ? sold[s] : snew[s - n_old]; }       // :90-93
        A.z[i] = zg[i];
        A.s[i] = v;
        sg[i] = v;
    }
    __syncthreads();
    compute_dstar(A, n);
    const float beta = beta_search(A, n, b0, W.beta[r], S);
    if (threadIdx.x == 0) {
        W.beta[r] = beta;
        atomicMax(reinterpret_cast<int*>(W.beta_max + k), __float_as_int(beta));    // beta > 0
    }
    if (k + 1 >= S.max_iters) return;                                               // :151-153: the last round never up-samples
    __syncthreads();
    resample(A, n, beta, true, S, S.u_up, S.n_eval, s_smp, nullptr);
    float* smp = W.samples + r * S.n_eval;
    for (int j = threadIdx.x; j < S.n_eval; j += kRayThreads) smp[j] = s_smp[j];
    merge_sorted(A.z, n, s_smp, S.n_eval, W.z[cur ^ 1] + r * W.zmax, W.src + r * W.zmax);
}

// ------------------------------------------------------------------------------------------------
// after the last round that ran (klast: first round whose batch max(beta) met the bound, or max_iters - 1): the final N_samples from the
// opacity pdf of that round's samples (ray_sampler.py:178-207 with the final weights).  only_k >= 0: staged API, launched after every
// round - does nothing unless that round is klast.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kRayThreads) sampler_resample_kernel(SamplerDev S, SamplerWs W, long long R, int only_k,
                                                                       const float* __restrict__ beta_param, const float* __restrict__ u_final_tape) {
    extern __shared__ __align__(16) float smem[];
    __shared__ float s_smp[128];
    const long long r = blockIdx.x;
    const float b0 = beta0_of(beta_param, S.beta_min);
    int k = 0;
    while (k + 1 < S.max_iters && (W.beta_max[k] > b0)) ++k;
    if (only_k >= 0 && only_k != k) return;
    const int n = S.n_eval * (k + 1), cur = k & 1;
    RayArrays A = carve_ray(smem);
    const float* zg = W.z[cur] + r * W.zmax;
    const float* sg = W.sdf[cur] + r * W.zmax;
    for (int i = threadIdx.x; i < n; i += kRayThreads) { A.z[i] = zg[i]; A.s[i] = sg[i]; }
    __syncthreads();
    const int ns = S.n_samples;
    const float* u = u_final_tape ? u_final_tape + r * S.n_samples : S.u_final;
    resample(A, n, W.beta[r], false, S, u, ns, s_smp, nullptr);
    float* smp = W.samples + r * S.n_eval;
    for (int j = threadIdx.x; j < ns; j += kRayThreads) smp[j] = s_smp[j];
}

// ------------------------------------------------------------------------------------------------
// final sample set   (ray_sampler.py:215-234)
// ------------------------------------------------------------------------------------------------
__global__ void sampler_finalize_kernel(SamplerDev S, SamplerWs W, long long R, const float* __restrict__ beta_param,
                                        const int* __restrict__ extra_tape, int extra_per_round, const int* __restrict__ eik_idx,
                                        float* __restrict__ out_z, float* __restrict__ out_z_eik, int* __restrict__ info) {
    __shared__ float vals[kWarpsPerCta][128];
    __shared__ float sorted[kWarpsPerCta][128];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long r = (long long)blockIdx.x * kWarpsPerCta + warp;
    const float b0 = beta0_of(beta_param, S.beta_min);
    int klast = 0;                                            // index of the round that drew the final samples
    while (klast + 1 < S.max_iters && (W.beta_max[klast] > b0)) ++klast;
    const int n = S.n_eval * (klast + 1), cur = klast & 1;
    if (blockIdx.x == 0 && threadIdx.x == 0 && info) { info[0] = klast + 1; info[1] = n; }
    if (r >= R) return;
    const int ns = S.n_samples, ne = S.n_extra;
    const int total = ns + 2 + ne;
    const float* zg = W.z[cur] + r * W.zmax;
    const float* smp = W.samples + r * S.n_eval;
    // extra_tape: one index set, or (extra_per_round) one candidate set per possible round count [max_iters][ne]
    const int* eidx = extra_tape ? extra_tape + (extra_per_round ? klast * ne : 0) : (S.extra_idx + klast * ne);
    for (int i = lane; i < total; i += 32) {
        float v;
        if (i < ns) v = smp[i];
        else if (i == ns) v = S.near_;
        else if (i == ns + 1) v = S.far_;
        else v = zg[eidx[i - ns - 2]];
        vals[warp][i] = v;
    }
    __syncwarp();
    for (int i = lane; i < total; i += 32) {                  // rank sort (stable)
        float v = vals[warp][i];
        int rank = 0;
        for (int j = 0; j < total; ++j) {
            float w = vals[warp][j];
            rank += (w < v) || (w == v && j < i);
        }
        sorted[warp][rank] = v;
    }
    __syncwarp();
    for (int i = lane; i < total; i += 32) out_z[r * total + i] = sorted[warp][i];
    if (out_z_eik && eik_idx && lane == 0) out_z_eik[r] = sorted[warp][eik_idx[r]];
}

// ------------------------------------------------------------------------------------------------
// alpha compositing + per-ray reductions   (network/__init__.py:118-125, 204-219, 223-240)
// ------------------------------------------------------------------------------------------------
struct CompositeArgs {
    const float* z;        // [R][N+1]
    const float* dnorm;    // [R]
    const float* sdf;      // [R*N]
    const float* rgb;      // [R*N*3] or null
    const float* grad;     // [R*N*3] or null
    const float* lmask;    // [R*N] or null
    const float* beta_param;
    float beta_min;
    float* out_rgb; float* out_depth; float* out_wsum; float* out_normal; float* out_light; float* out_w;
    long long R; int N;
};

__global__ void composite_kernel(CompositeArgs C) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long r = (long long)blockIdx.x * kWarpsPerCta + warp;
    if (r >= C.R) return;
    const int N = C.N;
    const float beta = beta0_of(C.beta_param, C.beta_min);
    const float alpha = 1.0f / beta;
    const float* z = C.z + r * (N + 1);
    const int per = (N + 31) / 32;
    const int i0 = lane * per, i1 = min(i0 + per, N);
    float fe[8];
    double sum = 0.0;
    for (int i = i0, q = 0; i < i1; ++i, ++q) {
        float dist = z[i + 1] - z[i];                          // last one is z_max - z_last
        fe[q] = dist * laplace_density(C.sdf[r * N + i], beta, alpha);
        sum += (double)fe[q];
    }
    double tot;
    double acc = warp_excl_scan(sum, lane, &tot);
    float a_rgb[3] = {0.f, 0.f, 0.f}, a_n[3] = {0.f, 0.f, 0.f}, a_w = 0.f, a_d = 0.f, a_l = 0.f;
    for (int i = i0, q = 0; i < i1; ++i, ++q) {
        float T = EXPF(-(float)acc);
        acc += (double)fe[q];
        float w = (1.0f - EXPF(-fe[q])) * T;
        if (C.out_w) C.out_w[r * N + i] = w;
        a_w += w;
        a_d += w * z[i];
        if (C.rgb) {
            const float* c = C.rgb + (r * N + i) * 3;
            a_rgb[0] += w * c[0]; a_rgb[1] += w * c[1]; a_rgb[2] += w * c[2];
        }
        if (C.grad) {
            const float* g = C.grad + (r * N + i) * 3;
            float nn = fmaxf(sqrtf(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]), 1e-12f);     // F.normalize
            a_n[0] += w * (g[0] / nn); a_n[1] += w * (g[1] / nn); a_n[2] += w * (g[2] / nn);
        }
        if (C.lmask) a_l += w * C.lmask[r * N + i];
    }
    a_w = warp_sumf(a_w); a_d = warp_sumf(a_d); a_l = warp_sumf(a_l);
#pragma unroll
    for (int c = 0; c < 3; ++c) { a_rgb[c] = warp_sumf(a_rgb[c]); a_n[c] = warp_sumf(a_n[c]); }
    if (lane == 0) {
        if (C.out_rgb) { C.out_rgb[r * 3] = a_rgb[0]; C.out_rgb[r * 3 + 1] = a_rgb[1]; C.out_rgb[r * 3 + 2] = a_rgb[2]; }
        if (C.out_wsum) C.out_wsum[r] = a_w;
        if (C.out_depth) C.out_depth[r] = a_d / fmaxf(C.dnorm[r], 1e-6f);
        if (C.out_normal) {
            float nn = fmaxf(sqrtf(a_n[0] * a_n[0] + a_n[1] * a_n[1] + a_n[2] * a_n[2]), 1e-12f);
            C.out_normal[r * 3] = a_n[0] / nn; C.out_normal[r * 3 + 1] = a_n[1] / nn; C.out_normal[r * 3 + 2] = a_n[2] / nn;
        }
        if (C.out_light) C.out_light[r] = a_l;
    }
}

// ------------------------------------------------------------------------------------------------
// parity entry: one full round on caller-supplied (z, sdf)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kRayThreads) sampler_round_debug_kernel(SamplerDev S, const float* __restrict__ z, const float* __restrict__ sdf,
                                           long long R, int n, const float* __restrict__ beta_param,
                                           const float* __restrict__ beta_in, int upsample,
                                           const float* __restrict__ u_tape, float* out_beta, float* out_cdf,
                                           int* out_inds, float* out_samples, float* out_zm, int* out_src) {
    extern __shared__ __align__(16) float smem[];
    __shared__ float s_smp[128];
    const long long r = blockIdx.x;
    const float b0 = beta0_of(beta_param, S.beta_min);
    RayArrays A = carve_ray(smem);
    for (int i = threadIdx.x; i < n; i += kRayThreads) { A.z[i] = z[r * n + i]; A.s[i] = sdf[r * n + i]; }
    __syncthreads();
    compute_dstar(A, n);
    float beta = beta_search(A, n, b0, beta_in[r], S);
    if (threadIdx.x == 0 && out_beta) out_beta[r] = beta;
    const int ns = upsample ? S.n_eval : S.n_samples;
    const float* u = u_tape ? (u_tape + r * ns) : (upsample ? S.u_up : S.u_final);
    resample(A, n, beta, upsample != 0, S, u, ns, s_smp, out_inds ? out_inds + r * ns : nullptr);
    for (int j = threadIdx.x; j < ns; j += kRayThreads) out_samples[r * ns + j] = s_smp[j];
    if (out_cdf) for (int i = threadIdx.x; i < n; i += kRayThreads) out_cdf[r * n + i] = A.cdf[i];
    if (upsample && out_zm) merge_sorted(A.z, n, s_smp, ns, out_zm + r * (n + ns), out_src + r * (n + ns));
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static int ensure_smem_attrs() {
    static PerDeviceOnce once;
    if (!once.need()) return I2SDF_OK;
    I2SDF_CUDA_CHECK(cudaFuncSetAttribute(sampler_beta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSamplerSmem));
    I2SDF_CUDA_CHECK(cudaFuncSetAttribute(sampler_resample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSamplerSmem));
    I2SDF_CUDA_CHECK(cudaFuncSetAttribute(sampler_round_debug_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSamplerSmem));
    return I2SDF_OK;
}

static inline int ray_grid(long long R) { return (int)((R + kWarpsPerCta - 1) / kWarpsPerCta); }

int launch_rays(const float* uv, const float* pose, const float* intr, int B, int Pn, float* o, float* d, float* dnorm,
                cudaStream_t st) {
    long long n = (long long)B * Pn;
    if (n <= 0) return I2SDF_OK;
    rays_kernel<<<(int)((n + 255) / 256), 256, 0, st>>>(uv, pose, intr, B, Pn, o, d, dnorm);
    I2SDF_CUDA_CHECK(cudaGetLastError());
    return I2SDF_OK;
}

size_t sampler_ws_floats(const i2sdf_handle* h, long long R) {
    const int zmax = h->smp.n_eval * h->smp.max_iters;
    // z[2], sdf[2], src : 5 * R * zmax ; samples, sdf_new : 2 * R * n_eval ; beta : R ; beta_max : 8
    return (size_t)5 * R * zmax + (size_t)2 * R * h->smp.n_eval + (size_t)R + 8;
}

SamplerWs carve_sampler_ws(const i2sdf_handle* h, long long R, float* base) {
    SamplerWs W;
    W.zmax = h->smp.n_eval * h->smp.max_iters;
    size_t rz = (size_t)R * W.zmax;
    W.z[0] = base; W.z[1] = base + rz; W.sdf[0] = base + 2 * rz; W.sdf[1] = base + 3 * rz;
    W.src = reinterpret_cast<int*>(base + 4 * rz);
    W.samples = base + 5 * rz;
    W.sdf_new = W.samples + (size_t)R * h->smp.n_eval;
    W.beta = W.sdf_new + (size_t)R * h->smp.n_eval;
    W.beta_max = W.beta + R;
    return W;
}

int launch_sampler_init(const i2sdf_handle* h, const SamplerWs& W, long long R, const float* jitter, float coeff, cudaStream_t st) {
    sampler_init_kernel<<<ray_grid(R), kWarpsPerCta * 32, 0, st>>>(h->smp, W, R, jitter, coeff);
    I2SDF_CUDA_CHECK(cudaGetLastError());
    return I2SDF_OK;
}

int launch_sampler_round(const i2sdf_handle* h, const SamplerWs& W, long long R, int k, int phase, const float* beta_param,
                         const float* u_final_tape, cudaStream_t st) {
    int rc = ensure_smem_attrs();
    if (rc) return rc;
    if (R <= 0) return I2SDF_OK;
    // phase 0: round k (beta search + speculative up-sampling); phase 1: the final samples, k = -1: after whichever round was the last,
    // k >= 0: only if round k was the last (staged API: launched behind every round's convergence exchange)
    if (phase == 0) sampler_beta_kernel<<<(int)R, kRayThreads, kSamplerSmem, st>>>(h->smp, W, R, k, beta_param);
    else sampler_resample_kernel<<<(int)R, kRayThreads, kSamplerSmem, st>>>(h->smp, W, R, k, beta_param, u_final_tape);
    I2SDF_CUDA_CHECK(cudaGetLastError());
    return I2SDF_OK;
}

int launch_sampler_finalize(const i2sdf_handle* h, const SamplerWs& W, long long R, const float* beta_param,
                            const int* extra_tape, int extra_per_round, const int* eik_idx, float* out_z, float* out_z_eik, int* info, cudaStream_t st) {
    sampler_finalize_kernel<<<ray_grid(R), kWarpsPerCta * 32, 0, st>>>(h->smp, W, R, beta_param, extra_tape, extra_per_round, eik_idx, out_z, out_z_eik, info);
    I2SDF_CUDA_CHECK(cudaGetLastError());
    return I2SDF_OK;
}

int launch_sampler_round_debug(const i2sdf_handle* h, const float* z, const float* sdf, long long R, int n, const float* beta_param,
                               const float* beta_in, int upsample, const float* u_tape, float* out_beta, float* out_cdf,
                               int* out_inds, float* out_samples, float* out_zm, int* out_src, cudaStream_t st) {
    int rc = ensure_smem_attrs();
    if (rc) return rc;
    if (R <= 0) return I2SDF_OK;
    sampler_round_debug_kernel<<<(int)R, kRayThreads, kSamplerSmem, st>>>(h->smp, z, sdf, R, n, beta_param, beta_in,
        upsample, u_tape, out_beta, out_cdf, out_inds, out_samples, out_zm, out_src);
    I2SDF_CUDA_CHECK(cudaGetLastError());
    return I2SDF_OK;
}

int launch_composite(const i2sdf_handle* h, const float* z, const float* dnorm, const float* sdf, const float* rgb,
                     const float* grad, const float* lmask, const float* beta_param, long long R, int N, float* out_rgb,
                     float* out_depth, float* out_wsum, float* out_normal, float* out_light, float* out_w, cudaStream_t st) {
    if (N > 256) { set_error("composite: N=%d > 256 samples per ray unsupported", N); return I2SDF_E_INVALID; }
    CompositeArgs C;
    C.z = z; C.dnorm = dnorm; C.sdf = sdf; C.rgb = rgb; C.grad = grad; C.lmask = lmask; C.beta_param = beta_param;
    C.beta_min = h->smp.beta_min; C.out_rgb = out_rgb; C.out_depth = out_depth; C.out_wsum = out_wsum;
    C.out_normal = out_normal; C.out_light = out_light; C.out_w = out_w; C.R = R; C.N = N;
    composite_kernel<<<ray_grid(R), kWarpsPerCta * 32, 0, st>>>(C);
    I2SDF_CUDA_CHECK(cudaGetLastError());
    return I2SDF_OK;
}

}  // namespace i2sdf
