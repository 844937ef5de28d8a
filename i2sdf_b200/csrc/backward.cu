// Training backward of the per-ray path (fp32, FMA pipe): parameter gradients of the SDF / radiance / light-mask
// stacks and density.beta, including the SECOND-ORDER terms the eikonal and normal losses need.
//
// What autograd does in the reference (reference file:line):
//   * first order: backward through RenderingNetwork (mlp.py:208-229), the light head
//     (model/network/__init__.py:162-170), volume_rendering (:223-240) and ImplicitNetwork.forward (mlp.py:84-105);
//   * second order: grad_x sdf is built with autograd.grad(create_graph=True) (mlp.py:107-143) and feeds
//     grad_theta / diff_norm / normal_values (model/network/__init__.py:188-209), so backward differentiates the
//     reverse sweep again.
// Here the second-order part is evaluated as "tangent forward + joint reverse": for upstream gbar on g = grad_x sdf,
//   gbar . g(x; theta) = d/de sdf(x + e*gbar), so one forward tangent pass (adot_l) followed by a reverse pass that
//   carries the adjoints p_l (of a_l) and q_l (of adot_l) gives all parameter gradients:
//     q_{L-1} = e_sdf,  p_{L-1} = [sbar, fbar]
//     dW_l += p_l h~_l^T + q_l hdot~_l^T ,  db_l += p_l
//     u = W_l^T p_l , v = W_l^T q_l
//     q_{l-1} = s'(a_{l-1}) * v ,  p_{l-1} = s'(a_{l-1}) * u + s''(a_{l-1}) * adot_{l-1} * v
// All dense contractions go through one generic tiled SGEMM (NT / NN / TN split-K with atomics); activations are
// [M][256] fp32 arrays in the caller's workspace.  (The forward kernels are fused; this backward is the first
// correct version and is deliberately un-fused — DESIGN.md lists fusing it onto tcgen05 as the next step.)
#include <stdlib.h>
#include "common.cuh"
#include "tc_bwd.cuh"
#include "wgrad_planes.cuh"
#include "tc_common.cuh"

namespace i2sdf {

// tensor-core GEMMs (tc_gemm.cu)
size_t tc_wgrad_ws_floats(const i2sdf_handle* h);
int tc_gemm_pw(const i2sdf_handle* h, cudaStream_t st, long long M, const float* A, int lda, int kvalid, const TcBlock& blk, float* C, int ldc,
               int ncols, const float* bias, int relu);      // relu: bit 0 = ReLU on the output, bit 1 = ReLU on the input
int tc_gemm_wgrad(const i2sdf_handle* h, cudaStream_t st, long long M, const float* P0, int ldp0, const float* X0, int ldx0, const float* P1,
                  int ldp1, const float* X1, int ldx1, int n1, int n2, float* dW, int ldw, float* ws);
int tc_gemm_pw_ex(const i2sdf_handle* h, cudaStream_t st, long long M, const float* A, int lda, int kvalid, const TcBlock& blk, float* C, int ldc,
                  int ncols, const float* bias, int relu, const float* A2, int lda2, int is_skip, int nsplit, const float* E2);
int tc_gemm_wgrad_ex(const i2sdf_handle* h, cudaStream_t st, long long M, const float* P0, int ldp0, const float* X0, int ldx0, const float* P1,
                     int ldp1, const float* X1, int ldx1, int n1, int n2, float* dW, int ldw, float* ws, const float* Aprev, const float* Adprev,
                     int is_skip, int nsplit, const float* E0, const float* E1);

namespace bwd {

constexpr int LD = 256;
constexpr float SQRT2 = 1.41421356237309504880f;

// ------------------------------------------------------------------------------------------------
// generic SGEMM:  C[m][n] (+)= sum_k A(m,k) * B(k,n)
//   A(m,k) = TA ? A[k*lda + m] : A[m*lda + k] ;  B(k,n) = TB ? B[n*ldb + k] : B[k*ldb + n]
//   ATOMIC: grid.z splits K, results are atomically added into C (C must be pre-initialised)
// ------------------------------------------------------------------------------------------------
constexpr int BM = 128, BN = 128, BK = 16;

template <bool TA, bool TB, bool ATOMIC>
__global__ void __launch_bounds__(256) sgemm_kernel(int M, int N, int K, const float* __restrict__ A, int lda,
                                                    const float* __restrict__ B, int ldb, float* __restrict__ C, int ldc,
                                                    int kchunk, int accumulate) {
    __shared__ float As[BK][BM + 4];
    __shared__ float Bs[BK][BN + 4];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int kbeg = ATOMIC ? blockIdx.z * kchunk : 0;
    const int kend = ATOMIC ? min(K, kbeg + kchunk) : K;
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    for (int k0 = kbeg; k0 < kend; k0 += BK) {
        // A tile: BM x BK
#pragma unroll
        for (int it = 0; it < (BM * BK) / 256; ++it) {
            int idx = it * 256 + tid;
            int m, k;
            if (TA) { m = idx % BM; k = idx / BM; } else { k = idx % BK; m = idx / BK; }
            int gm = m0 + m, gk = k0 + k;
            float v = 0.f;
            if (gm < M && gk < kend) v = TA ? A[(size_t)gk * lda + gm] : A[(size_t)gm * lda + gk];
            As[k][m] = v;
        }
#pragma unroll
        for (int it = 0; it < (BN * BK) / 256; ++it) {
            int idx = it * 256 + tid;
            int n, k;
            if (TB) { k = idx % BK; n = idx / BK; } else { n = idx % BN; k = idx / BN; }
            int gn = n0 + n, gk = k0 + k;
            float v = 0.f;
            if (gn < N && gk < kend) v = TB ? B[(size_t)gn * ldb + gk] : B[(size_t)gk * ldb + gn];
            Bs[k][n] = v;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[8], b[8];
            float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 8]);
            float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * 8 + 4]);
            float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tx * 8]);
            float4 b1 = *reinterpret_cast<const float4*>(&Bs[k][tx * 8 + 4]);
            a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
            b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        int gm = m0 + ty * 8 + i;
        if (gm >= M) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            int gn = n0 + tx * 8 + j;
            if (gn >= N) continue;
            float* c = C + (size_t)gm * ldc + gn;
            if (ATOMIC) atomicAdd(c, acc[i][j]);
            else *c = accumulate ? (*c + acc[i][j]) : acc[i][j];
        }
    }
}

static int gemm_nt(cudaStream_t st, int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc) {
    dim3 g((N + BN - 1) / BN, (M + BM - 1) / BM, 1);
    sgemm_kernel<false, true, false><<<g, 256, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, K, 0);
    I2SDF_CUDA_CHECK(cudaGetLastError());
    return 0;
}
static int gemm_nn(cudaStream_t st, int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc) {
    dim3 g((N + BN - 1) / BN, (M + BM - 1) / BM, 1);
    sgemm_kernel<false, false, false><<<g, 256, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, K, 0);
    I2SDF_CUDA_CHECK(cudaGetLastError());
    return 0;
}
// C[N1][N2] += A[K][N1]^T B[K][N2]   (reduction over the K = points dimension, split across CTAs)
static int gemm_tn_acc(cudaStream_t st, int N1, int N2, long long K, const float* A, int lda, const float* B, int ldb, float* C, int ldc, int sms) {
    int tiles = ((N1 + BM - 1) / BM) * ((N2 + BN - 1) / BN);
    int splits = (int)max(1LL, min((long long)(4 * sms / max(tiles, 1)), (K + 511) / 512));
    int kchunk = (int)(((K + splits - 1) / splits + BK - 1) / BK * BK);
    splits = (int)((K + kchunk - 1) / kchunk);
    dim3 g((N2 + BN - 1) / BN, (N1 + BM - 1) / BM, splits);
    sgemm_kernel<true, false, true><<<g, 256, 0, st>>>(N1, N2, (int)K, A, lda, B, ldb, C, ldc, kchunk, 1);
    I2SDF_CUDA_CHECK(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------
// elementwise pieces
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float sp100(float a) { float t = a * 100.f; return t > 20.f ? a : __fdiv_rn(log1pf(expf(t)), 100.f); }
__device__ __forceinline__ float dsp100(float a) { float t = a * 100.f; if (t > 20.f) return 1.f; float z = expf(t); return __fdiv_rn(z, z + 1.f); }
// MUFU versions (ex2 / lg2 / rcp .approx.ftz, as in the tensor-core epilogues): softplus_100 and its derivative from ONE exp
__device__ __forceinline__ void sp100_fast(float a, float& sp, float& dsp) {
    const float e = tc::ex2_approx(-fabsf(a) * 144.26950408889634f);          // exp(-|100 a|)
    sp = fmaf(tc::lg2_approx(1.0f + e), 0.0069314718055994531f, fmaxf(a, 0.0f));
    const float rr = tc::rcp_approx(1.0f + e);
    dsp = (a >= 0.f) ? rr : e * rr;
}
__device__ __forceinline__ float d2sp100(float a) { float t = a * 100.f; if (t > 20.f) return 0.f; float z = expf(t); float q = z + 1.f; return 100.f * __fdiv_rn(z, q * q); }

struct PointSrc {             // explicit points, or rays: m -> (m / ns, m % ns)
    const float* pts; const float* o; const float* d; const float* z; int zstride; int ns;
};
__device__ __forceinline__ void load_point(const PointSrc& S, long long m, float (&x)[3]) {
    if (S.pts) { x[0] = S.pts[m * 3]; x[1] = S.pts[m * 3 + 1]; x[2] = S.pts[m * 3 + 2]; return; }
    long long r = m / S.ns; int j = (int)(m - r * S.ns);
    float t = S.z[r * S.zstride + j];
#pragma unroll
    for (int c = 0; c < 3; ++c) x[c] = __fadd_rn(S.o[r * 3 + c], __fmul_rn(t, S.d[r * 3 + c]));
}

// E [M][40]: embedding of x ;  Edot [M][40]: J(x) gbar   (either output may be null)
__global__ void embed_kernel(PointSrc S, long long M, int mx, const float* __restrict__ gbar, float* __restrict__ E, float* __restrict__ Ed) {
    long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    float x[3];
    load_point(S, m, x);
    float g[3] = {0.f, 0.f, 0.f};
    if (gbar) { g[0] = gbar[m * 3]; g[1] = gbar[m * 3 + 1]; g[2] = gbar[m * 3 + 2]; }
    float* e = E ? E + m * 40 : nullptr;
    float* ed = Ed ? Ed + m * 40 : nullptr;
    for (int c = 0; c < 3; ++c) { if (e) e[c] = x[c]; if (ed) ed[c] = g[c]; }
    for (int k = 0; k < mx; ++k) {
        float f = (float)(1 << k);
        for (int c = 0; c < 3; ++c) {
            float s, co;
            sincosf(__fmul_rn(x[c], f), &s, &co);
            if (e) { e[3 + 6 * k + c] = s; e[6 + 6 * k + c] = co; }
            if (ed) { ed[3 + 6 * k + c] = f * co * g[c]; ed[6 + 6 * k + c] = -f * s * g[c]; }
        }
    }
    for (int i = 3 + 6 * mx; i < 40; ++i) { if (e) e[i] = 0.f; if (ed) ed[i] = 0.f; }
}

// layer input from the previous layer's saved pre-activation:
//   X0 = softplus(A_prev)           (+ skip: [X0(:nsplit), E] / sqrt2)
//   X1 = softplus'(A_prev) * Adot   (+ skip: [X1(:nsplit), Edot] / sqrt2)        (X1/Adot optional)
__global__ void layer_input_kernel(long long M, const float* __restrict__ Aprev, const float* __restrict__ Adot, int is_skip, int nsplit,
                                   const float* __restrict__ E, const float* __restrict__ Ed, float* __restrict__ X0, float* __restrict__ X1) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M * LD) return;
    long long m = i / LD; int f = (int)(i % LD);
    float a = Aprev[i];
    float x0 = sp100(a);
    float x1 = (X1 && Adot) ? dsp100(a) * Adot[i] : 0.f;
    if (is_skip) {
        if (f >= nsplit) { x0 = E[m * 40 + (f - nsplit)]; x1 = (X1 && Ed) ? Ed[m * 40 + (f - nsplit)] : 0.f; }
        x0 = __fdiv_rn(x0, SQRT2);
        x1 = __fdiv_rn(x1, SQRT2);
    }
    if (X0) X0[i] = x0;
    if (X1) X1[i] = x1;
}

// adjoints through the activation of hidden layer l (width wo):
//   Q = s'(A) * V ;  P = s'(A) * U + s''(A) * Adot * V        (V / Adot null -> first-order only)
//   U, V may carry the 1/sqrt2 of a skip concat downstream (scale).  Columns >= wo are zeroed.
__global__ void act_adjoint_kernel(long long M, int wo, const float* __restrict__ A, const float* __restrict__ Adot, const float* __restrict__ U,
                                   const float* __restrict__ V, const float* __restrict__ Vrow, const float* __restrict__ sbar,
                                   const float* __restrict__ wsdf, int div_sqrt2, float* __restrict__ P, float* __restrict__ Q) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M * LD) return;
    long long m = i / LD; int f = (int)(i % LD);
    float p = 0.f, q = 0.f;
    if (f < wo) {
        float a = A[i];
        float u = U ? U[i] : 0.f;
        if (sbar) u = fmaf(sbar[m], wsdf[f], u);          // rank-1 part of W_last^T [sbar; fbar]
        float v = V ? V[i] : (Vrow ? Vrow[f] : 0.f);
        if (div_sqrt2) { u = __fdiv_rn(u, SQRT2); v = __fdiv_rn(v, SQRT2); }
        float d1 = dsp100(a);
        q = d1 * v;
        p = d1 * u;
        if (Adot && (V || Vrow)) p = fmaf(d2sp100(a) * Adot[i], v, p);
    }
    P[i] = p;
    if (Q) Q[i] = q;
}

// out[j] += sum_m w[m] * X[m][j]   (w null -> 1) ; j < width
__global__ void colsum_kernel(long long M, int width, const float* __restrict__ X, int ldx, const float* __restrict__ w, float* __restrict__ out, int rows_per_cta) {
    int j = threadIdx.x;
    long long r0 = (long long)blockIdx.x * rows_per_cta, r1 = min(M, r0 + rows_per_cta);
    if (j >= width) return;
    float s = 0.f;
    for (long long m = r0; m < r1; ++m) s = fmaf(w ? w[m] : 1.f, X[m * ldx + j], s);
    atomicAdd(out + j, s);
}
static int colsum(cudaStream_t st, long long M, int width, const float* X, int ldx, const float* w, float* out) {
    if (M <= 0) return 0;
    int rows = 256;
    int nthreads = (width + 31) / 32 * 32;
    colsum_kernel<<<(int)((M + rows - 1) / rows), nthreads, 0, st>>>(M, width, X, ldx, w, out, rows);
    I2SDF_CUDA_CHECK(cudaGetLastError());
    return 0;
}
__global__ void sum_kernel(long long M, const float* __restrict__ x, float* __restrict__ out) {
    float s = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < M; i += (long long)gridDim.x * blockDim.x) s += x[i];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, s);
}

static inline int blocks(long long n, int t = 256) { return (int)((n + t - 1) / t); }

}  // namespace bwd

// ================================================================================================
// SDF stack backward.  W[l]: effective weights row-major [out_l][in_l] (device pointers, host array).
// sbar [M] / fbar [M][ldf] / gbar [M][3] : upstream grads (any may be null).  dW[l] / db[l]: accumulated into.
// act: saved pre-activations [L-1][M][256].  ws: >= sdf_backward_ws_floats(M).
// ================================================================================================
size_t sdf_backward_ws_floats(const i2sdf_handle* h, long long M) {
    // Adot[L-1] + U,V,P,Q,X0,X1 (each M*256) + E,Ed (M*40 each)
    return (size_t)(h->net.L - 1 + 6) * M * 256 + (size_t)2 * M * 40 + 64 + (h->use_tc ? tc_wgrad_ws_floats(h) : 0);
}

int sdf_backward(const i2sdf_handle* h, const bwd::PointSrc& src, long long M, const float* const* W, const float* act,
                 const float* sbar, const float* fbar, int ldf, const float* gbar, float* const* dW, float* const* db, float* ws,
                 cudaStream_t st) {
    using namespace bwd;
    if (M <= 0) return I2SDF_OK;
    const NetDev& n = h->net;
    const int L = n.L, ex = n.ex, nsplit = 256 - ex;
    const size_t MB = (size_t)M * 256;
    float* Adot = ws;                       // [L-1][M][256]
    float* U = Adot + (size_t)(L - 1) * MB;
    float* V = U + MB;
    float* P = V + MB;
    float* Q = P + MB;
    float* X0 = Q + MB;
    float* X1 = X0 + MB;
    float* E = X1 + MB;                     // [M][40]
    float* Ed = E + (size_t)M * 40;
    float* WGP = Ed + (size_t)M * 40 + 16;  // per-CTA weight-gradient partials (tensor-core path)
    const bool tc = h->use_tc;
    static const bool fuse = !(getenv("I2SDF_BWD_FUSE") && getenv("I2SDF_BWD_FUSE")[0] == '0');
    const bool second = gbar != nullptr;
    auto wo = [&](int l) { return h->lay_out[l]; };     // 256, 217 (feeds skip) or 257 (last)
    auto wi = [&](int l) { return h->lay_in[l]; };      // 39 or 256
    int rc;
    embed_kernel<<<blocks(M), 256, 0, st>>>(src, M, n.mx, gbar, E, second ? Ed : nullptr);
    I2SDF_CUDA_CHECK(cudaGetLastError());
    // ---- tangent forward: Adot_l = Hdot~_l W_l^T
    if (second) {
        if (tc) rc = tc_gemm_pw(h, st, M, Ed, 40, ex, tc_block(h, TCB_FWD_SDF, 0), Adot, 256, 256, nullptr, 0);
        else rc = gemm_nt(st, (int)M, wo(0), ex, Ed, 40, W[0], wi(0), Adot, 256);
        if (rc) return rc;
        for (int l = 1; l < L - 1; ++l) {
            if (tc && fuse) {     // hdot~_l = softplus'(a_{l-1}) * adot_{l-1} (+ skip) is formed while staging the A operand
                rc = tc_gemm_pw_ex(h, st, M, act + (size_t)(l - 1) * MB, 256, 256, tc_block(h, TCB_FWD_SDF, l), Adot + (size_t)l * MB, 256, 256, nullptr, 0,
                                   Adot + (size_t)(l - 1) * MB, 256, l == n.skip, nsplit, Ed);
                if (rc) return rc;
                continue;
            }
            layer_input_kernel<<<blocks(MB), 256, 0, st>>>(M, act + (size_t)(l - 1) * MB, Adot + (size_t)(l - 1) * MB, l == n.skip, nsplit, E, Ed, nullptr, X1);
            I2SDF_CUDA_CHECK(cudaGetLastError());
            if (tc) rc = tc_gemm_pw(h, st, M, X1, 256, 256, tc_block(h, TCB_FWD_SDF, l), Adot + (size_t)l * MB, 256, 256, nullptr, 0);
            else rc = gemm_nt(st, (int)M, wo(l), 256, X1, 256, W[l], wi(l), Adot + (size_t)l * MB, 256);
            if (rc) return rc;
        }
    }
    // ---- last layer (index L-1): p = [sbar; fbar], q = e_sdf
    {
        const int l = L - 1;
        layer_input_kernel<<<blocks(MB), 256, 0, st>>>(M, act + (size_t)(l - 1) * MB, second ? Adot + (size_t)(l - 1) * MB : nullptr, l == n.skip,
                                                        nsplit, E, Ed, X0, second ? X1 : nullptr);
        I2SDF_CUDA_CHECK(cudaGetLastError());
        if (sbar) {
            if ((rc = colsum(st, M, 256, X0, 256, sbar, dW[l]))) return rc;                      // dW[0,:] += sum sbar * h~
            sum_kernel<<<64, 256, 0, st>>>(M, sbar, db[l]);
            I2SDF_CUDA_CHECK(cudaGetLastError());
        }
        if (second && (rc = colsum(st, M, 256, X1, 256, nullptr, dW[l]))) return rc;             // q = e_sdf: dW[0,:] += sum hdot~
        if (fbar) {
            if (tc) rc = tc_gemm_wgrad(h, st, M, fbar, ldf, X0, 256, nullptr, 0, nullptr, 0, 256, 256, dW[l] + 256, 256, WGP);
            else rc = gemm_tn_acc(st, 256, 256, M, fbar, ldf, X0, 256, dW[l] + 256, 256, h->num_sms);              // rows 1..256
            if (rc) return rc;
            if ((rc = colsum(st, M, 256, fbar, ldf, nullptr, db[l] + 1))) return rc;
            if (tc) rc = tc_gemm_pw(h, st, M, fbar, ldf, 256, tc_block(h, TCB_REV_FEAT, 0), U, 256, 256, nullptr, 0);
            else rc = gemm_nn(st, (int)M, 256, 256, fbar, ldf, W[l] + 256, 256, U, 256);                           // U = fbar W_feat
            if (rc) return rc;
        }
    }
    // ---- hidden layers, top down
    for (int l = L - 2; l >= 0; --l) {
        const bool top = (l == L - 2);
        const float* Al = act + (size_t)l * MB;
        const float* Adl = second ? Adot + (size_t)l * MB : nullptr;
        const bool from_skip = (l + 1 == n.skip);            // U,V come from the skip layer's input: scale by 1/sqrt2
        if (top) {
            act_adjoint_kernel<<<blocks(MB), 256, 0, st>>>(M, wo(l), Al, Adl, fbar ? U : nullptr, nullptr, second ? W[L - 1] : nullptr, sbar, W[L - 1],
                                                            from_skip, P, second ? Q : nullptr);
        } else {
            act_adjoint_kernel<<<blocks(MB), 256, 0, st>>>(M, wo(l), Al, Adl, U, second ? V : nullptr, nullptr, nullptr, nullptr, from_skip, P,
                                                            second ? Q : nullptr);
        }
        I2SDF_CUDA_CHECK(cudaGetLastError());
        // inputs of layer l
        const float* in0; const float* in1; int ldin;
        const bool fused_in = tc && fuse && l > 0;       // layer inputs are recomputed inside the weight-gradient loaders
        if (l == 0) { in0 = E; in1 = Ed; ldin = 40; }
        else {
            if (!fused_in) {
                layer_input_kernel<<<blocks(MB), 256, 0, st>>>(M, act + (size_t)(l - 1) * MB, second ? Adot + (size_t)(l - 1) * MB : nullptr, l == n.skip, nsplit,
                                                                E, Ed, X0, second ? X1 : nullptr);
                I2SDF_CUDA_CHECK(cudaGetLastError());
            }
            in0 = X0; in1 = X1; ldin = 256;
        }
        if (fused_in) {
            if ((rc = tc_gemm_wgrad_ex(h, st, M, P, 256, nullptr, 256, second ? Q : nullptr, 256, nullptr, 256, wo(l), wi(l), dW[l], wi(l), WGP,
                                       act + (size_t)(l - 1) * MB, second ? Adot + (size_t)(l - 1) * MB : nullptr, l == n.skip, nsplit, E, Ed))) return rc;
        } else if (tc) {
            if ((rc = tc_gemm_wgrad(h, st, M, P, 256, in0, ldin, second ? Q : nullptr, 256, second ? in1 : nullptr, ldin, wo(l), wi(l), dW[l], wi(l), WGP))) return rc;
        } else {
            if ((rc = gemm_tn_acc(st, wo(l), wi(l), M, P, 256, in0, ldin, dW[l], wi(l), h->num_sms))) return rc;
            if (second && (rc = gemm_tn_acc(st, wo(l), wi(l), M, Q, 256, in1, ldin, dW[l], wi(l), h->num_sms))) return rc;
        }
        if ((rc = colsum(st, M, wo(l), P, 256, nullptr, db[l]))) return rc;
        if (l > 0) {
            if (tc) {
                if ((rc = tc_gemm_pw(h, st, M, P, 256, 256, tc_block(h, TCB_REV_SDF, l), U, 256, 256, nullptr, 0))) return rc;
                if (second && (rc = tc_gemm_pw(h, st, M, Q, 256, 256, tc_block(h, TCB_REV_SDF, l), V, 256, 256, nullptr, 0))) return rc;
            } else {
                if ((rc = gemm_nn(st, (int)M, 256, wo(l), P, 256, W[l], wi(l), U, 256))) return rc;
                if (second && (rc = gemm_nn(st, (int)M, 256, wo(l), Q, 256, W[l], wi(l), V, 256))) return rc;
            }
        }
    }
    return I2SDF_OK;
}

// ================================================================================================
// radiance stack backward (recomputes the hidden activations from feat; mlp.py:208-229)
// ================================================================================================
namespace bwd {
// X [M][288] = [PE(dir) (ed) | feat (256) | 0]   (reference column order), or with feat_first: [feat (256) | PE(dir) | 0]
// (the column order of the packed tensor-core weight blocks)
__global__ void color_input_kernel(long long M, int ns, int md, int ed, const float* __restrict__ dirs, const float* __restrict__ feat, float* __restrict__ X,
                                   int feat_first) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M * 288) return;
    long long m = i / 288; int c = (int)(i % 288);
    const int pe0 = feat_first ? 256 : 0, f0 = feat_first ? 0 : ed;
    float v = 0.f;
    if (c >= pe0 && c < pe0 + ed) {
        const int cp = c - pe0;
        const float* d = dirs + (m / ns) * 3;
        if (cp < 3) v = d[cp];
        else { int q = cp - 3, k = q / 6, s = (q % 6) / 3, cc = q % 3; float arg = __fmul_rn(d[cc], (float)(1 << k)); v = s ? cosf(arg) : sinf(arg); }
    } else if (c >= f0 && c < f0 + 256) v = feat[m * 256 + (c - f0)];
    X[i] = v;
}
__global__ void bias_relu_kernel(long long M, const float* __restrict__ b, float* __restrict__ H) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M * LD) return;
    H[i] = fmaxf(H[i] + b[i % LD], 0.f);
}
__global__ void relu_mask_kernel(long long n, const float* __restrict__ H, float* __restrict__ G) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    G[i] = H[i] > 0.f ? G[i] : 0.f;
}
// delta [M][4] = grgb * rgb (1 - rgb)
__global__ void sigmoid_adjoint3_kernel(long long M, const float* __restrict__ rgb, const float* __restrict__ grgb, float* __restrict__ D, long long m_up = -1) {
    long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    const bool up = m_up < 0 || m < m_up;            // grgb covers the first m_up points only
    for (int c = 0; c < 3; ++c) { float y = rgb[m * 3 + c]; D[m * 4 + c] = up ? grgb[m * 3 + c] * y * (1.f - y) : 0.f; }
    D[m * 4 + 3] = 0.f;
}
}  // namespace bwd

size_t color_backward_ws_floats(const i2sdf_handle* h, long long M) {
    return (size_t)M * 288 * 2 + (size_t)(h->net.Lc - 1) * M * 256 + (size_t)M * 256 + (size_t)M * 4 + 64 + (h->use_tc ? tc_wgrad_ws_floats(h) : 0);
}

// gfeat_out: [M][288] buffer; the feature adjoint is columns ed..ed+255 (ld 288)
int color_backward(const i2sdf_handle* h, long long M, int ns, const float* dirs, const float* const* W, const float* const* b,
                   const float* feat, const float* rgb, const float* grgb, float* const* dW, float* const* db, float* ws, float** gfeat,
                   int* gfeat_ld, cudaStream_t st) {
    using namespace bwd;
    if (M <= 0) return I2SDF_OK;
    const NetDev& n = h->net;
    const int Lc = n.Lc, ed = n.ed, kin = 256 + ed;
    const size_t MB = (size_t)M * 256;
    float* X = ws;                              // [M][288]
    float* GX = X + (size_t)M * 288;            // [M][288] adjoint of X
    float* H = GX + (size_t)M * 288;            // [Lc-1][M][256]
    float* G = H + (size_t)(Lc - 1) * MB;       // [M][256] running adjoint
    float* D = G + MB;                          // [M][4]
    float* WGP = D + (size_t)M * 4 + 16;
    const bool tc = h->use_tc;
    int rc;
    color_input_kernel<<<blocks((long long)M * 288), 256, 0, st>>>(M, ns, n.md, ed, dirs, feat, X, tc ? 1 : 0);
    I2SDF_CUDA_CHECK(cudaGetLastError());
    for (int l = 0; l < Lc - 1; ++l) {          // recompute hidden activations
        const float* in = l == 0 ? X : H + (size_t)(l - 1) * MB;
        if (tc) {
            if ((rc = tc_gemm_pw(h, st, M, in, l == 0 ? 288 : 256, l == 0 ? 288 : 256, tc_block(h, TCB_FWD_COL, l), H + (size_t)l * MB, 256, 256, b[l], 1))) return rc;
        } else {
            if ((rc = gemm_nt(st, (int)M, 256, l == 0 ? kin : 256, in, l == 0 ? 288 : 256, W[l], l == 0 ? kin : 256, H + (size_t)l * MB, 256))) return rc;
            bias_relu_kernel<<<blocks(MB), 256, 0, st>>>(M, b[l], H + (size_t)l * MB);
            I2SDF_CUDA_CHECK(cudaGetLastError());
        }
    }
    sigmoid_adjoint3_kernel<<<blocks(M), 256, 0, st>>>(M, rgb, grgb, D);
    I2SDF_CUDA_CHECK(cudaGetLastError());
    const float* Hlast = H + (size_t)(Lc - 2) * MB;
    if ((rc = gemm_tn_acc(st, 3, 256, M, D, 4, Hlast, 256, dW[Lc - 1], 256, h->num_sms))) return rc;
    if ((rc = colsum(st, M, 3, D, 4, nullptr, db[Lc - 1]))) return rc;
    if ((rc = gemm_nn(st, (int)M, 256, 3, D, 4, W[Lc - 1], 256, G, 256))) return rc;
    for (int l = Lc - 2; l >= 0; --l) {
        relu_mask_kernel<<<blocks(MB), 256, 0, st>>>((long long)MB, H + (size_t)l * MB, G);
        I2SDF_CUDA_CHECK(cudaGetLastError());
        const float* in = l == 0 ? X : H + (size_t)(l - 1) * MB;
        const int K = l == 0 ? kin : 256, ldin = l == 0 ? 288 : 256;
        if (tc) {
            if (l == 0) {   // X is [feat | PE(dir)]: feature columns -> dW[:, ed:], embedding columns -> dW[:, :ed]
                if ((rc = tc_gemm_wgrad(h, st, M, G, 256, X, 288, nullptr, 0, nullptr, 0, 256, 256, dW[0] + ed, kin, WGP))) return rc;
                if ((rc = tc_gemm_wgrad(h, st, M, G, 256, X + 256, 288, nullptr, 0, nullptr, 0, 256, ed, dW[0], kin, WGP))) return rc;
            } else {
                if ((rc = tc_gemm_wgrad(h, st, M, G, 256, in, 256, nullptr, 0, nullptr, 0, 256, 256, dW[l], 256, WGP))) return rc;
            }
        } else {
            if ((rc = gemm_tn_acc(st, 256, K, M, G, 256, in, ldin, dW[l], K, h->num_sms))) return rc;
        }
        if ((rc = colsum(st, M, 256, G, 256, nullptr, db[l]))) return rc;
        if (l > 0) {
            // adjoint of H_{l-1}: G <- G W_l
            if (tc) rc = tc_gemm_pw(h, st, M, G, 256, 256, tc_block(h, TCB_REV_COL, l), GX, 256, 256, nullptr, 0);
            else rc = gemm_nn(st, (int)M, 256, 256, G, 256, W[l], 256, GX, 256);
            if (rc) return rc;
            I2SDF_CUDA_CHECK(cudaMemcpyAsync(G, GX, MB * sizeof(float), cudaMemcpyDeviceToDevice, st));
        } else {
            // adjoint of the stack input; only the feature columns are consumed downstream -> GX[:, ed:ed+256]
            if (tc) rc = tc_gemm_pw(h, st, M, G, 256, 256, tc_block(h, TCB_REV_COL, 0), GX + ed, 288, 256, nullptr, 0);
            else rc = gemm_nn(st, (int)M, kin, 256, G, 256, W[0], kin, GX, 288);
            if (rc) return rc;
        }
    }
    *gfeat = GX + ed;
    *gfeat_ld = 288;
    return I2SDF_OK;
}

// ================================================================================================
// light-mask head backward (features detached: only the head's own parameters get gradients)
// ================================================================================================
namespace bwd {
__global__ void relu_copy_kernel(long long n, const float* __restrict__ X, float* __restrict__ Y) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) Y[i] = fmaxf(X[i], 0.f);
}
// Ain [M][128] pre-activation (bias b0 added here unless null) -> Hs = softplus ; given glm, lm: delta1 [M], D0 = A [M][128]
template <bool FAST>
__global__ void light_mid_kernel(long long M, int lh, const float* __restrict__ b0, const float* __restrict__ w1, const float* __restrict__ lm,
                                 const float* __restrict__ glm, const float* Ain, float* A, float* __restrict__ Hs, float* __restrict__ d1) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M * lh) return;
    long long m = i / lh; int f = (int)(i % lh);
    float a = Ain[i] + (b0 ? b0[f] : 0.f);
    float y = lm[m];
    float delta = glm[m] * y * (1.f - y);
    float sp, dsp;
    if (FAST) sp100_fast(a, sp, dsp);
    else { sp = sp100(a); dsp = dsp100(a); }
    Hs[i] = sp;
    A[i] = delta * w1[f] * dsp;                // adjoint of the hidden pre-activation
    if (f == 0) d1[m] = delta;
}
}  // namespace bwd

size_t light_backward_ws_floats(const i2sdf_handle* h, long long M) {
    return (size_t)M * 256 + (size_t)2 * M * h->net.lh + (size_t)M + 64 + (h->use_tc ? tc_wgrad_ws_floats(h) : 0);
}

// hidden (optional, tensor-core path): the head's hidden pre-activations W0 relu(feat) + b0 [M][lh] as light_forward left
// them in the saved state; without it they are recomputed
int light_backward(const i2sdf_handle* h, long long M, const float* const* W, const float* const* b, const float* feat, const float* hidden,
                   const float* lm, const float* glm, float* const* dW, float* const* db, float* ws, cudaStream_t st) {
    using namespace bwd;
    if (M <= 0) return I2SDF_OK;
    const int lh = h->net.lh;
    float* LF = ws;                              // relu(feat) [M][256]
    float* A = LF + (size_t)M * 256;             // [M][lh]
    float* Hs = A + (size_t)M * lh;              // [M][lh]
    float* d1 = Hs + (size_t)M * lh;             // [M]
    float* WGP = d1 + (((size_t)M + 3) & ~(size_t)3) + 16;   // tensor-core weight-gradient partials (16-byte aligned: float4 stores)
    const TcBlock blk = h->use_tc ? tc_block(h, TCB_FWD_LIGHT, 0) : TcBlock{nullptr, 0, 0};
    const bool tc = blk.ptr != nullptr;
    int rc;
    // relu(feat) once (100 MB: it stays L2-resident for the weight-gradient pass below, which is why this copy is cheaper than
    // fusing the ReLU into that pass's HBM-latency-bound operand loader: measured 74 + 70 us vs 213 us)
    relu_copy_kernel<<<blocks((long long)M * 256), 256, 0, st>>>((long long)M * 256, feat, LF);
    I2SDF_CUDA_CHECK(cudaGetLastError());
    if (!tc) {
        if ((rc = gemm_nt(st, (int)M, lh, 256, LF, 256, W[0], 256, A, lh))) return rc;
    } else if (!hidden) {
        if ((rc = tc_gemm_pw(h, st, M, feat, 256, 256, blk, A, lh, lh, nullptr, /*relu: input*/ 2))) return rc;
    }
    // hidden already holds the bias; the recomputed products do not
    if (tc) light_mid_kernel<true><<<blocks((long long)M * lh), 256, 0, st>>>(M, lh, hidden ? nullptr : b[0], W[1], lm, glm, hidden ? hidden : A, A, Hs, d1);
    else light_mid_kernel<false><<<blocks((long long)M * lh), 256, 0, st>>>(M, lh, b[0], W[1], lm, glm, A, A, Hs, d1);
    I2SDF_CUDA_CHECK(cudaGetLastError());
    if ((rc = colsum(st, M, lh, Hs, lh, d1, dW[1]))) return rc;          // dW1[0,:] += sum delta * h
    sum_kernel<<<64, 256, 0, st>>>(M, d1, db[1]);
    I2SDF_CUDA_CHECK(cudaGetLastError());
    if (tc) rc = tc_gemm_wgrad(h, st, M, A, lh, LF, 256, nullptr, 0, nullptr, 0, lh, 256, dW[0], 256, WGP);      // dW0 += D0^T relu(feat)
    else rc = gemm_tn_acc(st, lh, 256, M, A, lh, LF, 256, dW[0], 256, h->num_sms);
    if (rc) return rc;
    if ((rc = colsum(st, M, lh, A, lh, nullptr, db[0]))) return rc;
    return I2SDF_OK;
}

// ================================================================================================
// light-mask head forward behind the tensor-core main pass (model/network/__init__.py:162-170):
//   lmask = sigmoid(w1 . softplus_100(W0 relu(feat) + b0) + b1)
// The main-pass chain kernel has no room for a second 256-wide A operand (relu(feat) next to feat), so the head runs as
// its own HBM-bound pass over the features that kernel writes: one tcgen05 GEMM (ReLU fused into the operand staging)
// + one warp-per-point epilogue kernel.  32 896 of the path's ~1.0 M MACs per point.
// ================================================================================================
namespace bwd {
// A [M][lh] = W0 relu(feat) + b0  ->  out[m] = sigmoid(w1 . softplus(A[m]) + b1);  head = [w1 (lh) | b1];  lh <= 128, lh % 4 == 0
__global__ void __launch_bounds__(256) light_out_kernel(long long M, int lh, const float* __restrict__ A, const float* __restrict__ head,
                                                        float* __restrict__ out) {
    const long long m = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (m >= M) return;
    float acc = 0.f;
    if (lane * 4 < lh) {
        const float4 a = *reinterpret_cast<const float4*>(A + (size_t)m * lh + lane * 4);
        const float4 w = __ldg(reinterpret_cast<const float4*>(head) + lane);
        float s0, s1, s2, s3, dd;
        sp100_fast(a.x, s0, dd); sp100_fast(a.y, s1, dd); sp100_fast(a.z, s2, dd); sp100_fast(a.w, s3, dd);
        acc = fmaf(s0, w.x, fmaf(s1, w.y, fmaf(s2, w.z, s3 * w.w)));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) out[m] = __fdiv_rn(1.0f, 1.0f + expf(-(acc + __ldg(head + lh))));
}
}  // namespace bwd

size_t light_forward_ws_floats(const i2sdf_handle* h, long long M) { return (size_t)M * h->net.lh + 64; }

int light_forward(const i2sdf_handle* h, long long M, const float* feat, float* out_light, float* ws, cudaStream_t st) {
    using namespace bwd;
    if (M <= 0) return I2SDF_OK;
    const NetDev& n = h->net;
    const TcBlock blk = h->use_tc ? tc_block(h, TCB_FWD_LIGHT, 0) : TcBlock{nullptr, 0, 0};
    if (!blk.ptr || n.lh > 128 || (n.lh & 3)) { set_error("light_forward: tensor-core light block unavailable"); return I2SDF_E_INVALID; }
    int rc;
    if ((rc = tc_gemm_pw(h, st, M, feat, 256, 256, blk, ws, n.lh, n.lh, n.light_b0, /*relu: input*/ 2))) return rc;
    light_out_kernel<<<blocks(M * 32), 256, 0, st>>>(M, n.lh, ws, n.light_head, out_light);
    I2SDF_CUDA_CHECK(cudaGetLastError());
    return I2SDF_OK;
}

// ================================================================================================
// compositing backward  (volume_rendering + reductions, model/network/__init__.py:118-125,162-170,204-209,223-240)
// ================================================================================================
namespace bwd {
__device__ __forceinline__ double wscan_excl(double v, int lane, double* total) {
    double x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { double y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    *total = __shfl_sync(0xffffffffu, x, 31);
    return x - v;
}
__device__ __forceinline__ float wsum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

struct CompBwdArgs {
    const float* z; const float* dnorm; const float* sdf; const float* rgb; const float* grad; const float* lmask; const float* beta_param;
    float beta_min;
    const float* g_rgb; const float* g_depth; const float* g_wsum; const float* g_normal; const float* g_light;   // upstream (null = 0)
    float* o_sdf; float* o_rgb; float* o_grad; float* o_lmask; float* o_beta;
    long long R; int N;
};

__global__ void composite_bwd_kernel(CompBwdArgs C) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long r = (long long)blockIdx.x * 4 + warp;
    if (r >= C.R) return;
    const int N = C.N;
    const float braw = *C.beta_param;
    const float beta = fabsf(braw) + C.beta_min;
    const float alpha = 1.0f / beta;
    const float* z = C.z + r * (N + 1);
    const int per = (N + 31) / 32;
    const int i0 = lane * per, i1 = min(i0 + per, N);
    float fe[8], sig[8], ex[8], dist[8];
    double sum = 0.0;
    for (int i = i0, q = 0; i < i1; ++i, ++q) {
        float s = C.sdf[r * N + i];
        float sg = (s > 0.f) ? 1.f : ((s < 0.f) ? -1.f : 0.f);
        ex[q] = expf(-fabsf(s) / beta);                               // exp(-|s|/beta)
        sig[q] = alpha * (0.5f + 0.5f * sg * (ex[q] - 1.0f));
        dist[q] = z[i + 1] - z[i];
        fe[q] = dist[q] * sig[q];
        sum += (double)fe[q];
    }
    double tot;
    double acc = wscan_excl(sum, lane, &tot);
    // forward quantities per sample + first pass: normal accumulation (weights detached)
    float T[8], w[8];
    float vn[3] = {0.f, 0.f, 0.f};
    for (int i = i0, q = 0; i < i1; ++i, ++q) {
        T[q] = expf(-(float)acc);
        acc += (double)fe[q];
        w[q] = (1.0f - expf(-fe[q])) * T[q];
        if (C.g_normal) {
            const float* g = C.grad + (r * N + i) * 3;
            float nn = fmaxf(sqrtf(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]), 1e-12f);
            vn[0] += w[q] * (g[0] / nn); vn[1] += w[q] * (g[1] / nn); vn[2] += w[q] * (g[2] / nn);
        }
    }
    float dv[3] = {0.f, 0.f, 0.f};
    if (C.g_normal) {       // out = v/|v| ; dv = (gout - out (out.gout)) / |v|
        for (int c = 0; c < 3; ++c) vn[c] = wsum(vn[c]);
        float nv = fmaxf(sqrtf(vn[0] * vn[0] + vn[1] * vn[1] + vn[2] * vn[2]), 1e-12f);
        float o[3] = {vn[0] / nv, vn[1] / nv, vn[2] / nv};
        const float* go = C.g_normal + r * 3;
        float dot = o[0] * go[0] + o[1] * go[1] + o[2] * go[2];
        for (int c = 0; c < 3; ++c) dv[c] = (go[c] - o[c] * dot) / nv;
    }
    const float gd = C.g_depth ? C.g_depth[r] / fmaxf(C.dnorm[r], 1e-6f) : 0.f;
    const float gw = C.g_wsum ? C.g_wsum[r] : 0.f;
    const float gl = C.g_light ? C.g_light[r] : 0.f;
    float gc[3] = {0.f, 0.f, 0.f};
    if (C.g_rgb) { gc[0] = C.g_rgb[r * 3]; gc[1] = C.g_rgb[r * 3 + 1]; gc[2] = C.g_rgb[r * 3 + 2]; }
    // dL/dw_i and the suffix sums S_k = sum_{i>k} gw_i w_i
    float gwi[8];
    double part = 0.0;
    for (int i = i0, q = 0; i < i1; ++i, ++q) {
        float g = gw + gd * z[i];
        if (C.g_rgb) { const float* c = C.rgb + (r * N + i) * 3; g += gc[0] * c[0] + gc[1] * c[1] + gc[2] * c[2]; }
        gwi[q] = g;
        part += (double)(g * w[q]);
    }
    double total;
    double before = wscan_excl(part, lane, &total);     // sum over lanes < lane
    double run = before;                                // inclusive prefix up to current element grows below
    float gbeta = 0.f;
    for (int i = i0, q = 0; i < i1; ++i, ++q) {
        run += (double)(gwi[q] * w[q]);
        float suffix = (float)(total - run);            // sum_{j>i} gw_j w_j
        float gfe = gwi[q] * expf(-fe[q]) * T[q] - suffix;
        float gsig = gfe * dist[q];
        // sigma = alpha (0.5 + 0.5 sg (ex - 1)) ; d sigma/d s = -0.5 alpha^2 ex ; d sigma/d beta = -sigma/beta + 0.5 alpha sg ex |s| / beta^2
        float s = C.sdf[r * N + i];
        float sg = (s > 0.f) ? 1.f : ((s < 0.f) ? -1.f : 0.f);
        float dsig_ds = -0.5f * alpha * alpha * ex[q] * (sg != 0.f ? 1.f : 0.f);
        float dsig_db = -sig[q] * alpha + 0.5f * alpha * sg * ex[q] * fabsf(s) * alpha * alpha;
        C.o_sdf[r * N + i] = gsig * dsig_ds;
        gbeta += gsig * dsig_db;
        if (C.o_rgb) { float* o = C.o_rgb + (r * N + i) * 3; o[0] = w[q] * gc[0]; o[1] = w[q] * gc[1]; o[2] = w[q] * gc[2]; }
        if (C.o_lmask) C.o_lmask[r * N + i] = w[q] * gl;
        if (C.o_grad) {
            float* og = C.o_grad + (r * N + i) * 3;
            if (C.g_normal) {
                const float* g = C.grad + (r * N + i) * 3;
                float nrm = sqrtf(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]);
                float nn = fmaxf(nrm, 1e-12f);
                float nh[3] = {g[0] / nn, g[1] / nn, g[2] / nn};
                float dn[3] = {w[q] * dv[0], w[q] * dv[1], w[q] * dv[2]};
                float dot = nh[0] * dn[0] + nh[1] * dn[1] + nh[2] * dn[2];
                if (nrm > 1e-12f) { og[0] = (dn[0] - nh[0] * dot) / nn; og[1] = (dn[1] - nh[1] * dot) / nn; og[2] = (dn[2] - nh[2] * dot) / nn; }
                else { og[0] = dn[0] / nn; og[1] = dn[1] / nn; og[2] = dn[2] / nn; }
            } else { og[0] = 0.f; og[1] = 0.f; og[2] = 0.f; }
        }
    }
    gbeta = wsum(gbeta);
    if (lane == 0 && C.o_beta) atomicAdd(C.o_beta, gbeta * ((braw > 0.f) ? 1.f : ((braw < 0.f) ? -1.f : 0.f)));
}
}  // namespace bwd

int launch_composite_backward(const i2sdf_handle* h, const float* z, const float* dnorm, const float* sdf, const float* rgb, const float* grad,
                              const float* lmask, const float* beta_param, long long R, int N, const float* g_rgb, const float* g_depth,
                              const float* g_wsum, const float* g_normal, const float* g_light, float* o_sdf, float* o_rgb, float* o_grad,
                              float* o_lmask, float* o_beta, cudaStream_t st) {
    if (N > 256) { set_error("composite_backward: N > 256"); return I2SDF_E_INVALID; }
    if (R <= 0) return I2SDF_OK;
    bwd::CompBwdArgs C;
    C.z = z; C.dnorm = dnorm; C.sdf = sdf; C.rgb = rgb; C.grad = grad; C.lmask = lmask; C.beta_param = beta_param; C.beta_min = h->smp.beta_min;
    C.g_rgb = g_rgb; C.g_depth = g_depth; C.g_wsum = g_wsum; C.g_normal = g_normal; C.g_light = g_light;
    C.o_sdf = o_sdf; C.o_rgb = o_rgb; C.o_grad = o_grad; C.o_lmask = o_lmask; C.o_beta = o_beta; C.R = R; C.N = N;
    bwd::composite_bwd_kernel<<<(int)((R + 3) / 4), 128, 0, st>>>(C);
    I2SDF_CUDA_CHECK(cudaGetLastError());
    return I2SDF_OK;
}

// ================================================================================================
// Fused training backward on plane slots (tensor-core path): one chain kernel (mlp_tc_bwd.cu) for every per-point
// product, one weight-gradient launch for all layers (wgrad_planes.cu), one weighted column-sum launch for the rank-1
// pieces (sdf row of the last SDF layer, rgb head), and two tiny reductions.
// saved: forward state written by the tensor-core main pass in save mode (planes::Layout); ws: backward workspace.
// dW / db (SDF stack) and dWc / dbc (radiance stack; null without g_rgb) are accumulated into.
// ================================================================================================
size_t fused_backward_ws_bytes(const i2sdf_handle* h, long long M, bool color) {
    planes::Layout SL = planes::make_layout(M, h->net.L - 1, h->net.Lc, color, nullptr, nullptr);
    return SL.bwd_total() + ((size_t)M * 4 + 64) * sizeof(float);
}

int fused_backward(const i2sdf_handle* h, const bwd::PointSrc& src, long long M, void* saved, const float* s_rgb, const float* g_sdf,
                   const float* g_grad, const float* g_rgb, float* const* dW, float* const* db, float* const* dWc, float* const* dbc, void* ws,
                   int phases, long long m_rays, cudaStream_t st, long long m_up, const float* g_grad_tail) {
    if (m_up < 0 || m_up > M) m_up = M;
    // phases (bit mask, for per-phase timing by the caller): 1 = chain kernel, 2 = weight gradients, 4 = rank-1 pieces
    using namespace bwd;
    if (M <= 0) return I2SDF_OK;
    const NetDev& n = h->net;
    const int L = n.L, NL = L - 1, Lc = n.Lc;
    const bool color = g_rgb != nullptr;
    planes::Layout SL = planes::make_layout(M, NL, Lc, color, saved, ws);
    float* D = reinterpret_cast<float*>((uint8_t*)ws + SL.bwd_total());          // [M][4] rgb pre-sigmoid adjoints
    int rc;
    BwdParams p{};
    p.pts = src.pts; p.ray_o = src.o; p.ray_d = src.d; p.zarr = src.z; p.zstride = src.zstride; p.ns = src.ns; p.M = M;
    p.m_rays = (src.pts && src.o) ? m_rays : 0;
    p.g_sdf = g_sdf; p.g_grad = g_grad; p.g_rgb = g_rgb; p.s_rgb = s_rgb; p.with_color = color ? 1 : 0;
    p.m_up = m_up; p.g_grad_tail = g_grad_tail;
    const bool any_grad = g_grad != nullptr || g_grad_tail != nullptr;
    p.sl = SL; p.net = n;
    if ((phases & 1) && (rc = tc_bwd_launch(h, p, st))) return rc;
    if (!(phases & 6)) return I2SDF_OK;

    // ---- weight (+ bias) gradients: every dense product with the point index as reduction dimension
    WgArgs a{};
    a.ntiles = planes::ntiles(M);
    auto job = [&](float* dst, int ld, int rows, int cols) { a.jobs[a.njobs] = WgJob{dst, ld, rows, cols}; return a.njobs++; };
    auto term = [&](int j, size_t Poff, bool Pws, size_t Xoff, bool Xws, int xchunks, float* colsum, int ncs) {
        // P operands in the workspace are adjoint slots (P, PC, FB: HI plane only); X operands in the workspace are tangent slots (HD, ED)
        const int pp = Pws ? planes::kPlanesAdj : 2;
        const int xp = (Xws && xchunks == planes::BIG_CHUNKS) ? planes::kPlanesHD : 2;
        a.terms[a.nterms++] = WgTerm{(Pws ? SL.wbase : SL.base) + Poff, (Xws ? SL.wbase : SL.base) + Xoff, j, xchunks, colsum, ncs, pp, xp};
    };
    for (int l = NL - 1; l >= 1; --l) {
        const int j = job(dW[l], h->lay_in[l], h->lay_out[l], h->lay_in[l]);
        term(j, SL.P(l), true, SL.H(l - 1), false, planes::BIG_CHUNKS, db[l], h->lay_out[l]);
        if (any_grad) term(j, SL.Q(l), false, SL.HD(l - 1), true, planes::BIG_CHUNKS, nullptr, 0);
    }
    {
        const int j = job(dW[0], h->lay_in[0], h->lay_out[0], h->lay_in[0]);
        term(j, SL.P(0), true, SL.E(), false, planes::SMALL_CHUNKS, db[0], h->lay_out[0]);
        if (any_grad) term(j, SL.Q(0), false, SL.ED(), true, planes::SMALL_CHUNKS, nullptr, 0);
    }
    if (color) {
        int j = job(dW[L - 1] + h->lay_in[L - 1], h->lay_in[L - 1], 256, 256);            // feature rows 1..256 of the last SDF layer
        term(j, SL.FB(), true, SL.H(NL - 1), false, planes::BIG_CHUNKS, db[L - 1] + 1, 256);
        for (int l = Lc - 2; l >= 1; --l) {
            j = job(dWc[l], 256, 256, 256);
            term(j, SL.PC(l), true, SL.C(l - 1), false, planes::BIG_CHUNKS, dbc[l], 256);
        }
        const int kin = h->lay_in[L];                                                   // ed + 256, reference column order [PE(dir) | feat]
        j = job(dWc[0] + n.ed, kin, 256, 256);
        term(j, SL.PC(0), true, SL.CF(), false, planes::BIG_CHUNKS, dbc[0], 256);
        j = job(dWc[0], kin, 256, n.ed);
        term(j, SL.PC(0), true, SL.DV(), false, planes::DV_CHUNKS, nullptr, 0);
    }
    if ((phases & 2) && (rc = wgrad_planes_launch(h, a, st))) return rc;
    if (!(phases & 4)) return I2SDF_OK;

    // ---- rank-1 pieces
    CsArgs c{};
    c.ntiles = a.ntiles; c.M = M;
    if (g_sdf) c.jobs[c.njobs++] = CsJob{SL.base + SL.H(NL - 1), g_sdf, 1, dW[L - 1], 256, 1, 0, 2, m_up};      // dW_last[0,:] += sum sbar h~
    if (any_grad) c.jobs[c.njobs++] = CsJob{SL.wbase + SL.HD(NL - 1), nullptr, 0, dW[L - 1], 256, 1, 0, planes::kPlanesHD};        // q_last = e_sdf: += sum hdot~
    if (color) {
        sigmoid_adjoint3_kernel<<<blocks(M), 256, 0, st>>>(M, s_rgb, g_rgb, D, m_up);
        I2SDF_CUDA_CHECK(cudaGetLastError());
        c.jobs[c.njobs++] = CsJob{SL.base + SL.C(Lc - 2), D, 4, dWc[Lc - 1], 256, 3, 256, 2};      // the three rgb-head rows: one pass over the slot
        if ((rc = colsum(st, M, 3, D, 4, nullptr, dbc[Lc - 1]))) return rc;
    }
    if ((rc = planes_colsum_launch(h, c, st))) return rc;
    if (g_sdf) {
        sum_kernel<<<64, 256, 0, st>>>(m_up, g_sdf, db[L - 1]);
        I2SDF_CUDA_CHECK(cudaGetLastError());
    }
    return I2SDF_OK;
}

}  // namespace i2sdf
