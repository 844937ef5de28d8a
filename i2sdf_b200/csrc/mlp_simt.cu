// Fused per-tile MLP chain in fp32 on the FMA pipe (sm_100a).
//
// One CTA keeps a tile of 64 ray points resident in shared memory (feature-major, [feature][point]) and walks
// the whole chain on it: positional encoding -> SDF stack (softplus_100, skip concat) -> sdf head ->
// [feature layer -> light-mask head -> radiance stack -> rgb head] -> [reverse sweep for d sdf/dx].
// Weights stream from L2 through a 3-stage cp.async ring; only per-point results go back to HBM.
// This is the exact-fp32 path: the parity baseline for the tcgen05 kernel (mlp_tc.cu) and the executor of the
// odd-shaped pieces.
//
// Replaces (reference file:line): Embedder.embed embedder.py:28-38; ImplicitNetwork.forward mlp.py:84-105,
// get_sdf_vals :145-151, get_outputs/gradient :107-143; RenderingNetwork.forward mlp.py:208-229;
// light head model/network/__init__.py:162-168.
#include "common.cuh"

namespace i2sdf {
namespace simt {

constexpr int TM = 64;            // points per tile
constexpr int HS = TM + 4;        // shared row stride (floats): conflict-free float4 column writes
constexpr int NT = 256;           // threads per CTA: 8 warps; warp = 8 points, lane = feature (+32*j)
constexpr int KC = 4;             // weight rows (k) per cp.async chunk
constexpr int NSTAGE = 3;
constexpr int HROWS = 296;        // 256 features + 27 dir-embedding rows (+pad); rows 256.. double as skip adjoint
constexpr int EROWS = 40;         // embedding rows (39 + pad)
constexpr int WBUF = KC * 256;    // floats per stage

constexpr size_t kSmemBytes = (size_t)(HROWS * HS + EROWS * HS + NSTAGE * WBUF + 3 * TM + 3 * TM) * sizeof(float);

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// nn.Softplus(beta=100, threshold=20): x if 100x > 20 else log1p(exp(100x))/100      (mlp.py:76)
__device__ __forceinline__ float softplus100(float a) {
    float t = a * 100.0f;
    return t > 20.0f ? a : __fdiv_rn(log1pf(expf(t)), 100.0f);
}
// its derivative as autograd computes it: z/(z+1), z = exp(100x); 1 above the threshold
__device__ __forceinline__ float dsoftplus100(float a) {
    float t = a * 100.0f;
    if (t > 20.0f) return 1.0f;
    float z = expf(t);
    return __fdiv_rn(z, z + 1.0f);
}
__device__ __forceinline__ float sigmoidf_(float x) { return __fdiv_rn(1.0f, 1.0f + expf(-x)); }

// acc[p][j] += sum_k H[k][ty*8+p] * W[k][tx+32j]   for k in [0,K), K % KC == 0.
// H: shared, feature-major rows of stride HS.  Wg: global [K][NJ*32] row-major.  Wb: NSTAGE*WBUF floats.
// Ends WITHOUT a barrier: callers __syncthreads() before overwriting H.
template <int NJ, bool RELU_IN>
__device__ __forceinline__ void tile_gemm(const float* __restrict__ H, const float* __restrict__ Wg, int K,
                                          float* __restrict__ Wb, float (&acc)[8][NJ]) {
    constexpr int LDW = NJ * 32;
    constexpr int F4_PER_CHUNK = KC * LDW / 4;
    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    const int nchunks = K / KC;
    auto issue = [&](int c) {
        if (c < nchunks) {
            const float* src = Wg + (size_t)c * KC * LDW;
            float* dst = Wb + (c % NSTAGE) * WBUF;
            for (int i = tid; i < F4_PER_CHUNK; i += NT) cp_async16(dst + i * 4, src + i * 4);
        }
        cp_async_commit();
    };
    issue(0);
    issue(1);
    for (int c = 0; c < nchunks; ++c) {
        cp_async_wait<1>();
        __syncthreads();                 // chunk c visible to all; everyone is done with chunk c-1
        issue(c + 2);                    // refills the stage chunk c-1 used
        const float* w = Wb + (c % NSTAGE) * WBUF;
        const float* hrow = H + (size_t)c * KC * HS + ty * 8;
#pragma unroll
        for (int kk = 0; kk < KC; ++kk) {
            float4 h0 = *reinterpret_cast<const float4*>(hrow + kk * HS);
            float4 h1 = *reinterpret_cast<const float4*>(hrow + kk * HS + 4);
            float hv[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
            if (RELU_IN) {
#pragma unroll
                for (int p = 0; p < 8; ++p) hv[p] = fmaxf(hv[p], 0.0f);
            }
            float wv[NJ];
#pragma unroll
            for (int j = 0; j < NJ; ++j) wv[j] = w[kk * LDW + tx + 32 * j];
#pragma unroll
            for (int p = 0; p < 8; ++p)
#pragma unroll
                for (int j = 0; j < NJ; ++j) acc[p][j] = fmaf(hv[p], wv[j], acc[p][j]);
        }
    }
    cp_async_wait<0>();
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ bool round_active(const MlpParams& P) {
    if (P.round_idx <= 0) return true;
    float b0 = fabsf(*P.beta_param) + P.beta_min;
    for (int j = 0; j < P.round_idx; ++j)
        if (!(P.beta_max[j] > b0)) return false;
    return true;
}

__global__ void __launch_bounds__(NT, 2) mlp_tile_kernel(const MlpParams P) {
    extern __shared__ __align__(16) float smem[];
    float* Hs = smem;                         // [HROWS][HS]
    float* Es = Hs + HROWS * HS;              // [EROWS][HS] embedding of the points
    float* Wb = Es + EROWS * HS;              // weight ring
    float* Xs = Wb + NSTAGE * WBUF;           // [3][TM] point coordinates
    float* Ds = Xs + 3 * TM;                  // [3][TM] view directions
    float* REs = Hs + 256 * HS;               // [EROWS][HS] skip adjoint (aliases the dir-embedding rows)

    if (!round_active(P)) return;

    const NetDev& net = P.net;
    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    const int L = net.L;
    const bool full = (P.out_feat != nullptr) || (P.out_grad != nullptr) || P.want_color || P.want_light;
    const bool want_grad = P.out_grad != nullptr;
    const bool keep_act = want_grad || (P.save_act != nullptr);
    const int nsplit = 256 - net.ex;          // width produced by the layer feeding the skip concat (217)
    const long long ntiles = (P.M + TM - 1) / TM;
    const float SQRT2 = 1.41421356237309504880f;

    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long m0 = tile * TM;
        // activation store for this tile: a[l][p][f]
        float* Abase;
        long long Apstride;   // stride between layers in floats
        if (P.save_act) { Abase = P.save_act + m0 * 256; Apstride = P.M * 256; }
        else { Abase = P.scratch + (size_t)blockIdx.x * (size_t)(L - 1) * TM * 256; Apstride = (long long)TM * 256; }

        __syncthreads();   // previous tile's readers of Xs/Es/Hs are done
        // ---- 1. points (+ view dirs)
        if (tid < TM) {
            long long m = m0 + tid;
            float x = 0.f, y = 0.f, z = 0.f, dx = 0.f, dy = 0.f, dz = 1.f;
            if (m < P.M) {
                if (P.pts) {
                    x = P.pts[m * 3 + 0]; y = P.pts[m * 3 + 1]; z = P.pts[m * 3 + 2];
                    if (P.ray_d) { long long r = m / P.ns; dx = P.ray_d[r * 3]; dy = P.ray_d[r * 3 + 1]; dz = P.ray_d[r * 3 + 2]; }
                } else {
                    long long r = m / P.ns;
                    int j = (int)(m - r * P.ns);
                    float t = P.zarr[r * P.zstride + j];
                    dx = P.ray_d[r * 3]; dy = P.ray_d[r * 3 + 1]; dz = P.ray_d[r * 3 + 2];
                    // o + z*d, mul then add as the reference does (no fma)   network/__init__.py:103
                    x = __fadd_rn(P.ray_o[r * 3 + 0], __fmul_rn(t, dx));
                    y = __fadd_rn(P.ray_o[r * 3 + 1], __fmul_rn(t, dy));
                    z = __fadd_rn(P.ray_o[r * 3 + 2], __fmul_rn(t, dz));
                }
            }
            Xs[tid] = x; Xs[TM + tid] = y; Xs[2 * TM + tid] = z;
            Ds[tid] = dx; Ds[TM + tid] = dy; Ds[2 * TM + tid] = dz;
        }
        __syncthreads();
        // ---- 2. embedding rows [x, sin(2^k x), cos(2^k x)]   embedder.py:28-38
        for (int i = tid; i < EROWS * TM; i += NT) {
            int r = i / TM, p = i - r * TM;
            float v = 0.f;
            if (r < 3) v = Xs[r * TM + p];
            else if (r < net.ex) {
                int q = r - 3, k = q / 6, s = (q % 6) / 3, c = q % 3;
                float arg = __fmul_rn(Xs[c * TM + p], (float)(1 << k));
                v = s ? cosf(arg) : sinf(arg);
            }
            Es[r * HS + p] = v;
        }
        __syncthreads();

        // ---- 3. SDF hidden layers
        float head[8];
#pragma unroll
        for (int p = 0; p < 8; ++p) head[p] = 0.f;
        for (int l = 0; l < L - 1; ++l) {
            float acc[8][8];
#pragma unroll
            for (int p = 0; p < 8; ++p)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[p][j] = 0.f;
            tile_gemm<8, false>(l == 0 ? Es : Hs, net.sdf_wt[l], net.sdf_kpad[l], Wb, acc);
            __syncthreads();
            const float* bias = net.sdf_b[l];
            const bool feeds_skip = (l + 1 == net.skip);
            const bool last_hidden = (l == L - 2);
            const bool write_h = !(last_hidden && !full);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int f = tx + 32 * j;
                const float b = bias[f];
                const float hw = last_hidden ? net.sdf_head[f] : 0.f;
#pragma unroll
                for (int p = 0; p < 8; ++p) {
                    const int pp = ty * 8 + p;
                    float a = acc[p][j] + b;
                    if (keep_act && (m0 + pp) < P.M) Abase[(size_t)l * Apstride + (size_t)pp * 256 + f] = a;
                    float hval = softplus100(a);
                    if (feeds_skip) {   // cat([x, embed]) / sqrt(2)   mlp.py:94-95
                        hval = (f < nsplit) ? hval : Es[(f - nsplit) * HS + pp];
                        hval = __fdiv_rn(hval, SQRT2);
                    }
                    if (last_hidden) head[p] = fmaf(hval, hw, head[p]);
                    if (write_h) Hs[f * HS + pp] = hval;
                }
            }
            __syncthreads();
        }
        // ---- 4. sdf head
        {
            const float hb = net.sdf_head[256];
#pragma unroll
            for (int p = 0; p < 8; ++p) {
                float s = warp_sum(head[p]) + hb;
                long long m = m0 + ty * 8 + p;
                if (tx == 0 && m < P.M) P.out_sdf[m] = s;
            }
        }
        if (!full) continue;

        // ---- 5. feature layer (rows 1..256 of the last SDF layer), result replaces Hs
        {
            float acc[8][8];
#pragma unroll
            for (int p = 0; p < 8; ++p)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[p][j] = 0.f;
            tile_gemm<8, false>(Hs, net.sdf_wt[L - 1], 256, Wb, acc);
            __syncthreads();
            const float* bias = net.sdf_b[L - 1];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int f = tx + 32 * j;
                const float b = bias[f];
#pragma unroll
                for (int p = 0; p < 8; ++p) {
                    const int pp = ty * 8 + p;
                    float v = acc[p][j] + b;
                    if (P.out_feat && (m0 + pp) < P.M) P.out_feat[(size_t)(m0 + pp) * 256 + f] = v;
                    Hs[f * HS + pp] = v;
                }
            }
            __syncthreads();
        }
        // ---- 6. light-mask head: sigmoid(w1 . softplus100(W0 relu(feat) + b0) + b1)   network/__init__.py:162-168
        if (P.want_light) {
            float acc[8][4];
#pragma unroll
            for (int p = 0; p < 8; ++p)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[p][j] = 0.f;
            tile_gemm<4, true>(Hs, net.light_wt0, 256, Wb, acc);
            float part[8];
#pragma unroll
            for (int p = 0; p < 8; ++p) part[p] = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int f = tx + 32 * j;
                const float b = net.light_b0[f], w = net.light_head[f];
#pragma unroll
                for (int p = 0; p < 8; ++p) part[p] = fmaf(softplus100(acc[p][j] + b), w, part[p]);
            }
            const float hb = net.light_head[net.lh];
#pragma unroll
            for (int p = 0; p < 8; ++p) {
                float s = warp_sum(part[p]) + hb;
                long long m = m0 + ty * 8 + p;
                if (tx == 0 && m < P.M) P.out_light[m] = sigmoidf_(s);
            }
            __syncthreads();   // light GEMM readers done before colour rows / Hs are rewritten
        }
        // ---- 7. radiance stack on [feat | PE(dir)]   mlp.py:208-229
        if (P.want_color) {
            for (int i = tid; i < (HROWS - 256) * TM; i += NT) {
                int r = i / TM, p = i - r * TM;
                float v = 0.f;
                if (r < 3) v = Ds[r * TM + p];
                else if (r < net.ed) {
                    int q = r - 3, k = q / 6, s = (q % 6) / 3, c = q % 3;
                    float arg = __fmul_rn(Ds[c * TM + p], (float)(1 << k));
                    v = s ? cosf(arg) : sinf(arg);
                }
                Hs[(256 + r) * HS + p] = v;
            }
            __syncthreads();
            float part[3][8];
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int p = 0; p < 8; ++p) part[c][p] = 0.f;
            const int Lc = net.Lc;
            for (int l = 0; l < Lc - 1; ++l) {
                float acc[8][8];
#pragma unroll
                for (int p = 0; p < 8; ++p)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[p][j] = 0.f;
                tile_gemm<8, false>(Hs, net.col_wt[l], net.col_kpad[l], Wb, acc);
                __syncthreads();
                const float* bias = net.col_b[l];
                const bool last_hidden = (l == Lc - 2);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int f = tx + 32 * j;
                    const float b = bias[f];
                    float w0 = 0.f, w1 = 0.f, w2 = 0.f;
                    if (last_hidden) { w0 = net.col_head[f]; w1 = net.col_head[256 + f]; w2 = net.col_head[512 + f]; }
#pragma unroll
                    for (int p = 0; p < 8; ++p) {
                        float hval = fmaxf(acc[p][j] + b, 0.f);
                        if (last_hidden) {
                            part[0][p] = fmaf(hval, w0, part[0][p]);
                            part[1][p] = fmaf(hval, w1, part[1][p]);
                            part[2][p] = fmaf(hval, w2, part[2][p]);
                        } else {
                            Hs[f * HS + ty * 8 + p] = hval;
                        }
                    }
                }
                __syncthreads();
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float hb = net.col_head[768 + c];
#pragma unroll
                for (int p = 0; p < 8; ++p) {
                    float s = warp_sum(part[c][p]) + hb;
                    long long m = m0 + ty * 8 + p;
                    if (tx == 0 && m < P.M) P.out_rgb[m * 3 + c] = sigmoidf_(s);
                }
            }
        }
        // ---- 8. reverse sweep: d sdf / d x   (what autograd.grad does at mlp.py:134-140)
        if (want_grad) {
            __syncthreads();
            // adjoint of a_{L-2}: w_sdf * softplus'(a_{L-2})
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int f = tx + 32 * j;
                const float w = net.sdf_head[f];
#pragma unroll
                for (int p = 0; p < 8; ++p) {
                    const int pp = ty * 8 + p;
                    float a = ((m0 + pp) < P.M) ? Abase[(size_t)(L - 2) * Apstride + (size_t)pp * 256 + f] : 0.f;
                    Hs[f * HS + pp] = w * dsoftplus100(a);
                }
            }
            for (int i = tid; i < EROWS * TM; i += NT) REs[(i / TM) * HS + (i % TM)] = 0.f;
            __syncthreads();
            for (int l = L - 2; l >= 1; --l) {
                float acc[8][8];
#pragma unroll
                for (int p = 0; p < 8; ++p)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[p][j] = 0.f;
                tile_gemm<8, false>(Hs, net.sdf_wr[l], 256, Wb, acc);   // adj(input of layer l)
                __syncthreads();
                const bool is_skip = (l == net.skip);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int f = tx + 32 * j;
#pragma unroll
                    for (int p = 0; p < 8; ++p) {
                        const int pp = ty * 8 + p;
                        float r = acc[p][j];
                        if (is_skip) {
                            r = __fdiv_rn(r, SQRT2);
                            if (f >= nsplit) { REs[(f - nsplit) * HS + pp] = r; r = 0.f; }
                        }
                        float a = ((m0 + pp) < P.M) ? Abase[(size_t)(l - 1) * Apstride + (size_t)pp * 256 + f] : 0.f;
                        Hs[f * HS + pp] = r * dsoftplus100(a);
                    }
                }
                __syncthreads();
            }
            {   // layer 0: adjoint of the embedding (39 wide, padded to 64 columns)
                float acc[8][2];
#pragma unroll
                for (int p = 0; p < 8; ++p) { acc[p][0] = 0.f; acc[p][1] = 0.f; }
                tile_gemm<2, false>(Hs, net.sdf_wr[0], 256, Wb, acc);
                __syncthreads();
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int f = tx + 32 * j;
                    if (f < EROWS) {
#pragma unroll
                        for (int p = 0; p < 8; ++p) {
                            const int pp = ty * 8 + p;
                            Hs[f * HS + pp] = acc[p][j] + REs[f * HS + pp];
                        }
                    }
                }
                __syncthreads();
            }
            // J^T r: d/dx of [x, sin(f x), cos(f x)]
            if (tid < 3 * TM) {
                const int c = tid / TM, p = tid - c * TM;
                float g = Hs[c * HS + p];
                for (int k = 0; k < net.mx; ++k) {
                    const float fk = (float)(1 << k);
                    const int rs = 3 + 6 * k + c, rc = rs + 3;
                    g += fk * (Es[rc * HS + p] * Hs[rs * HS + p] - Es[rs * HS + p] * Hs[rc * HS + p]);
                }
                if ((m0 + p) < P.M) P.out_grad[(m0 + p) * 3 + c] = g;
            }
        }
    }
}

}  // namespace simt

size_t mlp_simt_scratch_floats(const i2sdf_handle* h) {
    return (size_t)(2 * h->num_sms) * (size_t)(h->net.L - 1) * simt::TM * 256;
}

int launch_mlp_simt(const i2sdf_handle* h, const MlpParams& p, cudaStream_t stream) {
    static PerDeviceOnce once;
    if (once.need()) I2SDF_CUDA_CHECK(cudaFuncSetAttribute(simt::mlp_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)simt::kSmemBytes));
    if (p.M <= 0) return I2SDF_OK;
    long long ntiles = (p.M + simt::TM - 1) / simt::TM;
    int grid = (int)(ntiles < (long long)(2 * h->num_sms) ? ntiles : (long long)(2 * h->num_sms));
    simt::mlp_tile_kernel<<<grid, simt::NT, simt::kSmemBytes, stream>>>(p);
    I2SDF_CUDA_CHECK(cudaGetLastError());
    return I2SDF_OK;
}

}  // namespace i2sdf
