// Shared declarations of the i2sdf_b200 CUDA core (sm_100a).
#pragma once
#include <stdlib.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/i2sdf_b200.h"
#include "planes.cuh"

namespace i2sdf {

constexpr int kMaxLayers = 12;
constexpr int kHidden = 256;

// Packed device-side network (fp32 SIMT layouts).  All pointers are device memory owned by the handle.
struct NetDev {
    int L;            // SDF linear layers
    int skip;         // skip layer index or -1
    int mx, ex;       // multires of points, embedding width (3+6*mx)
    int exp_;         // ex padded to a multiple of 8
    int Lc;           // colour linear layers
    int md, ed;       // multires of dirs, embedding width
    int Ll;           // light layers (0 or 2)
    int lh;           // light hidden (128)
    // SDF net
    const float* sdf_wt[kMaxLayers];   // forward, k-major [kpad][256]; last layer: feature rows (W[1:])
    int sdf_kpad[kMaxLayers];
    const float* sdf_wr[kMaxLayers];   // reverse sweep, row-major [256][ldr] (ldr = 256, layer 0: 64)
    const float* sdf_b[kMaxLayers];    // [256] (last layer: b[1:])
    const float* sdf_head;             // [257]: W_last[0,:], then b_last[0]
    // colour net (input rows reordered: [feat(256) | PE(d)(ed) | 0-pad])
    const float* col_wt[kMaxLayers];
    int col_kpad[kMaxLayers];
    const float* col_b[kMaxLayers];
    const float* col_head;             // [3][256] then [3] bias
    // light head
    const float* light_wt0;            // [256][128] k-major
    const float* light_b0;             // [128]
    const float* light_head;           // [128] then bias
};

struct SamplerDev {
    int n_samples, n_eval, n_extra, beta_iters, max_iters;
    float near_, far_, eps, add_tiny, beta_min;
    const float* u_up;       // [n_eval]
    const float* u_final;    // [n_samples]
    const float* t_init;     // [n_eval]
    const int* extra_idx;    // [max_iters][n_extra]
};

}  // namespace i2sdf

namespace i2sdf {
// Grid of a persistent tile kernel (one CTA per SM, tiles dealt round-robin): every CTA runs ceil(ntiles / grid) rounds whatever the grid, so
// the smallest grid with the same number of rounds does the same work in the same number of tile times with fewer SMs contending for
// HBM / L2 and less power drawn (the B200 is power-capped under these kernels): 800 tiles -> 6 rounds on 134 CTAs instead of 148 with
// 60 CTAs busy in the last round.  I2SDF_GRID_BALANCE=0 restores min(ntiles, SMs).
inline int balanced_grid(int num_sms, long long ntiles) {
    static const bool on = [] { const char* e = getenv("I2SDF_GRID_BALANCE"); return !(e && e[0] == '0'); }();
    if (ntiles <= (long long)num_sms) return (int)(ntiles > 0 ? ntiles : 1);
    if (!on) return num_sms;
    const long long rounds = (ntiles + num_sms - 1) / num_sms;
    return (int)((ntiles + rounds - 1) / rounds);
}
}  // namespace i2sdf

struct i2sdf_handle {
    i2sdf_desc desc;
    int device;
    int num_sms;
    i2sdf::NetDev net;
    i2sdf::SamplerDev smp;
    float* pool;             // one allocation for all packed weights / tables
    size_t pool_floats;
    // layer bookkeeping for pack_weights: (out,in) of every layer in API order
    int n_layers;
    int lay_out[2 * i2sdf::kMaxLayers + 4];
    int lay_in[2 * i2sdf::kMaxLayers + 4];
    bool use_tc;             // tcgen05 path available (set by env I2SDF_SIMT=1 -> false)
    void* tc;                // opaque tcgen05 packed state (see mlp_tc.cu)
    void* prof;              // measurement hook state (c_abi.cu)
    void* tcmain;            // tcgen05 main-pass state (mlp_tc_main.cu), null if unavailable for this network
    bool fused;              // training state travels as plane slots + fused backward (env I2SDF_FUSED_BWD=0 -> fp32 layer-wise)
};

namespace i2sdf {

void set_error(const char* fmt, ...);

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a PER-DEVICE setting: a process-wide `static bool done` (round 1) left the kernels
// of a second device in the same process (model.to('cuda:1')) without their shared-memory opt-in.  One flag per device and call site.
struct PerDeviceOnce {
    bool done[64] = {};
    bool need() {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;
        if (done[dev]) return false;
        done[dev] = true;
        return true;
    }
};

#define I2SDF_CUDA_CHECK(expr)                                                                 \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            i2sdf::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return I2SDF_E_CUDA;                                                               \
        }                                                                                      \
    } while (0)

// ---- SIMT fused MLP (mlp_simt.cu) ----------------------------------------------------------------
struct MlpParams {
    // point source: explicit pts, or rays: point m = (ray m / ns, sample m % ns) -> o + z * d
    const float* pts;
    const float* ray_o;
    const float* ray_d;
    const float* zarr;
    int zstride;
    int ns;
    long long M;
    long long m_rays;         // with BOTH sources given: points m < m_rays come from the rays, m >= m_rays from pts[m - m_rays]
    // third source (sdf-only kernel): a regular grid generated on the device, numpy.meshgrid(x, y, z) order as utils/plots.py:440-451:
    // point m = (j, i, k) = (m / (nx nz), (m / nz) % nx, m % nz) -> (gx[i], gy[j], gz[k]), then p' = A p + t if grid_affine (12 floats:
    // row-major 3x3 A, then t)
    const float* gx; const float* gy; const float* gz;
    int nx, ny, nz;
    const float* grid_affine;
    // predication for sampler rounds: run only if round_idx < 0 or all beta_max[j] > beta0 for j < round_idx
    const float* beta_max;    // device [max_iters]
    const float* beta_param;  // device scalar (raw density.beta)
    float beta_min;
    int round_idx;
    // outputs (null = skip)
    float* out_sdf;
    float* out_feat;
    float* out_grad;
    float* out_rgb;
    float* out_light;
    float* save_act;          // [L-1][M][256] pre-activations (training) or null
    float* scratch;           // per-CTA [(L-1)][TM][256] when grad wanted without save_act
    int pf_op_ahead;          // main pass: pull the next op's stored h~ segments into L2 one op ahead (set by tcmain_launch; I2SDF_MAIN_PREFETCH=0 off)
    void* tl;                 // development probe (I2SDF_DEBUG_TIMELINE): 64 KB of clock64 stamps, tensor-core main pass only
    planes::Layout sl;        // tensor-core main pass in training: the saved state is plane slots (sl.base != null)
    int want_color;
    int want_light;
    NetDev net;
};

int launch_mlp_simt(const i2sdf_handle* h, const MlpParams& p, cudaStream_t stream);
size_t mlp_simt_scratch_floats(const i2sdf_handle* h);

// ---- tensor-core weight blocks (mlp_tc3.cu) used by the backward GEMMs (tc_gemm.cu) -----------------
enum { TCB_FWD_SDF = 0, TCB_FWD_FEAT, TCB_FWD_COL, TCB_REV_SDF, TCB_REV_FEAT, TCB_REV_COL, TCB_FWD_LIGHT };
struct TcBlock { const uint8_t* ptr; int ksteps; int n; };     // bf16 hi/lo k-step blocks of n*64 bytes
TcBlock tc_block(const i2sdf_handle* h, int role, int layer);

// ---- sampler / compositing (sampler.cu) -----------------------------------------------------------
struct SamplerWs {            // carved from the caller's workspace
    float* z[2];              // [R][zmax] ping-pong sorted z
    float* sdf[2];            // [R][zmax]
    float* samples;           // [R][n_eval] new samples of the round (unsorted order = inverse-CDF order)
    float* sdf_new;           // [R][n_eval]
    int* src;                 // [R][zmax] merge source index
    float* beta;              // [R]
    float* beta_max;          // [max_iters]  (atomicMax as int)
    int zmax;
};

}  // namespace i2sdf
