// I2SDFLoss forward + the gradient of the loss w.r.t. every model output, in ONE launch.
//
// Reference: model/network/__init__.py:289-406 (I2SDFLoss).  In the reference (and in plain PyTorch) the loss and its
// backward are ~60 tiny kernels per step; at B200 kernel speeds their launch gaps, not their arithmetic, show up in the
// training step.  All terms are reductions over at most a few thousand rays, so one CTA does everything:
//   phase 1: block-wide sums (double accumulators) of every term and the mask counts
//   phase 2: d loss / d input for every differentiable input (closed forms of what autograd would produce)
//
//   rgb      F.l1_loss(rgb, gt)                                  :308-311   d = sign(a - b) / (3R)
//   eikonal  ((|g|_2 - 1)^2).mean()                              :313-315   d = 2 (|g| - 1) g / |g| / n   (0 at |g| = 0, as torch)
//   smooth   diff_norm.mean()                                    :347-351
//   mask     BCE(clip(weight_sum, 1e-3, 1 - 1e-3), mask)         :317-318   d = (p - t) / (p (1 - p)) / R inside the clip range
//   depth    mse_loss(depth[mask], gt[mask])                     :320-324   masked mean (0/0 = NaN for an empty mask, as torch)
//   normal   |1 - <n, gt>| masked mean                           :326-329   used for BOTH normal_loss and angular_loss (:363-371)
//   bubble   surface_sdf.abs().mean()                            :373-376
//   light    BCE(clip(light_mask), light_gt)                     :378-381
#include "common.cuh"

namespace i2sdf {
namespace lossk {

constexpr int NT = 1024;
constexpr int NACC = 10;
enum { A_RGB = 0, A_EIK, A_SMOOTH, A_MASK, A_DEPTH, A_DEPTH_N, A_NORMAL, A_NORMAL_N, A_BUBBLE, A_LIGHT };

__device__ __forceinline__ float sgn(float x) { return (x > 0.f) ? 1.f : ((x < 0.f) ? -1.f : 0.f); }
// torch.nn.functional.binary_cross_entropy on p = clip(x, 1e-3, 1 - 1e-3) (log clamped at -100 as torch does; never active here)
__device__ __forceinline__ float bce(float x, float t, float* dx) {
    const float lo = 1e-3f, hi = 1.0f - 1e-3f;
    const float p = fminf(fmaxf(x, lo), hi);
    const float l = -(t * fmaxf(logf(p), -100.f) + (1.f - t) * fmaxf(logf(1.f - p), -100.f));
    *dx = (x >= lo && x <= hi) ? __fdiv_rn(p - t, p * (1.f - p)) : 0.f;
    return l;
}

__global__ void __launch_bounds__(NT, 1) loss_kernel(const i2sdf_loss_args a) {
    __shared__ double red[NACC][NT / 32];
    __shared__ double tot[NACC];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double acc[NACC];
#pragma unroll
    for (int k = 0; k < NACC; ++k) acc[k] = 0.0;
    const long long R = a.R;
    for (long long r = tid; r < R; r += NT) {
#pragma unroll
        for (int c = 0; c < 3; ++c) acc[A_RGB] += fabsf(a.rgb[r * 3 + c] - a.rgb_gt[r * 3 + c]);
        if (a.diff_norm) acc[A_SMOOTH] += a.diff_norm[r];
        if (a.weight_sum) { float d; acc[A_MASK] += bce(a.weight_sum[r], a.mask_gt[r], &d); }
        if (a.depth && a.depth_mask[r]) { const float e = a.depth[r] - a.depth_gt[r]; acc[A_DEPTH] += e * e; acc[A_DEPTH_N] += 1.0; }
        if (a.normal && a.normal_mask[r]) {
            const float dot = a.normal[r * 3] * a.normal_gt[r * 3] + a.normal[r * 3 + 1] * a.normal_gt[r * 3 + 1] + a.normal[r * 3 + 2] * a.normal_gt[r * 3 + 2];
            acc[A_NORMAL] += fabsf(1.f - dot); acc[A_NORMAL_N] += 1.0;
        }
        if (a.light) { float d; acc[A_LIGHT] += bce(a.light[r], a.light_gt[r], &d); }
    }
    if (a.grad_theta)
        for (long long i = tid; i < a.n_eik; i += NT) {
            const float x = a.grad_theta[i * 3], y = a.grad_theta[i * 3 + 1], z = a.grad_theta[i * 3 + 2];
            const float e = sqrtf(x * x + y * y + z * z) - 1.f;
            acc[A_EIK] += e * e;
        }
    if (a.surface_sdf)
        for (long long i = tid; i < a.n_bubble; i += NT) acc[A_BUBBLE] += fabsf(a.surface_sdf[i]);
#pragma unroll
    for (int k = 0; k < NACC; ++k) {
        double v = acc[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[k][warp] = v;
    }
    __syncthreads();
    if (tid < NACC) {
        double v = 0.0;
        for (int w = 0; w < NT / 32; ++w) v += red[tid][w];
        tot[tid] = v;
    }
    __syncthreads();
    // divisors of the means: this call's own counts, or (sharded batches) the caller's share of the global counts
    const double n_ray = a.denom ? (double)a.denom[I2SDF_LOSS_DENOM_RAYS] : (double)R;
    const double n_eik = a.denom ? (double)a.denom[I2SDF_LOSS_DENOM_EIK] : (double)a.n_eik;
    const double n_bub = a.denom ? (double)a.denom[I2SDF_LOSS_DENOM_BUBBLE] : (double)a.n_bubble;
    const double n_depth = a.denom ? (double)a.denom[I2SDF_LOSS_DENOM_DEPTH] : tot[A_DEPTH_N];
    const double n_normal = a.denom ? (double)a.denom[I2SDF_LOSS_DENOM_NORMAL] : tot[A_NORMAL_N];
    if (tid == 0) {
        const float t_rgb = (float)(tot[A_RGB] / (3.0 * n_ray));
        const float t_eik = a.grad_theta ? (float)(tot[A_EIK] / n_eik) : 0.f;
        const float t_smooth = a.diff_norm ? (float)(tot[A_SMOOTH] / n_ray) : 0.f;
        const float t_mask = a.weight_sum ? (float)(tot[A_MASK] / n_ray) : 0.f;
        const float t_depth = a.depth ? (float)(tot[A_DEPTH] / n_depth) : 0.f;
        const float t_normal = a.normal ? (float)(tot[A_NORMAL] / n_normal) : 0.f;
        const float t_bubble = a.surface_sdf ? (float)(tot[A_BUBBLE] / n_bub) : 0.f;
        const float t_light = a.light ? (float)(tot[A_LIGHT] / n_ray) : 0.f;
        const float t_n = (a.w_normal > 0.f) ? t_normal : 0.f, t_a = (a.w_angular > 0.f) ? t_normal : 0.f;
        // same association order as the reference's sum (:383-391)
        float loss = t_rgb;
        loss += a.w_eik * t_eik; loss += a.w_smooth * t_smooth; loss += a.w_mask * t_mask; loss += a.w_depth * t_depth;
        loss += a.w_normal * t_n; loss += a.w_angular * t_a; loss += a.w_bubble * t_bubble; loss += a.w_light * t_light;
        a.terms[0] = loss; a.terms[1] = t_rgb; a.terms[2] = t_eik; a.terms[3] = t_smooth; a.terms[4] = t_mask; a.terms[5] = t_depth;
        a.terms[6] = t_n; a.terms[7] = t_a; a.terms[8] = t_bubble; a.terms[9] = t_light;
    }
    // ---- phase 2: gradients
    const float k_rgb = (float)(1.0 / (3.0 * n_ray));
    const float k_ray = (float)(1.0 / n_ray);
    const float k_depth = (float)((double)a.w_depth * 2.0 / n_depth);
    const float k_normal = (float)((double)(a.w_normal + a.w_angular) / n_normal);
    for (long long r = tid; r < R; r += NT) {
        if (a.g_rgb) {
#pragma unroll
            for (int c = 0; c < 3; ++c) a.g_rgb[r * 3 + c] = sgn(a.rgb[r * 3 + c] - a.rgb_gt[r * 3 + c]) * k_rgb;
        }
        if (a.g_diff_norm) a.g_diff_norm[r] = a.w_smooth * k_ray;
        if (a.g_weight_sum) { float d; bce(a.weight_sum[r], a.mask_gt[r], &d); a.g_weight_sum[r] = a.w_mask * d * k_ray; }
        if (a.g_depth) a.g_depth[r] = a.depth_mask[r] ? (a.depth[r] - a.depth_gt[r]) * k_depth : 0.f;
        if (a.g_normal) {
            float gx = 0.f, gy = 0.f, gz = 0.f;
            if (a.normal_mask[r]) {
                const float nx = a.normal_gt[r * 3], ny = a.normal_gt[r * 3 + 1], nz = a.normal_gt[r * 3 + 2];
                const float dot = a.normal[r * 3] * nx + a.normal[r * 3 + 1] * ny + a.normal[r * 3 + 2] * nz;
                const float s = -sgn(1.f - dot) * k_normal;
                gx = s * nx; gy = s * ny; gz = s * nz;
            }
            a.g_normal[r * 3] = gx; a.g_normal[r * 3 + 1] = gy; a.g_normal[r * 3 + 2] = gz;
        }
        if (a.g_light) { float d; bce(a.light[r], a.light_gt[r], &d); a.g_light[r] = a.w_light * d * k_ray; }
    }
    if (a.g_grad_theta) {
        const float k_eik = (float)(2.0 * (double)a.w_eik / n_eik);
        for (long long i = tid; i < a.n_eik; i += NT) {
            const float x = a.grad_theta[i * 3], y = a.grad_theta[i * 3 + 1], z = a.grad_theta[i * 3 + 2];
            const float nrm = sqrtf(x * x + y * y + z * z);
            const float s = (nrm > 0.f) ? k_eik * __fdiv_rn(nrm - 1.f, nrm) : 0.f;
            a.g_grad_theta[i * 3] = s * x; a.g_grad_theta[i * 3 + 1] = s * y; a.g_grad_theta[i * 3 + 2] = s * z;
        }
    }
    if (a.g_surface_sdf) {
        const float k_b = (float)((double)a.w_bubble / n_bub);
        for (long long i = tid; i < a.n_bubble; i += NT) a.g_surface_sdf[i] = sgn(a.surface_sdf[i]) * k_b;
    }
}

}  // namespace lossk
}  // namespace i2sdf

extern "C" int i2sdf_loss_forward(const i2sdf_loss_args* a, void* stream) {
    using namespace i2sdf;
    if (!a || !a->rgb || !a->rgb_gt || !a->terms || a->R < 1) { set_error("loss_forward: rgb, rgb_gt, terms and R >= 1 are required"); return I2SDF_E_INVALID; }
    if ((a->weight_sum && !a->mask_gt) || (a->depth && (!a->depth_gt || !a->depth_mask)) || (a->normal && (!a->normal_gt || !a->normal_mask)) ||
        (a->light && !a->light_gt) || (a->grad_theta && a->n_eik < 1) || (a->surface_sdf && a->n_bubble < 1)) {
        set_error("loss_forward: a term's prediction was given without its target / mask / row count"); return I2SDF_E_INVALID;
    }
    if ((a->g_grad_theta && !a->grad_theta) || (a->g_diff_norm && !a->diff_norm) || (a->g_weight_sum && !a->weight_sum) || (a->g_depth && !a->depth) ||
        (a->g_normal && !a->normal) || (a->g_surface_sdf && !a->surface_sdf) || (a->g_light && !a->light)) {
        set_error("loss_forward: gradient output requested for an absent term"); return I2SDF_E_INVALID;
    }
    lossk::loss_kernel<<<1, lossk::NT, 0, (cudaStream_t)stream>>>(*a);
    I2SDF_CUDA_CHECK(cudaGetLastError());
    return I2SDF_OK;
}
