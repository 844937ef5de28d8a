// Weight-norm reparametrisation of ALL layers in one launch per direction.
//
// Reference: nn.utils.weight_norm(lin) with dim = 0 (mlp.py:71-72, 200-201): W[r, :] = g[r] * v[r, :] / ||v[r, :]||_2, recomputed
// before every forward; autograd differentiates it.  PyTorch runs one small kernel per layer and direction (14 + 14 launches
// per training step for config/synthetic.yml); here a job table covers every layer, one warp per row.
//   forward : norm[r] = sqrt(sum_k v[r,k]^2) ;  W[r,k] = (g[r] * v[r,k]) * (1 / norm[r])          (ATen WeightNorm.cu arithmetic)
//   backward: s = sum_k dW[r,k] v[r,k] ;  dg[r] = s / norm ;  dv[r,k] = g[r] * (dW[r,k] / norm - v[r,k] * s / norm^3)
#include "common.cuh"

namespace i2sdf {
namespace wn {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void __launch_bounds__(256) wn_fwd_kernel(const i2sdf_wnorm_batch B) {
    const i2sdf_wnorm_job& J = B.jobs[blockIdx.y];
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < J.rows; r += gridDim.x * wpb) {
        const float* v = J.v + (size_t)r * J.cols;
        float s = 0.f;
        for (int k = lane; k < J.cols; k += 32) { const float x = v[k]; s = fmaf(x, x, s); }
        const float nrm = sqrtf(warp_sum(s));
        if (lane == 0) J.norm[r] = nrm;
        const float gr = J.g[r], rn = __fdiv_rn(1.0f, nrm);
        float* w = J.W + (size_t)r * J.cols;
        for (int k = lane; k < J.cols; k += 32) w[k] = (gr * v[k]) * rn;
    }
}

__global__ void __launch_bounds__(256) wn_bwd_kernel(const i2sdf_wnorm_batch B) {
    const i2sdf_wnorm_job& J = B.jobs[blockIdx.y];
    if (!J.dW) return;                                          // this layer's W received no gradient: dg, dv stay as the caller left them
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < J.rows; r += gridDim.x * wpb) {
        const float* v = J.v + (size_t)r * J.cols;
        const float* dw = J.dW + (size_t)r * J.cols;
        float s = 0.f;
        for (int k = lane; k < J.cols; k += 32) s = fmaf(dw[k], v[k], s);
        s = warp_sum(s);
        const float rn = __fdiv_rn(1.0f, J.norm[r]), rn3 = rn * rn * rn, gr = J.g[r];
        if (lane == 0) J.dg[r] = s * rn;
        float* dv = J.dv + (size_t)r * J.cols;
        for (int k = lane; k < J.cols; k += 32) dv[k] = gr * (rn * dw[k] - rn3 * v[k] * s);
    }
}

}  // namespace wn
}  // namespace i2sdf

extern "C" int i2sdf_weight_norm(const i2sdf_wnorm_batch* b, int backward, void* stream) {
    using namespace i2sdf;
    if (!b || b->n < 1 || b->n > I2SDF_WNORM_MAX_JOBS) { set_error("weight_norm: bad job count"); return I2SDF_E_INVALID; }
    for (int i = 0; i < b->n; ++i) {
        const i2sdf_wnorm_job& J = b->jobs[i];
        if (!J.g || !J.v || !J.norm || J.rows < 1 || J.cols < 1 || (!backward && !J.W) || (backward && J.dW && (!J.dg || !J.dv))) {
            set_error("weight_norm: job %d has a null pointer or an empty shape", i); return I2SDF_E_INVALID;
        }
    }
    const dim3 grid(32, b->n);                                  // 32 blocks x 8 warps: one pass over <= 257 rows per layer
    if (backward) wn::wn_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*b);
    else wn::wn_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*b);
    I2SDF_CUDA_CHECK(cudaGetLastError());
    return I2SDF_OK;
}
