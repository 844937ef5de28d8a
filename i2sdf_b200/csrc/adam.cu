// Adam step of ALL parameter tensors in one launch.
//
// Reference: torch.optim.Adam(self.model.get_param_groups(lr), eps=1e-15) stepped once per training step
// (model/trainer/recon.py:201-207, 254-287).  PyTorch's fused path needs 4 multi_tensor_apply launches of ~28 us for the 44
// small tensors of this model; one job table covers them here.  Arithmetic = ATen's fused Adam (fused_adam_utils.cuh):
//   m = m + (1 - b1) (g - m) ;  v = b2 v + (1 - b2) g^2 ;  p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
#include "common.cuh"

namespace i2sdf {
namespace adamk {

// scalars (optional, device): {step_size, bias_correction2_sqrt} of THIS step, read at run time instead of the values baked into the
// launch - what a CUDA-graph replay needs (the host refreshes them through a captured pinned-memory copy, i2sdf_b200/graph.py)
__global__ void __launch_bounds__(256) adam_kernel(const i2sdf_adam_batch B, const float* __restrict__ scalars) {
    const i2sdf_adam_job& J = B.jobs[blockIdx.y];
    const float one_m_b1 = B.one_minus_beta1, one_m_b2 = B.one_minus_beta2;      // formed in double on the host, as ATen does
    const float step_size = scalars ? scalars[0] : B.step_size, bc2 = scalars ? scalars[1] : B.bias_correction2_sqrt;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < J.numel; i += (long long)gridDim.x * blockDim.x) {
        const float g = J.grad[i];
        float m = J.exp_avg[i], v = J.exp_avg_sq[i];
        m = fmaf(one_m_b1, g - m, m);                              // lerp(m, g, 1 - beta1)
        v = B.beta2 * v + one_m_b2 * g * g;
        const float denom = __fdiv_rn(sqrtf(v), bc2) + B.eps;
        J.param[i] = J.param[i] - step_size * __fdiv_rn(m, denom);
        J.exp_avg[i] = m;
        J.exp_avg_sq[i] = v;
    }
}

}  // namespace adamk
}  // namespace i2sdf

extern "C" int i2sdf_adam_step_dev(const i2sdf_adam_batch* b, const float* scalars_dev, void* stream);
extern "C" int i2sdf_adam_step(const i2sdf_adam_batch* b, void* stream) { return i2sdf_adam_step_dev(b, nullptr, stream); }
extern "C" int i2sdf_adam_step_dev(const i2sdf_adam_batch* b, const float* scalars_dev, void* stream) {
    using namespace i2sdf;
    if (!b || b->n < 1 || b->n > I2SDF_ADAM_MAX_JOBS) { set_error("adam_step: bad job count"); return I2SDF_E_INVALID; }
    long long most = 0;
    for (int i = 0; i < b->n; ++i) {
        const i2sdf_adam_job& J = b->jobs[i];
        if (!J.param || !J.grad || !J.exp_avg || !J.exp_avg_sq || J.numel < 1) { set_error("adam_step: job %d has a null pointer or is empty", i); return I2SDF_E_INVALID; }
        most = J.numel > most ? J.numel : most;
    }
    int gx = (int)((most + 1023) / 1024);
    if (gx > 128) gx = 128;
    adamk::adam_kernel<<<dim3(gx, b->n), 256, 0, (cudaStream_t)stream>>>(*b, scalars_dev);
    I2SDF_CUDA_CHECK(cudaGetLastError());
    return I2SDF_OK;
}
