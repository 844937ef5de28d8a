// tcgen05 SDF kernel — placeholder until the tensor-core path lands.
#include "common.cuh"
namespace i2sdf {
int tc_create(i2sdf_handle* h) { h->use_tc = false; return I2SDF_OK; }
void tc_destroy(i2sdf_handle*) {}
int tc_pack(i2sdf_handle*, const float* const*, const float* const*, cudaStream_t) { return I2SDF_OK; }
int tc_launch_sdf(const i2sdf_handle* h, const MlpParams& p, cudaStream_t st) { return launch_mlp_simt(h, p, st); }
}
