// Backward chain kernel (mlp_tc_bwd.cu): parameters, op kinds, host entry points.
#pragma once
#include "common.cuh"

namespace i2sdf {

namespace chain { struct OpTable; }

namespace tcb {
// epilogue kinds of the backward op tables (built in tc_create, mlp_tc3.cu)
enum { BK_TAN = 0, BK_COL_REV, BK_FEAT_ADJ, BK_P };
}

struct BwdParams {
    // point source (as MlpParams): explicit pts, or rays: point m = (ray m / ns, sample m % ns) -> o + z * d
    const float* pts;
    const float* ray_o;
    const float* ray_d;
    const float* zarr;
    int zstride;
    int ns;
    long long M;
    long long m_rays;      // as MlpParams: both sources given -> first m_rays points from the rays, the rest from pts
    // upstream gradients per point (null = 0)
    // the upstream arrays cover the first m_up points (m_up = M unless the caller appended points behind its ray samples); points
    // m >= m_up have no sdf / rgb upstream and take the upstream of grad_x sdf from g_grad_tail[m - m_up]
    long long m_up;
    const float* g_grad_tail;   // [M - m_up][3] or null
    const float* g_sdf;    // [m_up]
    const float* g_grad;   // [m_up][3]   upstream of grad_x sdf
    const float* g_rgb;    // [M][3]   (with_color)
    const float* s_rgb;    // [M][3]   forward rgb (sigmoid output), for its derivative
    int with_color;
    long long* tl;         // development probe (I2SDF_DEBUG_TIMELINE, tools/timeline.py bwd): 64 KB of clock64 stamps of CTA 0's second tile
    int pf_dist;           // slot segments are prefetched into L2 this many 16-column items ahead (set by tc_bwd_launch)
    planes::Layout sl;     // saved forward slots (sl.base) + backward workspace slots (sl.wbase)
    NetDev net;
};

const chain::OpTable* tc_bwd_table(const i2sdf_handle* h, bool with_color);
int tc_bwd_launch(const i2sdf_handle* h, const BwdParams& p, cudaStream_t st);
long long tc_bwd_timeline_read(long long* out, long long n);

}  // namespace i2sdf
