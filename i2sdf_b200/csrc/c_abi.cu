// C-ABI of the i2sdf_b200 core (include/i2sdf_b200.h): handle life-cycle, weight packing, entry points.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "common.cuh"
#include "planes.cuh"
#include "wgrad_planes.cuh"
#include "tc_bwd.cuh"

namespace i2sdf {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// declared in sampler.cu
int launch_rays(const float*, const float*, const float*, int, int, float*, float*, float*, cudaStream_t);
size_t sampler_ws_floats(const i2sdf_handle*, long long);
SamplerWs carve_sampler_ws(const i2sdf_handle*, long long, float*);
int launch_sampler_init(const i2sdf_handle*, const SamplerWs&, long long, const float*, float, cudaStream_t);
int launch_sampler_round(const i2sdf_handle*, const SamplerWs&, long long, int, int, const float*, const float*, cudaStream_t);
int launch_sampler_finalize(const i2sdf_handle*, const SamplerWs&, long long, const float*, const int*, int, const int*, float*,
                            float*, int*, cudaStream_t);
int launch_sampler_round_debug(const i2sdf_handle*, const float*, const float*, long long, int, const float*, const float*, int,
                               const float*, float*, float*, int*, float*, float*, int*, cudaStream_t);
int launch_composite(const i2sdf_handle*, const float*, const float*, const float*, const float*, const float*, const float*,
                     const float*, long long, int, float*, float*, float*, float*, float*, float*, cudaStream_t);
// declared in backward.cu
namespace bwd { struct PointSrc { const float* pts; const float* o; const float* d; const float* z; int zstride; int ns; }; }
size_t fused_backward_ws_bytes(const i2sdf_handle*, long long, bool);
int fused_backward(const i2sdf_handle*, const bwd::PointSrc&, long long, void*, const float*, const float*, const float*, const float*,
                   float* const*, float* const*, float* const*, float* const*, void*, int, long long, cudaStream_t, long long m_up, const float* g_grad_tail);
size_t sdf_backward_ws_floats(const i2sdf_handle*, long long);
size_t color_backward_ws_floats(const i2sdf_handle*, long long);
size_t light_backward_ws_floats(const i2sdf_handle*, long long);
int sdf_backward(const i2sdf_handle*, const bwd::PointSrc&, long long, const float* const*, const float*, const float*, const float*, int,
                 const float*, float* const*, float* const*, float*, cudaStream_t);
int color_backward(const i2sdf_handle*, long long, int, const float*, const float* const*, const float* const*, const float*, const float*,
                   const float*, float* const*, float* const*, float*, float**, int*, cudaStream_t);
int light_backward(const i2sdf_handle*, long long, const float* const*, const float* const*, const float*, const float*, const float*, const float*,
                   float* const*, float* const*, float*, cudaStream_t);
size_t light_forward_ws_floats(const i2sdf_handle*, long long);
int light_forward(const i2sdf_handle*, long long, const float*, float*, float*, cudaStream_t);
int launch_composite_backward(const i2sdf_handle*, const float*, const float*, const float*, const float*, const float*, const float*,
                              const float*, long long, int, const float*, const float*, const float*, const float*, const float*, float*,
                              float*, float*, float*, float*, cudaStream_t);
// declared in mlp_tc.cu
int tc_create(i2sdf_handle* h);
void tc_destroy(i2sdf_handle* h);
int tc_pack(i2sdf_handle* h, const float* const* W, const float* const* b, cudaStream_t st);
int tc_launch_sdf(const i2sdf_handle* h, const MlpParams& p, cudaStream_t st);
// declared in mlp_tc_main.cu
int tcmain_create(i2sdf_handle* h, void** out_state);
void tcmain_destroy(void* state);
int tcmain_pack(i2sdf_handle* h, void* state, const float* const* W, cudaStream_t st);
int tcmain_launch(const i2sdf_handle* h, void* state, const MlpParams& p, cudaStream_t st);
int tcmain_has_full(const void* state);

// ---- packing ------------------------------------------------------------------------------------
enum { PACK_T = 0, PACK_R = 1, PACK_V = 2 };
struct PackJob {
    int mode;
    float* dst;
    int rows, ld;          // dst is [rows][ld]
    const float* src;
    int out, in;           // src is [out][in] (or a vector of `out`)
    int row_off;           // PACK_T: feature f reads src row f+row_off ; PACK_V: element offset
    int nvalid;            // number of valid features / elements
    int feat_first;        // PACK_T colour layer 0: k<feat_first -> col ed+k ; k-feat_first<ed -> col k-feat_first
    int ed;
};

// All jobs of one i2sdf_pack_weights call run as ONE launch (grid.y = job): the training step re-packs after every
// optimizer step, so the ~45 small jobs must not cost ~45 launches.
constexpr int kMaxPackJobs = 3 * kMaxLayers + 2 + 2 * kMaxLayers + 2 + 4;
struct PackBatch { int n; PackJob jobs[kMaxPackJobs]; };

__global__ void pack_batch_kernel(const PackBatch B) {
    const PackJob& J = B.jobs[blockIdx.y];
    const long long n = (long long)J.rows * J.ld;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        int k = (int)(i / J.ld), f = (int)(i % J.ld);
        float v = 0.f;
        if (J.mode == PACK_T) {
            int col = -1;
            if (J.feat_first > 0) {
                if (k < J.feat_first) col = J.ed + k;
                else if (k - J.feat_first < J.ed) col = k - J.feat_first;
            } else if (k < J.in) col = k;
            if (col >= 0 && col < J.in && f < J.nvalid) v = J.src[(size_t)(f + J.row_off) * J.in + col];
        } else if (J.mode == PACK_R) {
            if (k < J.out && f < J.in) v = J.src[(size_t)k * J.in + f];
        } else {
            if (f < J.nvalid) v = J.src[J.row_off + f];
        }
        J.dst[i] = v;
    }
}

static int run_pack(PackBatch& B, const PackJob& J) {
    if (B.n >= kMaxPackJobs) { set_error("pack: too many jobs"); return I2SDF_E_INVALID; }
    B.jobs[B.n++] = J;
    return I2SDF_OK;
}
static int flush_pack(const PackBatch& B, cudaStream_t st) {
    if (B.n == 0) return I2SDF_OK;
    pack_batch_kernel<<<dim3(64, B.n), 256, 0, st>>>(B);
    I2SDF_CUDA_CHECK(cudaGetLastError());
    return I2SDF_OK;
}

static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

// ---- measurement hook ---------------------------------------------------------------------------
constexpr int kProfKinds = 8;
struct Prof {
    bool on = false;
    std::vector<cudaEvent_t> ev[kProfKinds];     // start/stop pairs
    size_t used[kProfKinds] = {};
    long long launches[kProfKinds] = {};
};
struct ProfScope {
    Prof* p; int kind; cudaStream_t st; int nlaunch;
    ProfScope(const i2sdf_handle* h, int kind_, cudaStream_t st_, int nlaunch_ = 1) : p((Prof*)h->prof), kind(kind_), st(st_), nlaunch(nlaunch_) {
        if (!p || !p->on) { p = nullptr; return; }
        if (p->used[kind] + 2 > p->ev[kind].size()) {
            cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
            p->ev[kind].push_back(a); p->ev[kind].push_back(b);
        }
        cudaEventRecord(p->ev[kind][p->used[kind]], st);
    }
    ~ProfScope() {
        if (!p) return;
        cudaEventRecord(p->ev[kind][p->used[kind] + 1], st);
        p->used[kind] += 2;
        p->launches[kind] += nlaunch;
    }
};

}  // namespace i2sdf

using namespace i2sdf;

namespace {
// Entry points run on the handle's device whatever the caller's current device is (and leave the caller's device current again):
// a module moved to cuda:1 in a process whose current device is cuda:0 must not launch its kernels on cuda:0.
struct DeviceScope {
    int prev = -1;
    explicit DeviceScope(const i2sdf_handle* h) {
        if (!h) return;
        int cur = -1;
        if (cudaGetDevice(&cur) == cudaSuccess && cur != h->device) { prev = cur; cudaSetDevice(h->device); }
    }
    ~DeviceScope() { if (prev >= 0) cudaSetDevice(prev); }
};
}  // namespace

extern "C" {

int i2sdf_abi_version(void) { return I2SDF_ABI_VERSION; }
const char* i2sdf_last_error(void) { return g_err; }

int i2sdf_create(const i2sdf_desc* d, int device, i2sdf_handle** out) {
    if (!d || !out) { set_error("null argument"); return I2SDF_E_INVALID; }
    if (d->abi_version != I2SDF_ABI_VERSION) { set_error("ABI version mismatch (%d vs %d)", d->abi_version, I2SDF_ABI_VERSION); return I2SDF_E_INVALID; }
    if (d->hidden != 256 || d->feature_size != 256) { set_error("only hidden=256 / feature_vector_size=256 networks are supported (got %d/%d)", d->hidden, d->feature_size); return I2SDF_E_INVALID; }
    const int ex = 3 + 6 * d->multires_x, ed = 3 + 6 * d->multires_d;
    if (d->n_sdf_layers < 3 || d->n_sdf_layers > kMaxLayers || d->n_color_layers < 2 || d->n_color_layers > kMaxLayers) { set_error("layer counts out of range"); return I2SDF_E_INVALID; }
    if (ex > 39 || ed > 39) { set_error("multires > 6 unsupported"); return I2SDF_E_INVALID; }
    if (d->sdf_skip_layer != -1 && (d->sdf_skip_layer < 1 || d->sdf_skip_layer > d->n_sdf_layers - 2)) { set_error("skip layer %d unsupported", d->sdf_skip_layer); return I2SDF_E_INVALID; }
    if (d->n_light_layers != 0 && (d->n_light_layers != 2 || d->light_hidden != 128)) { set_error("light head must be [256,128,1]"); return I2SDF_E_INVALID; }
    if (d->n_samples_eval != 128 || d->n_samples + 2 + d->n_samples_extra > 128 || d->max_total_iters > 5 || d->max_total_iters < 1) { set_error("sampler sizes unsupported (N_samples_eval must be 128, N_samples+2+N_extra <= 128, max_total_iters <= 5)"); return I2SDF_E_INVALID; }
    if (!d->u_up || !d->u_final || !d->t_init || !d->extra_idx) { set_error("sampler tables missing"); return I2SDF_E_INVALID; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { set_error("no CUDA device"); return I2SDF_E_NOGPU; }
    cudaDeviceProp prop;
    I2SDF_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) { set_error("device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor); return I2SDF_E_NOGPU; }
    I2SDF_CUDA_CHECK(cudaSetDevice(device));

    i2sdf_handle* h = (i2sdf_handle*)calloc(1, sizeof(i2sdf_handle));
    h->desc = *d;
    h->desc.u_up = h->desc.u_final = h->desc.t_init = nullptr; h->desc.extra_idx = nullptr;
    h->device = device;
    h->num_sms = prop.multiProcessorCount;
    NetDev& n = h->net;
    n.L = d->n_sdf_layers; n.skip = d->sdf_skip_layer; n.mx = d->multires_x; n.ex = ex; n.exp_ = round_up(ex, 8);
    n.Lc = d->n_color_layers; n.md = d->multires_d; n.ed = ed; n.Ll = d->n_light_layers; n.lh = d->light_hidden;
    // layer shapes in API order
    int li = 0;
    for (int l = 0; l < n.L; ++l, ++li) {
        int in = (l == 0) ? ex : 256;
        int outd = (l == n.L - 1) ? 257 : ((l + 1 == n.skip) ? 256 - ex : 256);
        h->lay_out[li] = outd; h->lay_in[li] = in;
    }
    for (int l = 0; l < n.Lc; ++l, ++li) { h->lay_out[li] = (l == n.Lc - 1) ? 3 : 256; h->lay_in[li] = (l == 0) ? 256 + ed : 256; }
    for (int l = 0; l < n.Ll; ++l, ++li) { h->lay_out[li] = (l == 0) ? n.lh : 1; h->lay_in[li] = (l == 0) ? 256 : n.lh; }
    h->n_layers = li;

    // pool layout
    size_t off = 0;
    auto take = [&](size_t nfloats) { size_t o = off; off += (nfloats + 63) / 64 * 64; return o; };
    size_t o_wt[kMaxLayers], o_wr[kMaxLayers], o_b[kMaxLayers], o_cwt[kMaxLayers], o_cb[kMaxLayers];
    for (int l = 0; l < n.L; ++l) {
        n.sdf_kpad[l] = (l == 0) ? n.exp_ : 256;
        o_wt[l] = take((size_t)n.sdf_kpad[l] * 256);
        o_wr[l] = take((size_t)256 * (l == 0 ? 64 : 256));
        o_b[l] = take(256);
    }
    size_t o_head = take(320);
    for (int l = 0; l < n.Lc - 1; ++l) {
        n.col_kpad[l] = (l == 0) ? round_up(256 + ed, 4) : 256;
        o_cwt[l] = take((size_t)n.col_kpad[l] * 256);
        o_cb[l] = take(256);
    }
    size_t o_chead = take(3 * 256 + 64);
    size_t o_lwt = take(256 * 128), o_lb = take(128), o_lhead = take(192);
    const int ne = d->n_samples_extra, mi = d->max_total_iters;
    size_t o_uup = take(d->n_samples_eval), o_ufin = take(d->n_samples), o_tin = take(d->n_samples_eval), o_eidx = take((size_t)mi * ne);
    h->pool_floats = off;
    if (cudaMalloc(&h->pool, off * sizeof(float)) != cudaSuccess) { free(h); set_error("cudaMalloc of %zu bytes failed", off * sizeof(float)); return I2SDF_E_CUDA; }
    cudaMemset(h->pool, 0, off * sizeof(float));
    for (int l = 0; l < n.L; ++l) { n.sdf_wt[l] = h->pool + o_wt[l]; n.sdf_wr[l] = h->pool + o_wr[l]; n.sdf_b[l] = h->pool + o_b[l]; }
    n.sdf_head = h->pool + o_head;
    for (int l = 0; l < n.Lc - 1; ++l) { n.col_wt[l] = h->pool + o_cwt[l]; n.col_b[l] = h->pool + o_cb[l]; }
    n.col_head = h->pool + o_chead;
    n.light_wt0 = h->pool + o_lwt; n.light_b0 = h->pool + o_lb; n.light_head = h->pool + o_lhead;
    SamplerDev& s = h->smp;
    s.n_samples = d->n_samples; s.n_eval = d->n_samples_eval; s.n_extra = ne; s.beta_iters = d->beta_iters; s.max_iters = mi;
    s.near_ = d->near_; s.far_ = d->far_; s.eps = d->eps; s.add_tiny = d->add_tiny; s.beta_min = d->beta_min;
    s.u_up = h->pool + o_uup; s.u_final = h->pool + o_ufin; s.t_init = h->pool + o_tin; s.extra_idx = reinterpret_cast<int*>(h->pool + o_eidx);
    cudaMemcpy(h->pool + o_uup, d->u_up, sizeof(float) * d->n_samples_eval, cudaMemcpyHostToDevice);
    cudaMemcpy(h->pool + o_ufin, d->u_final, sizeof(float) * d->n_samples, cudaMemcpyHostToDevice);
    cudaMemcpy(h->pool + o_tin, d->t_init, sizeof(float) * d->n_samples_eval, cudaMemcpyHostToDevice);
    cudaMemcpy(h->pool + o_eidx, d->extra_idx, sizeof(int) * mi * ne, cudaMemcpyHostToDevice);
    // The product library (libi2sdf_b200.so) has ONE backend: the tcgen05 chain kernels and the fused plane-slot backward.  The fp32 SIMT
    // kernel and the layer-by-layer backward exist only in the CHECK build (libi2sdf_b200_check.so, -DI2SDF_CHECK_BUILD: test
    // infrastructure, selected there by I2SDF_SIMT=1 / I2SDF_SIMT_MAIN=1 / I2SDF_FUSED_BWD=0) as the exact-fp32 cross-check.
#ifdef I2SDF_CHECK_BUILD
    const char* env = getenv("I2SDF_SIMT");
    h->use_tc = !(env && env[0] == '1');
    { const char* fe = getenv("I2SDF_FUSED_BWD"); h->fused = !(fe && fe[0] == '0'); }
#else
    h->use_tc = true;
    h->fused = true;
#endif
    h->tc = nullptr;
    h->prof = new Prof();
    h->tcmain = nullptr;
    if (h->use_tc) {
        int rc = tc_create(h);
        if (rc != I2SDF_OK) { cudaFree(h->pool); free(h); return rc; }
        // tc_create also sets h->tcmain when the full main pass is available (I2SDF_SIMT_MAIN=1 disables it)
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { set_error("create: %s", cudaGetErrorString(e)); cudaFree(h->pool); free(h); return I2SDF_E_CUDA; }
    *out = h;
    return I2SDF_OK;
}

int i2sdf_destroy(i2sdf_handle* h) {
    if (!h) return I2SDF_OK;
    if (h->tc) tc_destroy(h);
    if (h->prof) { Prof* p = (Prof*)h->prof; for (auto& v : p->ev) for (auto e : v) cudaEventDestroy(e); delete p; }
    cudaFree(h->pool);
    free(h);
    return I2SDF_OK;
}

int i2sdf_num_layers(const i2sdf_handle* h) { return h ? h->n_layers : 0; }
int i2sdf_uses_tensor_cores(const i2sdf_handle* h) { return h ? ((h->use_tc ? 1 : 0) | ((h->tcmain && tcmain_has_full(h->tcmain)) ? 2 : 0)) : 0; }

int i2sdf_pack_weights(i2sdf_handle* h, const float* const* W, const float* const* b, void* stream) {
    DeviceScope device_scope(h);
    if (!h || !W || !b) { set_error("null argument"); return I2SDF_E_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    const NetDev& n = h->net;
    int li = 0, rc;
    ProfScope ps(h, 3, st, 2);
    PackBatch PB{};
    for (int l = 0; l < n.L; ++l, ++li) {
        const int outd = h->lay_out[li], in = h->lay_in[li];
        const bool last = (l == n.L - 1);
        PackJob J{};
        J.mode = PACK_T; J.dst = (float*)n.sdf_wt[l]; J.rows = n.sdf_kpad[l]; J.ld = 256; J.src = W[li]; J.out = outd; J.in = in;
        J.row_off = last ? 1 : 0; J.nvalid = last ? 256 : outd;
        if ((rc = run_pack(PB, J))) return rc;
        PackJob R{};
        R.mode = PACK_R; R.dst = (float*)n.sdf_wr[l]; R.rows = 256; R.ld = (l == 0) ? 64 : 256; R.src = W[li]; R.out = last ? 1 : outd; R.in = in;
        if ((rc = run_pack(PB, R))) return rc;
        PackJob V{};
        V.mode = PACK_V; V.dst = (float*)n.sdf_b[l]; V.rows = 1; V.ld = 256; V.src = b[li]; V.row_off = last ? 1 : 0; V.nvalid = last ? 256 : outd;
        if ((rc = run_pack(PB, V))) return rc;
        if (last) {
            PackJob H{}; H.mode = PACK_V; H.dst = (float*)n.sdf_head; H.rows = 1; H.ld = 256; H.src = W[li]; H.row_off = 0; H.nvalid = 256;
            if ((rc = run_pack(PB, H))) return rc;
            PackJob HB{}; HB.mode = PACK_V; HB.dst = (float*)n.sdf_head + 256; HB.rows = 1; HB.ld = 1; HB.src = b[li]; HB.row_off = 0; HB.nvalid = 1;
            if ((rc = run_pack(PB, HB))) return rc;
        }
    }
    for (int l = 0; l < n.Lc; ++l, ++li) {
        const int outd = h->lay_out[li], in = h->lay_in[li];
        if (l < n.Lc - 1) {
            PackJob J{};
            J.mode = PACK_T; J.dst = (float*)n.col_wt[l]; J.rows = n.col_kpad[l]; J.ld = 256; J.src = W[li]; J.out = outd; J.in = in; J.nvalid = outd;
            if (l == 0) { J.feat_first = 256; J.ed = n.ed; }
            if ((rc = run_pack(PB, J))) return rc;
            PackJob V{}; V.mode = PACK_V; V.dst = (float*)n.col_b[l]; V.rows = 1; V.ld = 256; V.src = b[li]; V.nvalid = outd;
            if ((rc = run_pack(PB, V))) return rc;
        } else {
            PackJob H{}; H.mode = PACK_V; H.dst = (float*)n.col_head; H.rows = 1; H.ld = 768; H.src = W[li]; H.nvalid = 768;
            if ((rc = run_pack(PB, H))) return rc;
            PackJob HB{}; HB.mode = PACK_V; HB.dst = (float*)n.col_head + 768; HB.rows = 1; HB.ld = 3; HB.src = b[li]; HB.nvalid = 3;
            if ((rc = run_pack(PB, HB))) return rc;
        }
    }
    if (n.Ll == 2) {
        PackJob J{};
        J.mode = PACK_T; J.dst = (float*)n.light_wt0; J.rows = 256; J.ld = 128; J.src = W[li]; J.out = n.lh; J.in = 256; J.nvalid = n.lh;
        if ((rc = run_pack(PB, J))) return rc;
        PackJob V{}; V.mode = PACK_V; V.dst = (float*)n.light_b0; V.rows = 1; V.ld = 128; V.src = b[li]; V.nvalid = n.lh;
        if ((rc = run_pack(PB, V))) return rc;
        ++li;
        PackJob H{}; H.mode = PACK_V; H.dst = (float*)n.light_head; H.rows = 1; H.ld = 128; H.src = W[li]; H.nvalid = n.lh;
        if ((rc = run_pack(PB, H))) return rc;
        PackJob HB{}; HB.mode = PACK_V; HB.dst = (float*)n.light_head + n.lh; HB.rows = 1; HB.ld = 1; HB.src = b[li]; HB.nvalid = 1;
        if ((rc = run_pack(PB, HB))) return rc;
        ++li;
    }
    if ((rc = flush_pack(PB, st))) return rc;
    if (h->use_tc && (rc = tc_pack(h, W, b, st))) return rc;
    if (h->tcmain && (rc = tcmain_pack(h, h->tcmain, W, st))) return rc;
    return I2SDF_OK;
}

// workspace = [ sampler state | mlp scratch | per-sample temporaries (8 floats x R x 128) ]
static size_t ws_sampler_floats(const i2sdf_handle* h, int64_t R) { return (sampler_ws_floats(h, R) + 63) / 64 * 64; }
// per-CTA scratch of the eval main pass: one tile of h~_l per SDF hidden layer in plane format (and the same bytes as the fp32
// cross-check kernel's [2 CTAs per SM][L-1][64][256] floats)
static size_t ws_scratch_floats(const i2sdf_handle* h) { return ((size_t)h->num_sms * (size_t)(h->net.L - 1) * (planes::BIG_TILE / sizeof(float)) + 63) / 64 * 64; }

// Light-mask head behind the tensor-core main pass: the head is its own pass over the features the main pass writes
// (light_forward, backward.cu).  Workspace region = [ hidden pre-activations: rows x lh | features: rows x 256 ] for one
// chunk of at most kLightChunkRays rays (eval renders of more rays walk the chunks; training passes its own feature buffer).
constexpr int64_t kLightChunkRays = 4096;
static bool light_tc(const i2sdf_handle* h);
static int64_t light_chunk_rays(int64_t R) { return R < kLightChunkRays ? R : kLightChunkRays; }
static size_t ws_light_hidden_floats(const i2sdf_handle* h, int64_t R) { return (light_forward_ws_floats(h, light_chunk_rays(R) * 128) + 63) / 64 * 64; }
static size_t ws_light_floats(const i2sdf_handle* h, int64_t R) {
    if (!light_tc(h)) return 0;
    return ws_light_hidden_floats(h, R) + (size_t)light_chunk_rays(R) * 128 * 256;
}

size_t i2sdf_workspace_bytes(const i2sdf_handle* h, int64_t R, int training) {
    (void)training;
    if (!h || R < 0) return 0;
    return (ws_sampler_floats(h, R) + ws_scratch_floats(h) + (size_t)R * 128 * 8 + 64 + ws_light_floats(h, R)) * sizeof(float);
}
// light head over rows [0, M) of feat, in chunks that fit the workspace's hidden buffer
static int run_light(const i2sdf_handle* h, long long M, const float* feat, float* s_light, float* hidden, long long hidden_rows, cudaStream_t st) {
    ProfScope ps(h, 6, st, 2);
    for (long long m0 = 0; m0 < M; m0 += hidden_rows) {
        const long long mc = (M - m0 < hidden_rows) ? M - m0 : hidden_rows;
        int rc = light_forward(h, mc, feat + (size_t)m0 * 256, s_light + m0, hidden, st);
        if (rc) return rc;
    }
    return I2SDF_OK;
}

int i2sdf_rays(i2sdf_handle* h, const float* uv, const float* pose, const float* intr, int B, int P, float* o, float* d,
               float* dnorm, void* stream) {
    DeviceScope device_scope(h);
    if (!h || !uv || !pose || !intr || !o || !d || !dnorm) { set_error("null argument"); return I2SDF_E_INVALID; }
    ProfScope ps(h, 3, (cudaStream_t)stream);
    return launch_rays(uv, pose, intr, B, P, o, d, dnorm, (cudaStream_t)stream);
}

static int run_mlp(const i2sdf_handle* h, const MlpParams& p, cudaStream_t st) {
    const bool sdf_only = !p.out_feat && !p.out_grad && !p.want_color && !p.want_light && !p.save_act && !p.sl.base;
    ProfScope ps(h, sdf_only ? 0 : 1, st);
    if (h->use_tc && sdf_only) return tc_launch_sdf(h, p, st);
    // eval main pass (sdf + grad_x + rgb per sample, nothing saved) -> tensor-core kernel
    // (training: pre-activations + features are saved for the backward instead of the per-CTA scratch)
    // (a light head rides behind it as its own pass over p.out_feat: the callers clear want_light, see run_light)
    if (h->tcmain && tcmain_has_full(h->tcmain) && p.out_sdf && p.out_grad && p.out_rgb && p.want_color && !p.want_light &&
        (p.scratch || p.sl.base) && !p.save_act && (p.ray_d || p.pts))
        return tcmain_launch(h, h->tcmain, p, st);
    // sdf + features only (ImplicitNetwork.forward: mesh extraction, plots): F layers + feature layer; needs no scratch
    if (h->tcmain && p.out_sdf && p.out_feat && !p.out_grad && !p.want_color && !p.want_light && !p.save_act && !p.sl.base && (p.ray_d || p.pts))
        return tcmain_launch(h, h->tcmain, p, st);
    // sdf + grad_x only (eikonal points / ImplicitNetwork.gradient): F layers then the reverse sweep, no radiance stack
    if (h->tcmain && p.out_sdf && p.out_grad && !p.want_color && !p.want_light && !p.out_feat && (p.scratch || p.sl.base) && !p.save_act && (p.ray_d || p.pts))
        return tcmain_launch(h, h->tcmain, p, st);
    if (p.sl.base) { set_error("run_mlp: plane-slot save requested for a pass the tensor-core kernel does not cover"); return I2SDF_E_INVALID; }
#ifdef I2SDF_CHECK_BUILD
    return launch_mlp_simt(h, p, st);
#else
    set_error("this call is not covered by the tensor-core chain kernels (network too deep for the full chain: n_sdf_layers - 1 + n_color_layers > 13, "
              "or an output combination no reference call site uses); the product library has no second backend");
    return I2SDF_E_INVALID;
#endif
}

// format of the state a forward saves for the backward: 1 = plane slots (tensor-core chain kernels + fused backward),
// 0 = fp32 pre-activations [L-1][M][256] (fp32 kernels + the layer-by-layer backward).  kind 0: main pass, 1: SDF points
static bool planes_main(const i2sdf_handle* h) { return h->fused && h->tcmain && tcmain_has_full(h->tcmain); }
static bool planes_sdf(const i2sdf_handle* h) { return h->fused && h->tcmain != nullptr; }
static size_t planes_main_bytes(const i2sdf_handle* h, int64_t M) { return planes::make_layout(M, h->net.L - 1, h->net.Lc, true, nullptr, nullptr).saved_total(); }
static bool light_tc(const i2sdf_handle* h) { return h->net.Ll == 2 && h->use_tc && h->tcmain && tcmain_has_full(h->tcmain); }
int i2sdf_saved_format(const i2sdf_handle* h, int kind) { return h ? ((kind == 0 ? planes_main(h) : planes_sdf(h)) ? 1 : 0) : 0; }

size_t i2sdf_sdf_saved_bytes(const i2sdf_handle* h, int64_t M) {
    if (!h || M < 0) return 0;
    if (planes_sdf(h)) return planes::make_layout(M, h->net.L - 1, h->net.Lc, false, nullptr, nullptr).saved_total();
    return (size_t)(h->net.L - 1) * (size_t)M * 256 * sizeof(float);
}

int i2sdf_sdf_forward(i2sdf_handle* h, const float* pts, int64_t M, float* out_sdf, float* out_feat, float* out_grad,
                      float* save_act, void* workspace, size_t workspace_bytes, void* stream) {
    DeviceScope device_scope(h);
    if (!h) { set_error("null handle"); return I2SDF_E_INVALID; }
    if (M == 0) return I2SDF_OK;
    if (M < 0 || !pts || !out_sdf) { set_error("null argument"); return I2SDF_E_INVALID; }
    MlpParams p{};
    p.pts = pts; p.M = M; p.ns = 1; p.round_idx = -1; p.beta_min = h->smp.beta_min;
    p.out_sdf = out_sdf; p.out_feat = out_feat; p.out_grad = out_grad;
    if (save_act && planes_sdf(h) && out_grad && !out_feat) p.sl = planes::make_layout(M, h->net.L - 1, h->net.Lc, false, save_act, nullptr);
    else p.save_act = save_act;
    if (out_grad && !save_act) {
        if (!workspace || workspace_bytes < ws_scratch_floats(h) * sizeof(float)) { set_error("sdf_forward: workspace too small"); return I2SDF_E_WORKSPACE; }
        p.scratch = (float*)workspace;
    }
    // development probe: the sdf-only tensor-core kernel stamps its hand-offs into the first 64 KB of the workspace
    if (!out_grad && !out_feat && !save_act && workspace && workspace_bytes >= 65536 && getenv("I2SDF_DEBUG_TIMELINE")) p.scratch = (float*)workspace;
    p.net = h->net;
    return run_mlp(h, p, (cudaStream_t)stream);
}

int i2sdf_sdf_grid(i2sdf_handle* h, const float* gx, const float* gy, const float* gz, int nx, int ny, int nz, const float* affine,
                   float* out_sdf, void* stream) {
    DeviceScope device_scope(h);
    if (!h || !gx || !gy || !gz || !out_sdf) { set_error("null argument"); return I2SDF_E_INVALID; }
    if (nx < 1 || ny < 1 || nz < 1) { set_error("sdf_grid: empty grid"); return I2SDF_E_INVALID; }
    if (!h->use_tc) { set_error("sdf_grid: the tensor-core sdf kernel is not available for this handle"); return I2SDF_E_INVALID; }
    MlpParams p{};
    p.gx = gx; p.gy = gy; p.gz = gz; p.nx = nx; p.ny = ny; p.nz = nz; p.grid_affine = affine;
    p.M = (long long)nx * ny * nz; p.ns = 1; p.round_idx = -1; p.beta_min = h->smp.beta_min;
    p.out_sdf = out_sdf;
    p.net = h->net;
    return run_mlp(h, p, (cudaStream_t)stream);
}

int i2sdf_sampler_rounds(i2sdf_handle* h, const float* o, const float* d, int64_t R, const float* beta_param,
                         const float* jitter, const float* u_final, void* workspace, size_t workspace_bytes, void* stream) {
    DeviceScope device_scope(h);
    if (!h || !o || !d || !beta_param || !workspace) { set_error("null argument"); return I2SDF_E_INVALID; }
    if (workspace_bytes < i2sdf_workspace_bytes(h, R, 0)) { set_error("sampler: workspace too small"); return I2SDF_E_WORKSPACE; }
    cudaStream_t st = (cudaStream_t)stream;
    SamplerWs W = carve_sampler_ws(h, R, (float*)workspace);
    int rc;
    { ProfScope ps(h, 2, st); if ((rc = launch_sampler_init(h, W, R, jitter, h->desc.lemma2_coeff, st))) return rc; }
    for (int k = 0; k < h->smp.max_iters; ++k) {
        MlpParams p{};
        p.ray_o = o; p.ray_d = d; p.zarr = W.samples; p.zstride = h->smp.n_eval; p.ns = h->smp.n_eval;
        p.M = (long long)R * h->smp.n_eval; p.out_sdf = W.sdf_new;
        p.beta_max = W.beta_max; p.beta_param = beta_param; p.beta_min = h->smp.beta_min; p.round_idx = k;
        p.net = h->net;
        if ((rc = run_mlp(h, p, st))) return rc;
        ProfScope ps(h, 2, st, 1);
        if ((rc = launch_sampler_round(h, W, R, k, 0, beta_param, nullptr, st))) return rc;
    }
    // the final samples of whichever round turned out to be the last (decided on the device)
    ProfScope ps(h, 2, st, 1);
    return launch_sampler_round(h, W, R, -1, 1, beta_param, u_final, st);
}

int i2sdf_sampler_step(i2sdf_handle* h, const float* o, const float* d, int64_t R, const float* beta_param, const float* jitter,
                       const float* u_final, int stage, int k, void* workspace, size_t workspace_bytes, void* stream) {
    DeviceScope device_scope(h);
    if (!h || !o || !d || !beta_param || !workspace) { set_error("null argument"); return I2SDF_E_INVALID; }
    if (workspace_bytes < i2sdf_workspace_bytes(h, R, 0)) { set_error("sampler: workspace too small"); return I2SDF_E_WORKSPACE; }
    if (stage < 0 || stage > 2 || (stage > 0 && (k < 0 || k >= h->smp.max_iters))) { set_error("sampler_step: bad stage / round"); return I2SDF_E_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    SamplerWs W = carve_sampler_ws(h, R, (float*)workspace);
    if (stage == 0) { ProfScope ps(h, 2, st); return launch_sampler_init(h, W, R, jitter, h->desc.lemma2_coeff, st); }
    if (stage == 1) {
        MlpParams p{};
        p.ray_o = o; p.ray_d = d; p.zarr = W.samples; p.zstride = h->smp.n_eval; p.ns = h->smp.n_eval;
        p.M = (long long)R * h->smp.n_eval; p.out_sdf = W.sdf_new;
        p.beta_max = W.beta_max; p.beta_param = beta_param; p.beta_min = h->smp.beta_min; p.round_idx = k;
        p.net = h->net;
        int rc = run_mlp(h, p, st);
        if (rc) return rc;
        ProfScope ps(h, 2, st, 1);
        return launch_sampler_round(h, W, R, k, 0, beta_param, nullptr, st);
    }
    ProfScope ps(h, 2, st, 1);
    return launch_sampler_round(h, W, R, k, 1, beta_param, u_final, st);
}

float* i2sdf_sampler_beta_max(i2sdf_handle* h, int64_t R, void* workspace) {
    if (!h || !workspace) return nullptr;
    return carve_sampler_ws(h, R, (float*)workspace).beta_max;
}

int i2sdf_sampler_finalize(i2sdf_handle* h, int64_t R, const float* beta_param, const int32_t* extra_idx, const int32_t* eik_idx,
                           float* out_z, float* out_z_eik, int32_t* out_info, void* workspace, size_t workspace_bytes, void* stream) {
    DeviceScope device_scope(h);
    if (!h || !beta_param || !out_z || !workspace) { set_error("null argument"); return I2SDF_E_INVALID; }
    if (workspace_bytes < i2sdf_workspace_bytes(h, R, 0)) { set_error("sampler: workspace too small"); return I2SDF_E_WORKSPACE; }
    SamplerWs W = carve_sampler_ws(h, R, (float*)workspace);
    ProfScope ps(h, 2, (cudaStream_t)stream);
    return launch_sampler_finalize(h, W, R, beta_param, extra_idx, 0, eik_idx, out_z, out_z_eik, out_info, (cudaStream_t)stream);
}

int i2sdf_sampler_finalize_candidates(i2sdf_handle* h, int64_t R, const float* beta_param, const int32_t* extra_table, const int32_t* eik_idx,
                                      float* out_z, float* out_z_eik, int32_t* out_info, void* workspace, size_t workspace_bytes, void* stream) {
    DeviceScope device_scope(h);
    if (!h || !beta_param || !out_z || !workspace || !extra_table) { set_error("null argument"); return I2SDF_E_INVALID; }
    if (workspace_bytes < i2sdf_workspace_bytes(h, R, 0)) { set_error("sampler: workspace too small"); return I2SDF_E_WORKSPACE; }
    SamplerWs W = carve_sampler_ws(h, R, (float*)workspace);
    ProfScope ps(h, 2, (cudaStream_t)stream);
    return launch_sampler_finalize(h, W, R, beta_param, extra_table, 1, eik_idx, out_z, out_z_eik, out_info, (cudaStream_t)stream);
}

__global__ void sampler_info_kernel(SamplerDev S, const float* beta_max, const float* beta_param, int* info) {
    float b0 = fabsf(*beta_param) + S.beta_min;
    int klast = 0;
    while (klast + 1 < S.max_iters && (beta_max[klast] > b0)) ++klast;
    info[0] = klast + 1;
    info[1] = S.n_eval * (klast + 1);
}

int i2sdf_sampler_info(i2sdf_handle* h, int64_t R, const float* beta_param, int32_t* out_info, void* workspace,
                       size_t workspace_bytes, void* stream) {
    DeviceScope device_scope(h);
    if (!h || !beta_param || !out_info || !workspace) { set_error("null argument"); return I2SDF_E_INVALID; }
    if (workspace_bytes < i2sdf_workspace_bytes(h, R, 0)) { set_error("sampler: workspace too small"); return I2SDF_E_WORKSPACE; }
    SamplerWs W = carve_sampler_ws(h, R, (float*)workspace);
    sampler_info_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(h->smp, W.beta_max, beta_param, out_info);
    I2SDF_CUDA_CHECK(cudaGetLastError());
    return I2SDF_OK;
}

int i2sdf_sampler_round_debug(i2sdf_handle* h, const float* z, const float* sdf, int64_t R, int n, const float* beta_param,
                              const float* beta_in, int force_upsample, const float* u_tape, float* out_beta, float* out_cdf,
                              int32_t* out_inds, float* out_samples, float* out_z_merged, int32_t* out_src, void* stream) {
    DeviceScope device_scope(h);
    if (!h || !z || !sdf || !beta_param || !beta_in || !out_samples) { set_error("null argument"); return I2SDF_E_INVALID; }
    if (n < 2 || n > h->smp.n_eval * h->smp.max_iters) { set_error("round_debug: n=%d out of range", n); return I2SDF_E_INVALID; }
    return launch_sampler_round_debug(h, z, sdf, R, n, beta_param, beta_in, force_upsample, u_tape, out_beta, out_cdf, out_inds,
                                      out_samples, out_z_merged, out_src, (cudaStream_t)stream);
}

size_t i2sdf_backward_workspace_bytes(const i2sdf_handle* h, int64_t M) {
    if (!h || M < 0) return 0;
    size_t a = sdf_backward_ws_floats(h, M), b = color_backward_ws_floats(h, M), c = light_backward_ws_floats(h, M);
    size_t m = a > b ? a : b;
    m = m > c ? m : c;
    m = (m + 64) * sizeof(float);
    if (h->tcmain) { const size_t f = fused_backward_ws_bytes(h, M, planes_main(h)) + 256; m = m > f ? m : f; }
    return m;
}

int i2sdf_points_forward(i2sdf_handle* h, const float* o, const float* d, const float* z, int64_t R, int N, float* s_sdf, float* s_grad,
                         float* s_rgb, float* s_light, float* s_feat, float* save_act, void* workspace, size_t workspace_bytes, void* stream) {
    DeviceScope device_scope(h);
    return i2sdf_points_forward_ex(h, o, d, z, R, N, nullptr, 0, s_sdf, s_grad, s_rgb, s_light, s_feat, save_act, workspace, workspace_bytes, stream);
}

int i2sdf_points_forward_ex(i2sdf_handle* h, const float* o, const float* d, const float* z, int64_t R, int N, const float* extra_pts,
                            int64_t n_extra, float* s_sdf, float* s_grad, float* s_rgb, float* s_light, float* s_feat, float* save_act,
                            void* workspace, size_t workspace_bytes, void* stream) {
    DeviceScope device_scope(h);
    if (!h || !o || !d || !z || !s_sdf) { set_error("null argument"); return I2SDF_E_INVALID; }
    if (n_extra < 0 || (n_extra > 0 && (!extra_pts || !planes_main(h) || !s_rgb || !s_grad))) {
        set_error("points_forward_ex: extra points need the tensor-core main pass with s_rgb and s_grad"); return I2SDF_E_INVALID;
    }
    if (N < 1 || N > 128) { set_error("points_forward: N=%d unsupported", N); return I2SDF_E_INVALID; }
    if (s_light && h->net.Ll == 0) { set_error("points_forward: no light head"); return I2SDF_E_INVALID; }
    MlpParams p{};
    p.ray_o = o; p.ray_d = d; p.zarr = z; p.zstride = N + 1; p.ns = N; p.M = (long long)R * N + n_extra; p.round_idx = -1; p.beta_min = h->smp.beta_min;
    if (n_extra > 0) { p.pts = extra_pts; p.m_rays = (long long)R * N; }     // appended explicit points (eikonal / smoothness) ride the same launch
    // tensor-core main pass + light head: the head is a second pass over the features (s_feat, rows of the R*N ray samples)
    const bool ltc = s_light && light_tc(h) && s_rgb && s_grad && (!save_act || planes_main(h));
    if (ltc && (!s_feat || !workspace || workspace_bytes < i2sdf_workspace_bytes(h, R, 0))) {
        set_error("points_forward: the light head on the tensor-core path needs s_feat and the full workspace"); return I2SDF_E_WORKSPACE;
    }
    p.out_sdf = s_sdf; p.out_grad = s_grad; p.out_rgb = s_rgb; p.out_light = ltc ? nullptr : s_light; p.out_feat = s_feat;
    p.want_color = s_rgb != nullptr; p.want_light = (s_light != nullptr) && !ltc;
    if (save_act && planes_main(h) && (!s_light || ltc) && s_rgb && s_grad) p.sl = planes::make_layout(p.M, h->net.L - 1, h->net.Lc, true, save_act, nullptr);
    else p.save_act = save_act;
    if (s_grad && !save_act) {
        if (!workspace || workspace_bytes < ws_scratch_floats(h) * sizeof(float)) { set_error("points_forward: workspace too small"); return I2SDF_E_WORKSPACE; }
        p.scratch = (float*)workspace;
    }
    p.net = h->net;
    int rc = run_mlp(h, p, (cudaStream_t)stream);
    if (rc || !ltc) return rc;
    if (p.sl.base)      // training: the hidden pre-activations stay in the saved state for i2sdf_light_backward
        return run_light(h, (long long)R * N, s_feat, s_light, (float*)((uint8_t*)save_act + planes_main_bytes(h, p.M)), p.M, (cudaStream_t)stream);
    float* hidden = (float*)workspace + ws_sampler_floats(h, R) + ws_scratch_floats(h) + (size_t)R * 128 * 8 + 64;
    return run_light(h, (long long)R * N, s_feat, s_light, hidden, light_chunk_rays(R) * 128, (cudaStream_t)stream);
}

int i2sdf_composite_forward(i2sdf_handle* h, const float* z, const float* dnorm, const float* s_sdf, const float* s_rgb, const float* s_grad,
                            const float* s_light, const float* beta_param, int64_t R, int N, float* rgb, float* depth, float* weight_sum,
                            float* normal, float* light, float* s_w, void* stream) {
    DeviceScope device_scope(h);
    if (!h || !z || !dnorm || !s_sdf || !beta_param) { set_error("null argument"); return I2SDF_E_INVALID; }
    ProfScope ps(h, 3, (cudaStream_t)stream);
    return launch_composite(h, z, dnorm, s_sdf, s_rgb, s_grad, s_light, beta_param, R, N, rgb, depth, weight_sum, normal, light, s_w, (cudaStream_t)stream);
}

int i2sdf_composite_backward(i2sdf_handle* h, const float* z, const float* dnorm, const float* s_sdf, const float* s_rgb, const float* s_grad,
                             const float* s_light, const float* beta_param, int64_t R, int N, const float* g_rgb, const float* g_depth,
                             const float* g_wsum, const float* g_normal, const float* g_light, float* o_sdf, float* o_rgb, float* o_grad,
                             float* o_light, float* o_beta, void* stream) {
    DeviceScope device_scope(h);
    if (!h || !z || !dnorm || !s_sdf || !beta_param || !o_sdf) { set_error("null argument"); return I2SDF_E_INVALID; }
    if ((g_rgb && !s_rgb) || (g_normal && !s_grad)) { set_error("composite_backward: missing forward tensors"); return I2SDF_E_INVALID; }
    ProfScope ps(h, 3, (cudaStream_t)stream);
    return launch_composite_backward(h, z, dnorm, s_sdf, s_rgb, s_grad, s_light, beta_param, R, N, g_rgb, g_depth, g_wsum, g_normal, g_light,
                                     o_sdf, o_rgb, o_grad, o_light, o_beta, (cudaStream_t)stream);
}

int i2sdf_color_backward(i2sdf_handle* h, const float* const* W, const float* const* b, const float* dirs, int ns, const float* feat,
                         const float* s_rgb, const float* g_rgb, int64_t M, float* const* dW, float* const* db, float* g_x, void* workspace,
                         size_t workspace_bytes, void* stream) {
    DeviceScope device_scope(h);
    if (!h || !W || !b || !dirs || !feat || !s_rgb || !g_rgb || !dW || !db || !g_x || !workspace) { set_error("null argument"); return I2SDF_E_INVALID; }
    if (workspace_bytes < i2sdf_backward_workspace_bytes(h, M)) { set_error("color_backward: workspace too small"); return I2SDF_E_WORKSPACE; }
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope ps(h, 1, st, 4 * h->net.Lc);
    float* gf = nullptr; int ld = 0;
    int rc = color_backward(h, M, ns, dirs, W, b, feat, s_rgb, g_rgb, dW, db, (float*)workspace, &gf, &ld, st);
    if (rc) return rc;
    // the adjoint of the stack input lives at the head of the workspace's second [M][288] block: copy it out
    I2SDF_CUDA_CHECK(cudaMemcpyAsync(g_x, gf - h->net.ed, (size_t)M * 288 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return I2SDF_OK;
}

int i2sdf_light_backward(i2sdf_handle* h, const float* const* W, const float* const* b, const float* feat, const float* hidden, const float* s_light,
                         const float* g_light, int64_t M, float* const* dW, float* const* db, void* workspace, size_t workspace_bytes, void* stream) {
    DeviceScope device_scope(h);
    if (!h || !W || !b || !feat || !s_light || !g_light || !dW || !db || !workspace) { set_error("null argument"); return I2SDF_E_INVALID; }
    if (h->net.Ll != 2) { set_error("light_backward: network has no light head"); return I2SDF_E_INVALID; }
    if (workspace_bytes < i2sdf_backward_workspace_bytes(h, M)) { set_error("light_backward: workspace too small"); return I2SDF_E_WORKSPACE; }
    ProfScope ps(h, 6, (cudaStream_t)stream, 8);
    return light_backward(h, M, W, b, feat, hidden, s_light, g_light, dW, db, (float*)workspace, (cudaStream_t)stream);
}

int i2sdf_sdf_backward(i2sdf_handle* h, const float* const* W, const float* pts, const float* o, const float* d, const float* z, int zstride,
                       int ns, int64_t M, const float* act, const float* g_sdf, const float* g_feat, int g_feat_ld, const float* g_grad,
                       float* const* dW, float* const* db, void* workspace, size_t workspace_bytes, void* stream) {
    DeviceScope device_scope(h);
    if (!h || !W || !act || !dW || !db || !workspace || (!pts && (!o || !d || !z))) { set_error("null argument"); return I2SDF_E_INVALID; }
    if (workspace_bytes < i2sdf_backward_workspace_bytes(h, M)) { set_error("sdf_backward: workspace too small"); return I2SDF_E_WORKSPACE; }
    bwd::PointSrc S{pts, o, d, z, zstride, ns > 0 ? ns : 1};
    ProfScope ps(h, 1, (cudaStream_t)stream, 10 * h->net.L);
    return sdf_backward(h, S, M, W, act, g_sdf, g_feat, g_feat_ld, g_grad, dW, db, (float*)workspace, (cudaStream_t)stream);
}

int i2sdf_profile_enable(i2sdf_handle* h, int enable) {
    if (!h) { set_error("null handle"); return I2SDF_E_INVALID; }
    Prof* p = (Prof*)h->prof;
    p->on = enable != 0;
    if (enable) for (int k = 0; k < kProfKinds; ++k) { p->used[k] = 0; p->launches[k] = 0; }
    return I2SDF_OK;
}

int i2sdf_profile_read(i2sdf_handle* h, float ms[4], int64_t launches[4]) { return i2sdf_profile_read_n(h, 4, ms, launches); }

int i2sdf_profile_read_n(i2sdf_handle* h, int n, float* ms, int64_t* launches) {
    if (!h || !ms || !launches || n < 1 || n > kProfKinds) { set_error("profile_read: bad argument"); return I2SDF_E_INVALID; }
    Prof* p = (Prof*)h->prof;
    I2SDF_CUDA_CHECK(cudaDeviceSynchronize());
    for (int k = 0; k < n; ++k) {
        float tot = 0.f;
        for (size_t i = 0; i + 1 < p->used[k]; i += 2) { float t = 0.f; cudaEventElapsedTime(&t, p->ev[k][i], p->ev[k][i + 1]); tot += t; }
        ms[k] = tot;
        launches[k] = p->launches[k];
    }
    return I2SDF_OK;
}

int64_t i2sdf_debug_bwd_timeline(int64_t* out, int64_t n) {
    cudaDeviceSynchronize();
    return (int64_t)tc_bwd_timeline_read((long long*)out, (long long)n);
}

size_t i2sdf_saved_bytes(const i2sdf_handle* h, int64_t R, int N) { return i2sdf_saved_bytes_points(h, R * N); }

// plane slots of the main pass (+ for networks with a light head: the head's hidden pre-activations [M][lh] fp32 behind them)
size_t i2sdf_saved_bytes_points(const i2sdf_handle* h, int64_t M) {
    if (!h || M < 0) return 0;
    if (planes_main(h)) return planes_main_bytes(h, M) + (light_tc(h) ? (size_t)M * h->net.lh * sizeof(float) : 0);
    return (size_t)(h->net.L - 1) * (size_t)M * 256 * sizeof(float);
}

int i2sdf_fused_backward_ex(i2sdf_handle* h, const float* pts, const float* o, const float* d, const float* z, int zstride, int ns, int64_t M,
                            int64_t m_rays, void* saved, const float* s_rgb, const float* g_sdf, const float* g_grad, const float* g_rgb, int64_t m_up,
                            const float* g_grad_tail, float* const* dW_sdf, float* const* db_sdf, float* const* dW_col, float* const* db_col,
                            void* workspace, size_t workspace_bytes, void* stream) {
    DeviceScope device_scope(h);
    if (!h || !saved || !dW_sdf || !db_sdf || !workspace || (!pts && (!o || !d || !z))) { set_error("fused_backward: null argument"); return I2SDF_E_INVALID; }
    if (pts && o && (m_rays < 0 || m_rays > M)) { set_error("fused_backward: m_rays out of range"); return I2SDF_E_INVALID; }
    if (m_up < 0 || m_up > M || (g_grad_tail && m_up == M)) { set_error("fused_backward: m_up out of range"); return I2SDF_E_INVALID; }
    if (g_rgb && (!s_rgb || !dW_col || !db_col)) { set_error("fused_backward: g_rgb needs s_rgb, dW_col, db_col"); return I2SDF_E_INVALID; }
    if (!(g_rgb ? planes_main(h) : planes_sdf(h))) { set_error("fused_backward: this handle saves fp32 pre-activations (use i2sdf_sdf_backward / i2sdf_color_backward)"); return I2SDF_E_INVALID; }
    if (workspace_bytes < i2sdf_backward_workspace_bytes(h, M)) { set_error("fused_backward: workspace too small"); return I2SDF_E_WORKSPACE; }
    bwd::PointSrc S{pts, o, d, z, zstride, ns > 0 ? ns : 1};
    cudaStream_t st = (cudaStream_t)stream;
    int rc;
    { ProfScope ps(h, 4, st, 1); if ((rc = fused_backward(h, S, M, saved, s_rgb, g_sdf, g_grad, g_rgb, dW_sdf, db_sdf, dW_col, db_col, workspace, 1, m_rays, st, m_up, g_grad_tail))) return rc; }
    { ProfScope ps(h, 5, st, 1); if ((rc = fused_backward(h, S, M, saved, s_rgb, g_sdf, g_grad, g_rgb, dW_sdf, db_sdf, dW_col, db_col, workspace, 2, m_rays, st, m_up, g_grad_tail))) return rc; }
    { ProfScope ps(h, 3, st, g_rgb ? 4 : 2); if ((rc = fused_backward(h, S, M, saved, s_rgb, g_sdf, g_grad, g_rgb, dW_sdf, db_sdf, dW_col, db_col, workspace, 4, m_rays, st, m_up, g_grad_tail))) return rc; }
    return I2SDF_OK;
}

int i2sdf_fused_backward(i2sdf_handle* h, const float* pts, const float* o, const float* d, const float* z, int zstride, int ns, int64_t M,
                         int64_t m_rays, void* saved, const float* s_rgb, const float* g_sdf, const float* g_grad, const float* g_rgb, float* const* dW_sdf,
                         float* const* db_sdf, float* const* dW_col, float* const* db_col, void* workspace, size_t workspace_bytes, void* stream) {
    return i2sdf_fused_backward_ex(h, pts, o, d, z, zstride, ns, M, m_rays, saved, s_rgb, g_sdf, g_grad, g_rgb, M, nullptr, dW_sdf, db_sdf, dW_col, db_col,
                                   workspace, workspace_bytes, stream);
}

int i2sdf_render_forward(i2sdf_handle* h, const float* o, const float* d, const float* dnorm, const float* z, int64_t R, int N,
                         const float* beta_param, float* rgb, float* depth, float* weight_sum, float* normal, float* light,
                         float* s_sdf, float* s_grad, float* s_rgb, float* s_w, float* s_light, void* save, size_t save_bytes,
                         void* workspace, size_t workspace_bytes, void* stream) {
    DeviceScope device_scope(h);
    if (!h || !o || !d || !dnorm || !z || !beta_param || !workspace) { set_error("null argument"); return I2SDF_E_INVALID; }
    if (N < 1 || N > 128) { set_error("render_forward: N=%d samples per ray unsupported (1..128)", N); return I2SDF_E_INVALID; }
    if (workspace_bytes < i2sdf_workspace_bytes(h, R, 0)) { set_error("render_forward: workspace too small"); return I2SDF_E_WORKSPACE; }
    if (light && h->net.Ll == 0) { set_error("render_forward: light output requested but the network has no light head"); return I2SDF_E_INVALID; }
    if (save && save_bytes < i2sdf_saved_bytes(h, R, N)) { set_error("render_forward: save buffer too small"); return I2SDF_E_WORKSPACE; }
    cudaStream_t st = (cudaStream_t)stream;
    float* base = (float*)workspace;
    float* scratch = base + ws_sampler_floats(h, R);
    float* tmp = scratch + ws_scratch_floats(h);
    const size_t RN = (size_t)R * N;
    if (!s_sdf) s_sdf = tmp;
    if (!s_grad && normal) s_grad = tmp + (size_t)R * 128;
    if (!s_rgb && rgb) s_rgb = tmp + (size_t)R * 128 * 4;
    if (!s_light && light) s_light = tmp + (size_t)R * 128 * 7;
    (void)RN;
    // tensor-core main pass + light head: walk the rays in chunks; each chunk's features go to the workspace and the head
    // (light_forward) turns them into the per-sample light mask
    const bool ltc = s_light && light_tc(h) && s_rgb && s_grad && (!save || planes_main(h));
    const int64_t CH = ltc ? light_chunk_rays(R) : R;
    if (ltc && save && R > CH) { set_error("render_forward: saving state with a light head is limited to %lld rays per call", (long long)CH); return I2SDF_E_INVALID; }
    float* lhidden = tmp + (size_t)R * 128 * 8 + 64;
    float* lfeat = lhidden + ws_light_hidden_floats(h, R);
    int rc = I2SDF_OK;
    for (int64_t r0 = 0; r0 < R && rc == I2SDF_OK; r0 += (CH > 0 ? CH : 1)) {
        const int64_t rc_rays = (R - r0 < CH) ? R - r0 : CH;
        MlpParams p{};
        p.ray_o = o + r0 * 3; p.ray_d = d + r0 * 3; p.zarr = z + r0 * (N + 1); p.zstride = N + 1; p.ns = N; p.M = (long long)rc_rays * N; p.round_idx = -1;
        p.beta_min = h->smp.beta_min;
        p.out_sdf = s_sdf + r0 * N; p.out_grad = s_grad ? s_grad + r0 * N * 3 : nullptr; p.out_rgb = s_rgb ? s_rgb + r0 * N * 3 : nullptr;
        p.out_light = ltc ? nullptr : s_light; p.out_feat = ltc ? lfeat : nullptr;
        p.want_color = s_rgb != nullptr; p.want_light = (s_light != nullptr) && !ltc;
        if (save && planes_main(h) && (!s_light || ltc) && s_rgb && s_grad) p.sl = planes::make_layout(p.M, h->net.L - 1, h->net.Lc, true, save, nullptr);
        else p.save_act = (float*)save;
        p.scratch = scratch;
        // development probe: the tensor-core main pass stamps its hand-offs into the first 64 KB of the workspace (the sampler's state,
        // dead by now)
        if (!save && getenv("I2SDF_DEBUG_TIMELINE")) p.tl = workspace;
        p.net = h->net;
        rc = run_mlp(h, p, st);
        if (rc == I2SDF_OK && ltc) rc = run_light(h, p.M, lfeat, s_light + r0 * N, lhidden, CH * 128, st);
    }
    if (rc) return rc;
    ProfScope ps(h, 3, st);
    return launch_composite(h, z, dnorm, s_sdf, s_rgb, s_grad, s_light, beta_param, R, N, rgb, depth, weight_sum, normal, light, s_w, st);
}

// ---- plane slots (diagnostics / tests) --------------------------------------------------------------------------
size_t i2sdf_planes_slot_bytes(int64_t M, int columns) {
    if (M < 0) return 0;
    return columns == 256 ? planes::big_slot_bytes(M) : (columns == 48 ? planes::small_slot_bytes(M) : 0);
}
static int planes_chunks(int columns) { return columns == 256 ? planes::BIG_CHUNKS : (columns == 48 ? planes::SMALL_CHUNKS : 0); }

int i2sdf_planes_pack(i2sdf_handle* h, const float* X, int ld, int width, int64_t M, int columns, void* slot, void* stream) {
    DeviceScope device_scope(h);
    if (!h || !X || !slot || !planes_chunks(columns) || width > columns) { set_error("planes_pack: bad argument"); return I2SDF_E_INVALID; }
    return planes_pack_launch(X, ld, width, M, (uint8_t*)slot, planes_chunks(columns), (cudaStream_t)stream);
}
int i2sdf_planes_unpack(i2sdf_handle* h, const void* slot, int columns, int64_t M, float* X, int ld, int width, void* stream) {
    DeviceScope device_scope(h);
    if (!h || !X || !slot || !planes_chunks(columns) || width > columns) { set_error("planes_unpack: bad argument"); return I2SDF_E_INVALID; }
    return planes_unpack_launch((const uint8_t*)slot, planes_chunks(columns), M, X, ld, width, (cudaStream_t)stream);
}
int i2sdf_planes_wgrad(i2sdf_handle* h, int nterms, const void* const* P, const void* const* X, int x_columns, int64_t M,
                       float* dW, int ld, int rows, int cols, float* colsum, void* stream) {
    DeviceScope device_scope(h);
    if (!h || !P || !X || !dW || nterms < 1 || nterms > 2 || !planes_chunks(x_columns) || rows > 256 || cols > x_columns) {
        set_error("planes_wgrad: bad argument"); return I2SDF_E_INVALID;
    }
    if (!h->use_tc) { set_error("planes_wgrad: tensor-core path disabled"); return I2SDF_E_INVALID; }
    WgArgs a{};
    a.ntiles = planes::ntiles(M); a.nterms = nterms; a.njobs = 1;
    a.jobs[0] = WgJob{dW, ld, rows, cols};
    for (int t = 0; t < nterms; ++t) a.terms[t] = WgTerm{(const uint8_t*)P[t], (const uint8_t*)X[t], 0, planes_chunks(x_columns), t == 0 ? colsum : nullptr, rows, 2, 2};
    return wgrad_planes_launch(h, a, (cudaStream_t)stream);
}

}  // extern "C"
