/*
 * i2sdf_b200 — C ABI of the B200-native I2-SDF volume-rendering core.
 *
 * The reference (jingsenzhu/i2-sdf) is pure Python/PyTorch and has no FFI of its own; the entry points
 * below are what a binding for its per-ray hot path would call.  Each one names the reference
 * function(s) it replaces (file:line relative to the reference tree).
 *
 * Conventions
 *   - plain C, `extern "C"`, no C++/torch types: raw DEVICE pointers (unless marked host), sizes, a stream.
 *   - every function returns 0 on success, a negative I2SDF_E_* code otherwise; i2sdf_last_error() gives text.
 *     No exception crosses this boundary.  Nothing here falls back to the CPU.
 *   - the caller owns every buffer, including the workspace (size from i2sdf_workspace_bytes()).
 *     The library's only state is the opaque handle: packed weights, tables, device properties.
 *   - calls are asynchronous on `stream` (a cudaStream_t passed as void*); no call synchronises the host.
 *   - all floating point tensors are fp32, row-major, densely packed; index tensors are int32 unless noted.
 */
#ifndef I2SDF_B200_H
#define I2SDF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define I2SDF_ABI_VERSION 1

enum {
    I2SDF_OK = 0,
    I2SDF_E_INVALID = -1,     /* bad argument / unsupported network shape */
    I2SDF_E_CUDA = -2,        /* a CUDA runtime call or launch failed */
    I2SDF_E_NOGPU = -3,       /* no sm_100 device */
    I2SDF_E_WORKSPACE = -4    /* workspace too small */
};

/* Network + sampler description: the `model:` node of a reference yaml
 * (config/synthetic.yml:30-74; read by I2SDFNetwork.__init__, model/network/__init__.py:20-47). */
typedef struct i2sdf_desc {
    int32_t abi_version;       /* I2SDF_ABI_VERSION */
    int32_t hidden;            /* width of every hidden layer; must be 256 */
    int32_t feature_size;      /* feature_vector_size; must be 256 */
    int32_t n_sdf_layers;      /* number of Linear layers of ImplicitNetwork (9 or 7) */
    int32_t sdf_skip_layer;    /* skip_in[0] (4 or 3), -1 for none */
    int32_t multires_x;        /* 6 -> 39-wide embedding of points */
    int32_t n_color_layers;    /* Linear layers of RenderingNetwork (5 or 4), mode 'nerf' */
    int32_t multires_d;        /* 4 -> 27-wide embedding of view dirs */
    int32_t n_light_layers;    /* 0, or 2 for the light-mask head [256,128,1] */
    int32_t light_hidden;      /* 128 */
    /* ErrorBoundSampler (ray_sampler.py:46-65) */
    int32_t n_samples;         /* 64 */
    int32_t n_samples_eval;    /* 128 */
    int32_t n_samples_extra;   /* 32 */
    int32_t beta_iters;        /* 10 */
    int32_t max_total_iters;   /* 5 */
    float near_;               /* 0 */
    float far_;                /* 2 * scene_bounding_sphere */
    float eps;                 /* 0.1 */
    float add_tiny;            /* 1e-6 */
    float beta_min;            /* LaplaceDensity.beta_min, 1e-4 (density.py:19) */
    float lemma2_coeff;        /* 1/(4*log(1+eps)) evaluated in fp32 by the caller (ray_sampler.py:76) */
    /* host tables computed by the caller with torch.linspace so that index arithmetic is bit-identical
     * to the reference: u_up[n_samples_eval], u_final[n_samples] (ray_sampler.py:188), t_init[n_samples_eval]
     * (ray_sampler.py:30), extra_idx[max_total_iters][n_samples_extra] = linspace(0, n-1, 32).long() for
     * n = 128,256,... (ray_sampler.py:225).  HOST pointers, copied at create time. */
    const float* u_up;
    const float* u_final;
    const float* t_init;
    const int32_t* extra_idx;
} i2sdf_desc;

typedef struct i2sdf_handle i2sdf_handle;

int i2sdf_abi_version(void);
const char* i2sdf_last_error(void);

/* Replaces: I2SDFNetwork.__init__ (model/network/__init__.py:20-47) for the device-side state. */
int i2sdf_create(const i2sdf_desc* desc, int device, i2sdf_handle** out);
int i2sdf_destroy(i2sdf_handle* h);

/* Bit 0: the sampler's SDF evaluations run on the tcgen05 tensor-core kernel; bit 1: the eval main pass
 * (SDF + grad_x + radiance) does.  0 = everything on the fp32 FMA-pipe kernels (environment I2SDF_SIMT=1 at create
 * time forces that, I2SDF_SIMT_MAIN=1 only the main pass; all of it is CUDA — there is no CPU path). */
int i2sdf_uses_tensor_cores(const i2sdf_handle* h);

/* Number of Linear layers the weight arrays below must hold: n_sdf + n_color + n_light, in that order. */
int i2sdf_num_layers(const i2sdf_handle* h);

/* Replaces: the weight-norm forward pre-hook + nn.Linear parameter reads (mlp.py:71-72,97,200-201,222).
 * W[i] : device pointer to the EFFECTIVE weight of layer i, row-major [out,in] (W = g*v/||v||, computed by
 * the caller so autograd sees it); b[i] : device pointer to bias [out].  `W`/`b` are HOST arrays of device
 * pointers.  Re-packs (transposes / pads / splits) into the handle's kernel layouts.  Call after every
 * optimizer step. */
int i2sdf_pack_weights(i2sdf_handle* h, const float* const* W, const float* const* b, void* stream);

/* Workspace size in bytes for a call over R rays (sampler + render) or M points (sdf_forward: pass R=M, 1 sample). */
size_t i2sdf_workspace_bytes(const i2sdf_handle* h, int64_t R, int training);

/* Replaces: utils.get_camera_params + lift (utils/rend_util.py:92-147) and the ray flattening /
 * normalisation at model/network/__init__.py:86-93.
 * uv [B,P,2], pose [B,4,4], intr [B,4,4]  ->  o [B*P,3], d [B*P,3] (unit), dnorm [B*P]. */
int i2sdf_rays(i2sdf_handle* h, const float* uv, const float* pose, const float* intr, int B, int P,
               float* o, float* d, float* dnorm, void* stream);

/* Replaces: ImplicitNetwork.forward / get_sdf_vals / get_outputs / gradient (mlp.py:84-151).
 * pts [M,3] -> out_sdf [M] (required), out_feat [M,256] or NULL, out_grad [M,3] (= d sdf/d x) or NULL.
 * save_act: NULL, or [n_sdf_layers-1, M, 256] to keep the pre-activations for i2sdf_sdf_backward. */
int i2sdf_sdf_forward(i2sdf_handle* h, const float* pts, int64_t M, float* out_sdf, float* out_feat,
                      float* out_grad, float* save_act, void* workspace, size_t workspace_bytes, void* stream);

/* Replaces: the SDF evaluation of a regular grid for mesh extraction / validation plots - `implicit_network(pnts)[:, 0]` over
 * `get_grid_uniform` / `get_grid` points (model/eval/recon.py:46-56, 87-90; utils/plots.py:188-225, 440-489).
 * The points are generated on the device from the three axis arrays in numpy.meshgrid(x, y, z) order - point m = (j, i, k) =
 * (m / (nx nz), (m / nz) % nx, m % nz) -> (gx[i], gy[j], gz[k]) - optionally mapped p' = A p + t (affine: 12 device floats, row-major
 * 3x3 then t: the PCA-aligned grid of eval/recon.py:80-84), and go through the sdf-only tensor-core chain: no point array is read, no
 * feature is computed or written (the reference evaluates and discards 256 of them per point).  out_sdf [nx*ny*nz]. */
int i2sdf_sdf_grid(i2sdf_handle* h, const float* gx, const float* gy, const float* gz, int nx, int ny, int nz, const float* affine,
                   float* out_sdf, void* stream);

/* Replaces: ErrorBoundSampler.get_z_vals rounds (ray_sampler.py:67-212): uniform init (+ stratified jitter),
 * up to max_total_iters rounds of {SDF of new samples, d*, beta line search, opacity-bound pdf, inverse CDF,
 * merge}.  No host synchronisation: the batch-global convergence test (ray_sampler.py:151) is evaluated on
 * the device and later rounds predicate themselves off.
 * beta_param : device pointer to density.beta (raw parameter; beta0 = |beta| + beta_min on device).
 * jitter [R,n_samples_eval] / u_final [R,n_samples] : training RNG tapes (ray_sampler.py:39,190) or NULL (eval).
 * State is left in the workspace for i2sdf_sampler_finalize. */
int i2sdf_sampler_rounds(i2sdf_handle* h, const float* o, const float* d, int64_t R, const float* beta_param,
                         const float* jitter, const float* u_final,
                         void* workspace, size_t workspace_bytes, void* stream);

/* The same rounds one stage at a time, for callers that shard the rays of ONE batch over several GPUs and want the
 * reference's batch-global convergence test (`beta.max() > beta0`, ray_sampler.py:151) instead of a per-shard one:
 *   stage 0: initial z's (jitter);  stage 1, round k: SDF on the round's samples + beta search, leaves max_rays(beta) in
 *   i2sdf_sampler_beta_max(...)[k], and (k + 1 < max_total_iters) up-samples every ray for round k + 1 in the same launch - the
 *   reference up-samples all rays or none (ray_sampler.py:151-176), so the draw is done speculatively and round k + 1 predicates
 *   itself off if the batch turns out converged;  stage 2, round k: the final N_samples draw, only if that word says round k was
 *   the last one (otherwise the launch returns at once).
 * Between stage 1 and stage 2 the caller MAX-all-reduces beta_max[k] across its ranks (4 bytes, stream-ordered; see
 * i2sdf_b200/core.py::sample(group=...)).  Positive floats: integer and float MAX agree.  i2sdf_sampler_rounds is the same
 * sequence with ONE final-draw launch behind the last round. */
int i2sdf_sampler_step(i2sdf_handle* h, const float* o, const float* d, int64_t R, const float* beta_param,
                       const float* jitter, const float* u_final, int stage, int k,
                       void* workspace, size_t workspace_bytes, void* stream);
float* i2sdf_sampler_beta_max(i2sdf_handle* h, int64_t R, void* workspace);

/* Replaces: ray_sampler.py:215-234 — extra samples, near/far, final sort, eikonal pick.
 * extra_idx : device int32[n_samples_extra] (training: randperm(n)[:32], ray_sampler.py:223) or NULL (eval table).
 * eik_idx   : device int32[R] (ray_sampler.py:233) or NULL.
 * out_z [R, n_samples+2+n_samples_extra] sorted; out_z_eik [R] or NULL.
 * out_info (device int32[2] or NULL): [0] = rounds executed, [1] = z's per ray when sampling stopped (128*rounds). */
int i2sdf_sampler_finalize(i2sdf_handle* h, int64_t R, const float* beta_param, const int32_t* extra_idx,
                           const int32_t* eik_idx, float* out_z, float* out_z_eik, int32_t* out_info,
                           void* workspace, size_t workspace_bytes, void* stream);

/* Same without the host round trip a training caller would otherwise need to learn n before drawing randperm(n)[:32]:
 * extra_table = device int32[max_total_iters][n_samples_extra], row k = the index set to use if sampling stopped after
 * k+1 rounds (n = 128 (k+1)); the kernel picks the row on the device.  The caller draws every candidate from the SAME
 * host generator state and re-draws the one that applied once out_info has been copied back (i2sdf_b200/core.py). */
int i2sdf_sampler_finalize_candidates(i2sdf_handle* h, int64_t R, const float* beta_param, const int32_t* extra_table,
                                      const int32_t* eik_idx, float* out_z, float* out_z_eik, int32_t* out_info,
                                      void* workspace, size_t workspace_bytes, void* stream);

/* Device int32[2] as above WITHOUT finalising: lets a training caller read n (one 8-byte D2H, the only host
 * sync of the path; the reference syncs once per round at ray_sampler.py:151) to draw randperm(n)[:32]. */
int i2sdf_sampler_info(i2sdf_handle* h, int64_t R, const float* beta_param, int32_t* out_info,
                       void* workspace, size_t workspace_bytes, void* stream);

/* Debug/parity entry: one sampler round on caller-supplied sorted z [R,n] and sdf [R,n] (identical inputs to
 * the oracle's round): outputs beta [R], cdf [R,n], inds [R,ns] (searchsorted right=True, ray_sampler.py:193),
 * samples [R,ns], and if upsample: z_merged [R,n+ns], src [R,n+ns] (merge source index, ray_sampler.py:212).
 * beta_in [R]: upper bound entering the round.  force_upsample: 1 = opacity-bound pdf (:153-171), 0 = final (:173-183). */
int i2sdf_sampler_round_debug(i2sdf_handle* h, const float* z, const float* sdf, int64_t R, int n,
                              const float* beta_param, const float* beta_in, int force_upsample,
                              const float* u_tape, float* out_beta, float* out_cdf, int32_t* out_inds,
                              float* out_samples, float* out_z_merged, int32_t* out_src, void* stream);

/* Replaces: the main pass of I2SDFNetwork.forward (model/network/__init__.py:99-125,162-170,204-219) +
 * volume_rendering (:223-240): points o+z*d -> SDF(+grad_x) -> radiance MLP -> Laplace density -> alpha
 * compositing (+ light-mask head).
 * z [R,N+1] (last column = z_max).  Per-ray outputs (any may be NULL): rgb [R,3], depth [R], weight_sum [R],
 * normal [R,3] (normalize(sum w * normalize(grad))), light [R].
 * Per-sample outputs kept for backward / parity (any may be NULL): s_sdf [R*N], s_grad [R*N,3], s_rgb [R*N,3],
 * s_w [R*N], s_light [R*N].   save: NULL or an i2sdf_saved-sized buffer (see i2sdf_saved_bytes). */
int i2sdf_render_forward(i2sdf_handle* h, const float* o, const float* d, const float* dnorm, const float* z,
                         int64_t R, int N, const float* beta_param,
                         float* rgb, float* depth, float* weight_sum, float* normal, float* light,
                         float* s_sdf, float* s_grad, float* s_rgb, float* s_w, float* s_light,
                         void* save, size_t save_bytes, void* workspace, size_t workspace_bytes, void* stream);

size_t i2sdf_saved_bytes(const i2sdf_handle* h, int64_t R, int N);

/* ---- training: split forward + backward ------------------------------------------------------------------
 * The reference obtains these through torch.autograd (loss.backward(), model/trainer/recon.py:254-287), including
 * the double-backward through autograd.grad(create_graph=True) at mlp.py:107-143.  W / b arguments are HOST arrays
 * of DEVICE pointers to the effective weights [out,in] / biases of the addressed stack, in layer order; dW / db are
 * accumulated INTO (caller zero-initialises).  All need workspace >= i2sdf_backward_workspace_bytes(h, M). */
size_t i2sdf_backward_workspace_bytes(const i2sdf_handle* h, int64_t M);

/* Main-pass MLP chain only (no compositing): per-sample s_sdf [R*N], s_grad [R*N,3] (or NULL), s_rgb [R*N,3],
 * s_light [R*N] (or NULL), s_feat [R*N,256] (or NULL), save_act [n_sdf_layers-1, R*N, 256] (or NULL).
 * Replaces model/network/__init__.py:103-116,162-168. */
int i2sdf_points_forward(i2sdf_handle* h, const float* o, const float* d, const float* z, int64_t R, int N,
                         float* s_sdf, float* s_grad, float* s_rgb, float* s_light, float* s_feat, float* save_act,
                         void* workspace, size_t workspace_bytes, void* stream);

/* Same with n_extra explicit points [n_extra,3] appended after the R*N ray samples (the eikonal / smoothness points of
 * model/network/__init__.py:175-193 ride the main-pass launch): every per-sample output has R*N + n_extra rows, the
 * saved state is sized by i2sdf_saved_bytes_points(h, R*N + n_extra).  Tensor-core main pass only.
 * Networks with a light-mask head: s_light keeps R*N rows (the head is evaluated on the ray samples only, as its own pass
 * over s_feat, which is then REQUIRED together with workspace >= i2sdf_workspace_bytes(h, R, 0)). */
int i2sdf_points_forward_ex(i2sdf_handle* h, const float* o, const float* d, const float* z, int64_t R, int N,
                            const float* extra_pts, int64_t n_extra, float* s_sdf, float* s_grad, float* s_rgb,
                            float* s_light, float* s_feat, float* save_act, void* workspace, size_t workspace_bytes,
                            void* stream);
size_t i2sdf_saved_bytes_points(const i2sdf_handle* h, int64_t M);

/* Compositing only (model/network/__init__.py:118-125,169,204-219,223-240) on per-sample inputs. */
int i2sdf_composite_forward(i2sdf_handle* h, const float* z, const float* dnorm, const float* s_sdf, const float* s_rgb,
                            const float* s_grad, const float* s_light, const float* beta_param, int64_t R, int N,
                            float* rgb, float* depth, float* weight_sum, float* normal, float* light, float* s_w, void* stream);

/* Backward of i2sdf_composite_forward.  Upstream g_* may be NULL (= 0).  Outputs: o_sdf [R*N], o_rgb [R*N,3],
 * o_grad [R*N,3] (normal path; weights are detached there as in the reference, :207), o_light [R*N] (weights
 * detached, :169), o_beta [1] (accumulated; d/d density.beta). */
int i2sdf_composite_backward(i2sdf_handle* h, const float* z, const float* dnorm, const float* s_sdf, const float* s_rgb,
                             const float* s_grad, const float* s_light, const float* beta_param, int64_t R, int N,
                             const float* g_rgb, const float* g_depth, const float* g_wsum, const float* g_normal,
                             const float* g_light, float* o_sdf, float* o_rgb, float* o_grad, float* o_light, float* o_beta,
                             void* stream);

/* Radiance stack backward (mlp.py:208-229).  dirs [M/ns, 3]; feat [M,256]; s_rgb [M,3] (forward output);
 * g_rgb [M,3].  g_x [M,288] receives the adjoint of the stack's input [PE(dir) | feat | pad]: the feature adjoint
 * is g_x + (3 + 6*multires_d) with leading dimension 288. */
int i2sdf_color_backward(i2sdf_handle* h, const float* const* W, const float* const* b, const float* dirs, int ns,
                         const float* feat, const float* s_rgb, const float* g_rgb, int64_t M, float* const* dW,
                         float* const* db, float* g_x, void* workspace, size_t workspace_bytes, void* stream);

/* Light-mask head backward (model/network/__init__.py:162-168; features detached -> head parameters only).
 * feat [M,256] fp32 features; hidden: NULL, or (tensor-core path) the head's hidden pre-activations [M,light_hidden] that
 * i2sdf_points_forward_ex left behind the plane slots of its saved state (at byte offset
 * i2sdf_saved_bytes_points(h, M_total) - M_total * light_hidden * 4), which saves recomputing them. */
int i2sdf_light_backward(i2sdf_handle* h, const float* const* W, const float* const* b, const float* feat,
                         const float* hidden, const float* s_light, const float* g_light, int64_t M, float* const* dW,
                         float* const* db, void* workspace, size_t workspace_bytes, void* stream);

/* SDF stack backward incl. second order.  Points: pts [M,3], or (pts NULL) rays o,d [M/ns,3] with z [.., zstride].
 * act: pre-activations saved by the forward.  g_sdf [M] / g_feat [M, ld g_feat_ld] / g_grad [M,3] upstream (NULL = 0);
 * g_grad is the upstream of grad_x sdf (eikonal / normal terms). */
int i2sdf_sdf_backward(i2sdf_handle* h, const float* const* W, const float* pts, const float* o, const float* d,
                       const float* z, int zstride, int ns, int64_t M, const float* act, const float* g_sdf,
                       const float* g_feat, int g_feat_ld, const float* g_grad, float* const* dW, float* const* db,
                       void* workspace, size_t workspace_bytes, void* stream);

/* Measurement hook (bench.py): when enabled, every kernel launch is bracketed by CUDA events recorded on the
 * launching stream and counted, per class: 0 = sampler SDF evaluations (the dominant kernel), 1 = main-pass MLP
 * chain, 2 = per-ray sampler kernels, 3 = rays / compositing / weight packing.  enable=1 also resets the counters.
 * i2sdf_profile_read synchronises the device and returns summed event durations (ms) and launch counts. */
int i2sdf_profile_enable(i2sdf_handle* h, int enable);
int i2sdf_profile_read(i2sdf_handle* h, float ms[4], int64_t launches[4]);
/* same with n <= 8 classes: 4 = backward chain kernel (fused training path), 5 = weight-gradient kernel,
 * 6 = light-mask head (forward pass over the features + its backward). */
int i2sdf_profile_read_n(i2sdf_handle* h, int n, float* ms, int64_t* launches);
/* Development probe (no reference counterpart): with I2SDF_DEBUG_TIMELINE set in the environment the backward chain kernel stamps clock64
 * values of one CTA's second tile into a device buffer (csrc/mlp_tc_bwd.cu: tc_bwd8_kernel<true>); this copies up to n int64 of it to the
 * host (synchronous) and returns how many were copied, 0 if the probe never ran.  tools/timeline.py bwd prints the table. */
int64_t i2sdf_debug_bwd_timeline(int64_t* out, int64_t n);

/* What a forward in training mode saves for the backward.  format 1 = plane slots (tensor-core chain kernels; consumed
 * by i2sdf_fused_backward), 0 = fp32 pre-activations [L-1][M][256] (fp32 kernels; consumed by i2sdf_sdf_backward /
 * i2sdf_color_backward).  kind 0: main pass (i2sdf_points_forward / i2sdf_render_forward, size i2sdf_saved_bytes),
 * kind 1: stand-alone SDF points (i2sdf_sdf_forward with out_grad, size i2sdf_sdf_saved_bytes). */
int i2sdf_saved_format(const i2sdf_handle* h, int kind);
size_t i2sdf_sdf_saved_bytes(const i2sdf_handle* h, int64_t M);

/* Fused backward of the per-point networks on plane slots: the tangent pass of the SDF stack along g_grad, the
 * reverse passes of the radiance and SDF stacks, all weight and bias gradients — what loss.backward() runs through
 * ImplicitNetwork (mlp.py:84-143, incl. the second-order graph) and RenderingNetwork (mlp.py:208-229) in the reference.
 * Points as in i2sdf_sdf_backward; with BOTH pts and rays given the first m_rays points are ray samples and the rest
 * explicit (the layout i2sdf_points_forward_ex produced).  saved: the buffer the forward filled (format 1).  g_sdf [M], g_grad [M,3],
 * g_rgb [M,3] upstream (NULL = 0; g_rgb NULL: SDF stack only, e.g. eikonal points); s_rgb [M,3] forward rgb.
 * dW_sdf / db_sdf / dW_col / db_col: per-layer gradient buffers in API order, ACCUMULATED into.
 * workspace >= i2sdf_backward_workspace_bytes(h, M). */
int i2sdf_fused_backward(i2sdf_handle* h, const float* pts, const float* o, const float* d, const float* z, int zstride,
                         int ns, int64_t M, int64_t m_rays, void* saved, const float* s_rgb, const float* g_sdf, const float* g_grad,
                         const float* g_rgb, float* const* dW_sdf, float* const* db_sdf, float* const* dW_col,
                         float* const* db_col, void* workspace, size_t workspace_bytes, void* stream);

/* The same when the caller appended explicit points behind its ray samples (i2sdf_points_forward_ex: the eikonal / smoothness points of
 * model/network/__init__.py:175-193 ride the main-pass launches): the upstream arrays g_sdf / g_grad / g_rgb cover only the first m_up
 * points (the ray samples), points m >= m_up have no sdf / rgb upstream and take the upstream of grad_x sdf from g_grad_tail [M - m_up, 3]
 * (NULL = 0).  Saves the caller three concatenations with zero blocks per training step. */
int i2sdf_fused_backward_ex(i2sdf_handle* h, const float* pts, const float* o, const float* d, const float* z, int zstride,
                            int ns, int64_t M, int64_t m_rays, void* saved, const float* s_rgb, const float* g_sdf, const float* g_grad,
                            const float* g_rgb, int64_t m_up, const float* g_grad_tail, float* const* dW_sdf, float* const* db_sdf,
                            float* const* dW_col, float* const* db_col, void* workspace, size_t workspace_bytes, void* stream);

/* Plane slots: the HBM format of the fused training path (bf16 hi + lo planes per 128-point tile in the tensor
 * cores' SMEM layout, i2sdf_b200/csrc/planes.cuh).  pack / unpack convert from / to plain fp32 [M][ld] arrays
 * (columns = 256 or 48); planes_wgrad is the weight-gradient kernel on its own:
 *   dW[rows][ld] += sum_t P_t^T X_t   (P_t: 256-column slots, X_t: x_columns-column slots, t < nterms <= 2),
 *   colsum[j] += sum_m P_0[m][j] when colsum != NULL.
 * These three are what the parity tests drive; the training path calls the same kernels internally. */
size_t i2sdf_planes_slot_bytes(int64_t M, int columns);
int i2sdf_planes_pack(i2sdf_handle* h, const float* X, int ld, int width, int64_t M, int columns, void* slot, void* stream);
int i2sdf_planes_unpack(i2sdf_handle* h, const void* slot, int columns, int64_t M, float* X, int ld, int width, void* stream);
int i2sdf_planes_wgrad(i2sdf_handle* h, int nterms, const void* const* P, const void* const* X, int x_columns, int64_t M,
                       float* dW, int ld, int rows, int cols, float* colsum, void* stream);

/* ---- loss ---------------------------------------------------------------------------------
 * Replaces: I2SDFLoss.forward (model/network/__init__.py:338-406) AND the backward autograd runs through it: one launch
 * computes every term, the weighted total and d loss / d input for every model output that carries gradient.
 * All pointers are device memory; a NULL prediction pointer switches its term off (the caller applies the reference's
 * "key present and weight > 0" rules, e.g. diff_norm is passed only once current_step > smooth_iter).  Masks are bytes
 * (torch.bool).  g_* outputs may be NULL.  No handle: the loss has no network state.
 * denom (optional): rays sharded over several GPUs (SURVEY.md §8(e) caveat 2).  The reference's means run over the WHOLE
 * batch; a shard's own masked means differ from them as soon as the shards' mask counts differ.  With denom != NULL the
 * kernel divides this shard's sums by the five given divisors instead of its local counts; the caller passes
 * (global count) / (number of shards), so that the plain average over the shards of the returned loss - and of the
 * parameter gradients, which the gradient all-reduce averages anyway - is the single-GPU loss / gradient of the whole batch. */
#define I2SDF_LOSS_DENOM_RAYS 0    /* divisor of the per-ray means: rgb (x3), smooth, mask BCE, light BCE */
#define I2SDF_LOSS_DENOM_EIK 1     /* eikonal rows */
#define I2SDF_LOSS_DENOM_BUBBLE 2  /* bubble points */
#define I2SDF_LOSS_DENOM_DEPTH 3   /* rays with depth_mask */
#define I2SDF_LOSS_DENOM_NORMAL 4  /* rays with normal_mask */
typedef struct i2sdf_loss_args {
    int64_t R;                   /* rays */
    int64_t n_eik;               /* rows of grad_theta (2R) */
    int64_t n_bubble;            /* rows of surface_sdf */
    const float* rgb;            /* [R,3] rgb_values */
    const float* rgb_gt;         /* [R,3] */
    const float* grad_theta;     /* [n_eik,3] or NULL */
    const float* diff_norm;      /* [R] or NULL */
    const float* weight_sum;     /* [R] or NULL  (mask BCE) */
    const float* mask_gt;        /* [R] */
    const float* depth;          /* [R] or NULL */
    const float* depth_gt;       /* [R] */
    const uint8_t* depth_mask;   /* [R] */
    const float* normal;         /* [R,3] normal_values or NULL */
    const float* normal_gt;      /* [R,3] */
    const uint8_t* normal_mask;  /* [R] */
    const float* surface_sdf;    /* [n_bubble] or NULL */
    const float* light;          /* [R] light_mask or NULL */
    const float* light_gt;       /* [R] */
    float w_eik, w_smooth, w_mask, w_depth, w_normal, w_angular, w_bubble, w_light;
    float* terms;                /* [10]: loss, rgb, eikonal, smooth, mask, depth, normal, angular, bubble, light_mask */
    float* g_rgb;                /* [R,3] */
    float* g_grad_theta;         /* [n_eik,3] */
    float* g_diff_norm;          /* [R] */
    float* g_weight_sum;         /* [R] */
    float* g_depth;              /* [R] */
    float* g_normal;             /* [R,3] */
    float* g_surface_sdf;        /* [n_bubble] */
    float* g_light;              /* [R] */
    const float* denom;          /* [5] device, I2SDF_LOSS_DENOM_* order, or NULL = this call's own counts */
} i2sdf_loss_args;
int i2sdf_loss_forward(const i2sdf_loss_args* args, void* stream);

/* ---- weight norm --------------------------------------------------------------------------
 * Replaces: the nn.utils.weight_norm pre-forward hook and its autograd backward for EVERY layer of the model
 * (mlp.py:71-72, 200-201; torch._weight_norm with dim = 0), one launch per direction.
 * forward (backward = 0): W = g * v / ||v||_row, norm[r] = ||v[r,:]|| (kept for the backward).
 * backward (backward = 1): given dW: dg, dv (written, not accumulated); a job with dW == NULL is skipped. */
#define I2SDF_WNORM_MAX_JOBS 28
typedef struct i2sdf_wnorm_job {
    const float* g;       /* [rows]      weight_g */
    const float* v;       /* [rows,cols] weight_v */
    float* W;             /* [rows,cols] out (forward) */
    float* norm;          /* [rows]      out (forward) / in (backward) */
    const float* dW;      /* [rows,cols] in (backward) or NULL */
    float* dg;            /* [rows]      out (backward) */
    float* dv;            /* [rows,cols] out (backward) */
    int32_t rows, cols;
} i2sdf_wnorm_job;
typedef struct i2sdf_wnorm_batch {
    int32_t n;
    int32_t pad_;
    i2sdf_wnorm_job jobs[I2SDF_WNORM_MAX_JOBS];
} i2sdf_wnorm_batch;
int i2sdf_weight_norm(const i2sdf_wnorm_batch* batch, int backward, void* stream);

/* ---- optimizer --------------------------------------------------------------------------
 * Replaces: torch.optim.Adam(...).step() of the reconstruction trainer (model/trainer/recon.py:201-207) for every
 * parameter tensor in one launch.  The host passes the step's scalars: step_size = lr / (1 - beta1^t),
 * bias_correction2_sqrt = sqrt(1 - beta2^t).  No weight decay / amsgrad (the reference uses neither). */
#define I2SDF_ADAM_MAX_JOBS 64
typedef struct i2sdf_adam_job {
    float* param;           /* [numel] updated in place */
    const float* grad;      /* [numel] */
    float* exp_avg;         /* [numel] */
    float* exp_avg_sq;      /* [numel] */
    int64_t numel;
} i2sdf_adam_job;
typedef struct i2sdf_adam_batch {
    int32_t n;
    float beta1, beta2, eps, step_size, bias_correction2_sqrt;
    float one_minus_beta1, one_minus_beta2;   /* 1 - beta formed in double by the host (1 - 0.999f != (float)0.001) */
    i2sdf_adam_job jobs[I2SDF_ADAM_MAX_JOBS];
} i2sdf_adam_batch;
int i2sdf_adam_step(const i2sdf_adam_batch* batch, void* stream);
/* The same with the two per-step scalars {step_size, bias_correction2_sqrt} read from device memory at run time (scalars_dev [2];
 * NULL: the values in the batch): a launch captured into a CUDA graph stays valid from step to step. */
int i2sdf_adam_step_dev(const i2sdf_adam_batch* batch, const float* scalars_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* I2SDF_B200_H */
