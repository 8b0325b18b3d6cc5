/*
 * panonerf_b200 — C ABI of the B200-native mip-NeRF volumetric-rendering hot path of Pano-NeRF.
 *
 * This is the drop-in boundary.  The reference has no FFI: its "operator API" for this path is
 * the Python function surface of models/mip.py, models/{mip_nerf,pano_mip_nerf}.py and
 * utils/surface_rendering.py.  Each entry point below replaces the arithmetic of the reference function cited
 * beside it; the Python host (panonerf_b200/models/*.py) keeps the reference's names and signatures and calls
 * these through ctypes (INTEGRATION.md shows the binding).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer (sm_100a, current device) unless the comment says "host";
 *   - tensors are contiguous row-major; `ld*` arguments are row strides in ELEMENTS;
 *   - fp32 unless a `dtype` argument says otherwise (PNB_F32 / PNB_BF16);
 *   - nothing here allocates device memory: outputs and workspaces are caller-owned;
 *   - `stream` is a cudaStream_t passed as void*; kernels are launched asynchronously on it;
 *   - return value: 0 on success, otherwise a cudaError_t value (or PNB_ERR_ARG) — pnb_last_error() gives text.
 *   - there is NO CPU fallback anywhere behind this ABI.
 */
#ifndef PANONERF_B200_H_
#define PANONERF_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PNB_ABI_VERSION 1
#define PNB_F32 0
#define PNB_BF16 1
#define PNB_ERR_ARG 10001

/* epilogue flags of the GEMM entry points */
#define PNB_EPI_BIAS 1       /* C += bias[n]                                                       */
#define PNB_EPI_RELU 2       /* C = max(C, 0)                                                      */
#define PNB_EPI_MASK 4       /* C = mask_src[m,n] > 0 ? C : 0   (ReLU backward / Jacobian chain)   */
#define PNB_EPI_ACCUM 8      /* C += previous C (fp32 outputs only)                                */

int pnb_abi_version(void);
const char* pnb_last_error(void);
/* number of kernels this library has launched since load (bench.py's `gpu_launches`). */
long long pnb_launch_count(void);

/* ---- K1  equirectangular ray generation: datasets/pano_datasets.py:152-216 -------------------------------
 * rows [row0,row0+nrows) of an H x W panorama; c2w = 12 floats (3x4 row-major, HOST pointer).
 * Outputs are [nrows*W, 3] / [nrows*W, 1] blocks of the 8 `Rays` fields (datasets/base_datasets.py:13-16). */
int pnb_raygen_equirect(int H, int W, int row0, int nrows, const float* c2w_host, float near_v, float far_v,
                        float* origins, float* directions, float* viewdirs, float* radii, float* lossmult,
                        float* near_o, float* far_o, float* noise_var, void* stream);

/* ---- K2  stratified sampling + conical-frustum Gaussians: models/mip.py:113-151, 154-194, 67-89, 36-58, 8-22
 * ray r uses origins[r / o_div] and directions|radii|near|far[(d_mod ? r % d_mod : r)]
 * (main rays: o_div=1,d_mod=0; env rays of sample_each_points: o_div=D,d_mod=D).
 * s_lin = torch.linspace(0,1,N+1); t_rand (nullable) has row stride rand_ld (0 = one row shared by all rays). */
int pnb_sample_cast(int R, int N, const float* origins, int o_div, const float* directions, const float* radii,
                    const float* near_v, const float* far_v, int d_mod, const float* s_lin, const float* t_rand,
                    int rand_ld, int disparity, float* t_out, float* means, float* covs, void* stream);
/* sample_each_points_hemisp, models/mip.py:197-237: as above with o_div = d_mod = D, but every ray brings its own
 * direction, directions[R,3] (a hemisphere of D directions rotated onto each surface normal). */
int pnb_sample_cast_hemisp(int R, int N, const float* origins, int o_div, const float* directions, const float* radii,
                           const float* near_v, const float* far_v, int d_mod, const float* s_lin,
                           const float* t_rand, int rand_ld, float* t_out, float* means, float* covs, void* stream);
/* cast_rays on given fence-posts t[R,N+1] (after resampling): models/mip.py:67-89 */
int pnb_cast_rays(int R, int N, const float* t, const float* origins, int o_div, const float* directions,
                  const float* radii, int d_mod, float* means, float* covs, void* stream);

/* ---- integrated positional encoding: models/mip.py:394-428 + 355-361 ; pos_enc: models/mip.py:431-441 -----
 * enc row = [sin(3L) | cos(3L)], L = max_deg-min_deg, written at enc + m*ld with element type `dtype`. */
int pnb_ipe_fwd(int M, const float* means, const float* covs, int min_deg, int max_deg, void* enc, int ld,
                int dtype, void* stream);
/* d_means[m,c] = sum_f d_enc[m,f] * d enc_f / d mean_c   (vector-Jacobian product, used for normals and for
 * the env-branch gradient into the surface point) */
int pnb_ipe_vjp(int M, const float* means, const float* covs, int min_deg, int max_deg, const void* d_enc,
                int ld, int dtype, float* d_means, void* stream);
/* out[m,f] = sum_c (d enc_f / d mean_c) * v[m,c]   (Jacobian-vector product: the adjoint of pnb_ipe_vjp); bf16 rows
 * (the tensor-core path's operand) take sin / cos / exp2 from the SFU, fp32 rows keep ~1 ulp cosines on the exact phase */
int pnb_ipe_jvp(int M, const float* means, const float* covs, int min_deg, int max_deg, const float* v,
                void* out, int ld, int dtype, void* stream);
int pnb_pos_enc(int R, const float* x, int deg, float* out, void* stream);

/* ---- activations of compute_graph: models/pano_mip_nerf.py:264-278, models/mip_nerf.py:238-241 ------------
 * raw_den is [M,C] (C=1 mipnerf, C=5 panonerf: sigma | albedo(3) | roughness).  sp1 (nullable) receives
 * softplus'(raw_den[:,0]+bias) which the normals need. */
int pnb_act_fwd(int M, int C, const float* raw_rgb, const float* raw_den, float density_bias, float rgb_padding,
                float* rgb, float* density, float* albedo, void* stream);
int pnb_act_bwd(int M, int C, const float* raw_rgb, const float* raw_den, float density_bias, float rgb_padding,
                const float* d_rgb, const float* d_density, const float* d_albedo, float* d_raw_rgb,
                float* d_raw_den, void* stream);
/* normals_raw = -softplus'(raw0+bias) * v ; and its backward (models/pano_mip_nerf.py:299-302) */
int pnb_density_grad_fwd(int M, int C, const float* raw_den, float density_bias, const float* v, float* n_raw,
                         void* stream);
int pnb_density_grad_bwd(int M, int C, const float* raw_den, float density_bias, const float* v,
                         const float* d_n_raw, float* d_raw0, float* d_v, void* stream);

/* ---- K6  alpha compositing: models/mip.py:444-483 ---------------------------------------------------------
 * density is [R,N] (channel 0 already selected); dirs[(d_mod ? r % d_mod : r)].
 * white_bkgd is a flag word: bit 0 = white background (mip.py:477-478), bit 1 (PNB_COMPOSITE_ATTENUATE) = every
 * sample's colour is attenuated by 1 / (1 + t_mid^2): `volumetric_lighting_composing`, models/mip.py:486-527. */
#define PNB_COMPOSITE_WHITE_BKGD 1
#define PNB_COMPOSITE_ATTENUATE 2
int pnb_composite_fwd(int R, int N, const float* rgb, const float* density, const float* t, const float* dirs,
                      int d_mod, int white_bkgd, float* comp_rgb, float* distance, float* acc, float* weights,
                      void* stream);
int pnb_composite_bwd(int R, int N, const float* rgb, const float* density, const float* t, const float* dirs,
                      int d_mod, int white_bkgd, const float* g_comp, const float* g_dist, const float* g_acc,
                      const float* g_weights, float* d_rgb, float* d_density, void* stream);

/* K6 with compute_graph's activations fused in (models/pano_mip_nerf.py:264-278 + models/mip.py:444-483 in one pass):
 * raw_rgb [R*N,3] and raw_den [R*N,C] are the MLP's head outputs; rgb = softplus(raw_rgb) (1 + 2 pad) - pad and
 * density = softplus(raw_den[:,0] + density_bias) live in registers only; albedo (nullable, [R*N,3], needs C >= 4)
 * receives sigmoid(raw_den[:,1:4]) 0.77 + 0.03.  Results are bit-identical to pnb_act_fwd + pnb_composite_fwd
 * (N <= 256) and pnb_composite_bwd + pnb_act_bwd.  d_raw_den is [R*N,C]: channel 0 = density, 1..3 = g_albedo through
 * the sigmoid (zero when g_albedo is null), the rest zero.  d_t (nullable, [R,N+1]) receives the gradient w.r.t. the
 * fence-posts (delta_i and t_mid_i of mip.py:456-462, the clamp of :476): `stop_resample_grad = False`. */
int pnb_act_composite_fwd(int R, int N, int C, const float* raw_rgb, const float* raw_den, float density_bias,
                          float rgb_padding, const float* t, const float* dirs, int d_mod, int white_bkgd,
                          float* comp_rgb, float* distance, float* acc, float* weights, float* albedo, void* stream);
int pnb_act_composite_bwd(int R, int N, int C, const float* raw_rgb, const float* raw_den, float density_bias,
                          float rgb_padding, const float* t, const float* dirs, int d_mod, int white_bkgd,
                          const float* g_comp, const float* g_dist, const float* g_acc, const float* g_weights,
                          const float* g_albedo, float* d_raw_rgb, float* d_raw_den, float* d_t, void* stream);

/* ---- `stop_resample_grad = False` (models/mip.py:336-350: the fine level's loss reaches the coarse weights) ----------
 * pnb_ipe_cov_hess: d_covs = (d enc / d cov)^T d_enc (d_enc nullable); with h_enc (= d raw_sigma / d enc, the
 *   Jacobian sweep's output, fp32) and d_v (= dL/d v, v = J_ipe^T h_enc) also the explicit second-order terms of the
 *   density-gradient normals, d_v * d v / d mean into d_means and d_v * d v / d cov into d_covs (the ReLU network is
 *   piece-wise linear in enc: what autograd through vmap(jacrev), pano_mip_nerf.py:295-302, adds for them).
 *   accumulate != 0: += into d_means / d_covs (either may be null).
 * pnb_cast_rays_bwd: d_t[R,N+1] (+)= (d means / d t)^T g_means + (d covs / d t)^T g_covs  (mip.py:67-89, cone).
 * pnb_resample_bwd: d_weights[R,N] = (d new_t / d weights)^T g_new_t through the lerp, min(1, cumsum), the
 *   normalisation (incl. the 1e-5 padding branch) and the blur-pool maxima, with torch's tie rules; same t / weights /
 *   padding / blur_pool / u arguments as the forward call. */
int pnb_ipe_cov_hess(int M, const float* means, const float* covs, int min_deg, int max_deg, const void* d_enc, int ld,
                     int dtype, const float* h_enc, int ldh, const float* d_v, float* d_means, float* d_covs,
                     int accumulate, void* stream);
int pnb_cast_rays_bwd(int R, int N, const float* t, const float* directions, const float* radii, const float* g_means,
                      const float* g_covs, float* d_t, int accumulate, void* stream);
int pnb_resample_bwd(int R, int N, const float* t, const float* weights, float padding, int blur_pool, const float* u,
                     int u_ld, const float* g_new_t, float* d_weights, void* stream);

/* ---- K7  hierarchical resampling: models/mip.py:304-352 (blur-pool) + 240-301 (PDF/CDF/searchsorted/lerp) --
 * u: [N+1] when u_ld==0 (deterministic linspace(0,1-eps,N+1)) or [R,N+1] (u_ld=N+1, randomized).
 * inds (nullable) receives torch.searchsorted(cdf,u,right=True) as int64 — bit-exact contract.
 * blur_pool=0 skips the blur-pool/padding step (plain sorted_piecewise_constant_pdf on the given weights). */
int pnb_resample(int R, int N, const float* t, const float* weights, float padding, int blur_pool, const float* u,
                 int u_ld, float* new_t, long long* inds, void* stream);
/* Same, followed by cast_rays on the new fence-posts in the same kernel (models/mip.py:351 -> :67-89): the whole of
 * resample_along_rays in one launch.  means/covs [R,N,3] (both nullable together: then identical to pnb_resample);
 * origins/directions [R,3], radii [R]. */
int pnb_resample_cast(int R, int N, const float* t, const float* weights, float padding, int blur_pool,
                      const float* u, int u_ld, float* new_t, long long* inds, const float* origins,
                      const float* directions, const float* radii, float* means, float* covs, void* stream);

/* ---- normals / orientation loss / albedo compositing: models/pano_mip_nerf.py:296-317 ---------------------- */
int pnb_normals_fwd(int R, int N, const float* n_raw, const float* weights, const float* dirs,
                    const float* albedos, float* normal, float* ort, float* albedo, void* stream);
int pnb_normals_bwd(int R, int N, const float* n_raw, const float* weights, const float* dirs,
                    const float* albedos, const float* g_normal, const float* g_ort, const float* g_albedo,
                    float* d_n_raw, float* d_weights, float* d_albedos, void* stream);

/* ---- surface point + Lambertian shading: models/pano_mip_nerf.py:321-324, utils/surface_rendering.py:104-165 */
int pnb_surface_point_fwd(int R, const float* origins, const float* dirs, const float* dist, float* pts,
                          void* stream);
/* d_dist[r] = dirs[r] . sum_k d_means[r,k,:]   (k runs over the D*Ne env samples of ray r) */
int pnb_surface_point_bwd(int R, int K, const float* dirs, const float* d_means, float* d_dist, void* stream);
int pnb_shade_fwd(int R, int D, const float* env_rgb, const float* albedo, const float* normal,
                  const float* light_dirs, const float* solid_angle, float* surface_rgb, float* shading,
                  void* stream);
int pnb_shade_bwd(int R, int D, const float* env_rgb, const float* albedo, const float* normal,
                  const float* light_dirs, const float* solid_angle, const float* g_rgb, const float* g_shading,
                  float* d_env, float* d_albedo, float* d_normal, void* stream);

/* ---- tone mapping + losses: utils/surface_rendering.py:319-344, systems/panonerf_system.py:39-67 ---------- */
int pnb_hdr_to_ldr(long long n, const float* x, int quantize_u8, float* out, void* stream);
/* partial[r] = mask[r] * sum_c (ldr(pred[r,c]) - gt[r,c])^2 ; d_pred = scale * mask * 2 (ldr-gt) * ldr'  */
int pnb_tonemap_se_fwd(int R, const float* pred, const float* gt_ldr, const float* mask, float* partial,
                       void* stream);
int pnb_tonemap_se_bwd(int R, const float* pred, const float* gt_ldr, const float* mask, const float* g_scale,
                       float* d_pred, void* stream);
/* partial[r] = sum_c (normalize(gt)[r,c] - normalize(alb)[r,c])^2 */
int pnb_chroma_fwd(int R, const float* gt_ldr, const float* albedo, float* partial, void* stream);
int pnb_chroma_bwd(int R, const float* gt_ldr, const float* albedo, const float* g_scale, float* d_albedo,
                   void* stream);
/* deterministic sum of n floats -> out[0] (two-pass, workspace ws >= 1024 floats) */
int pnb_sum(long long n, const float* x, float scale, float* out, float* ws, void* stream);

/* ---- K10  fused Adam on the flat parameter buffer: systems/base_system.py:81-87 ---------------------------
 * g is multiplied by grad_scale first (1/world_size after the NCCL sum). step is 1-based. */
int pnb_adam_step(long long n, float* p, const float* g, float* m, float* v, float lr, float beta1, float beta2,
                  float eps, int step, float grad_scale, void* stream);
/* same update, step-dependent scalars from DEVICE memory: hyper = {lr, 1 - beta1^step, sqrt(1 - beta2^step)}
 * (lets the launch sit inside a CUDA graph while the learning-rate schedule advances) */
int pnb_adam_step_dev(long long n, float* p, const float* g, float* m, float* v, const float* hyper, float beta1,
                      float beta2, float eps, float grad_scale, void* stream);

/* ---- element-wise helpers of the MLP backward ------------------------------------------------------------- */
/* out[m,n] = src[m,n] > 0 ? w[n]*g[m] : 0   (seed of the density-Jacobian chain; g nullable => 1) */
int pnb_mask_scale(long long M, int N, const void* src, int ld_src, const float* w, const float* g, void* out,
                   int ld_out, int dtype, void* stream);
/* out = src>0 ? x : 0 */
int pnb_mask_mul(long long M, int N, const void* x, int ldx, const void* src, int ld_src, void* out, int ld_out,
                 int dtype, void* stream);
/* out[n] (+)= sum_m x[m,n] (fp32 out, atomic accumulate into pre-zeroed/accumulating buffer) */
int pnb_colsum(long long M, int N, const void* x, int ldx, int dtype, float* out, void* stream);
/* dtype conversion with strides (fp32 <-> bf16) */
int pnb_convert(long long M, int N, const void* src, int ld_src, int src_dtype, void* dst, int ld_dst,
                int dst_dtype, void* stream);

/* ---- K3/K4/K5  MLP GEMMs: models/pano_mip_nerf.py:78-114 (forward) and its hand-derived backward ----------
 * fp32 SIMT path ("parity mode"): C[M,N] = op(A) * op(B) with
 *   mode 0 (NT, forward)  : A[M,K] (lda) , B = W[N,K] (ldb)          -> C[M,N]
 *   mode 1 (NN, dgrad)    : A[M,K] (lda) , B = W[K,N] (ldb)          -> C[M,N]
 *   mode 2 (TN, wgrad)    : A[K,M] (lda) , B[K,N] (ldb), K = samples -> C[M,N]   (C is accumulated atomically)
 * epilogue: PNB_EPI_* flags; bias[N]; row_bias [M/row_group, N] (nullable); mask_src [M,N] (ld_mask). */
int pnb_gemm_f32(int mode, long long M, int N, long long K, const float* A, int lda, const float* B, int ldb,
                 float* C, int ldc, const float* bias, const float* row_bias, int row_group, const float* mask_src,
                 int ld_mask, int flags, void* stream);

/* the same three GEMM modes on bf16-stored A/B/mask (fp32 FFMA arithmetic, C bf16 or fp32): the CUDA-core twin of
 * the tensor-core path below, used to validate it on identical bf16 data. */
int pnb_gemm_bf16_simt(int mode, long long M, int N, long long K, const void* A, int lda, const void* B, int ldb,
                       void* C, int ldc, int c_dtype, const float* bias, const float* row_bias, int row_group,
                       const void* mask_src, int ld_mask, int flags, void* stream);

/* bf16 tensor-core path (tcgen05.mma, TMEM accumulators, TMA-fed):
 * C[M,Nout] = A[M,K](bf16, lda) * W[Nout,K]^T (bf16, ldw) with fp32 accumulation; epilogue flags as above;
 * C dtype = c_dtype (bf16 or fp32), bias fp32 [Nout], mask_src bf16 [M,Nout].  16 <= K <= 384 (K % 16 == 0).
 * row_bias (nullable) is an fp32 [M/row_group, Nout] addend shared by each group of `row_group` consecutive rows
 * (the per-ray view-direction term of the view layer, models/pano_mip_nerf.py:109-112).
 * colsum_out (nullable, fp32 [Nout], bf16 outputs with Nout % 64 == 0 only) is incremented by the column sums of C:
 * in the backward pass that is the bias gradient of the layer whose dZ this call produces, for free.
 * The same entry point serves forward (W) and dgrad (pre-transposed W^T). */
int pnb_linear_tc(long long M, int Nout, int K, const void* A, int lda, const void* W, int ldw, void* C, int ldc,
                  int c_dtype, const float* bias, const float* row_bias, int row_group, const void* mask_src,
                  int ld_mask, int flags, float* colsum_out, void* stream);
/* dW[Nw,Kw] (fp32, accumulated) += dZ[M,Nw]^T (bf16) * X[M,Kw] (bf16) ; reduction over the M samples.
 * Nw in {128,256}; 16 <= Kw <= 256.  workspace: pnb_wgrad_tc_workspace(Nw,Kw) bytes (per-CTA partials, summed in a
 * fixed order => deterministic). */
long long pnb_wgrad_tc_workspace(int Nw, int Kw);
int pnb_wgrad_tc(long long M, int Nw, int Kw, const void* dZ, int ldz, const void* X, int ldx, float* dW, int ldw,
                 float* workspace, void* stream);
/* All weight (and bias) gradients of one backward pass in one launch: job j computes
 * dW_j[Nw,Kw] += Z_j^T X_j and, when want_colsum, db_j[Nw] += column sums of Z_j over the M samples.
 * map_base: HOST array of n_maps DEVICE pointers to bf16 tensors [planes][M][ld]; map_desc: HOST int64 [n_maps][3] =
 * {planes, ld, cols}.  jobs: HOST int64 [n_jobs][8] = {zmap, zplane, xmap, xplane, Nw, Kw, ldw, want_colsum};
 * dW / db: HOST arrays of n_jobs DEVICE pointers (db entries may be null).  Nw in {128,256}, 16 <= Kw <= 256, Kw%16==0.
 * workspace: pnb_wgrad_batch_workspace() bytes (one slot per segment, see pnb_wgrad_batch_plan).  Partials are summed in
 * a fixed order (deterministic). */
long long pnb_wgrad_batch_workspace(void);
/* Host-only test hook: the work split of pnb_wgrad_batch for (M, jobs).  The (job, 64-sample block) units, in job-major
 * order and weighted by their operand bytes, are cut into 148 pieces of equal bytes; a CTA's piece may straddle job
 * boundaries ("segments").  out_segments: HOST int64 [max_segments][5] = {cta, job, first block, end block, slot};
 * returns the number of segments (or -1). */
int pnb_wgrad_batch_plan(long long M, int n_jobs, const long long* jobs, long long* out_segments, int max_segments);
int pnb_wgrad_batch(long long M, int n_maps, const void* const* map_base, const long long* map_desc, int n_jobs,
                    const long long* jobs, const void* const* dW, const void* const* db, float* workspace,
                    void* stream);
/* ---- fused MLP (the whole network of models/pano_mip_nerf.py:78-114 per 128-sample tile in one kernel) ------
 * Weights are consumed from a packed blob of bf16 tiles (pre-swizzled shared-memory images, in MMA order) that
 * pnb_mlp_fused_pack builds from the 24 fp32 parameters; `params_host` is a HOST array of 24 DEVICE pointers in
 * state-dict order (layers.i.0.weight, layers.i.0.bias for i=0..7, density_layer.*, extra_layer.*,
 * view_layers.0.0.*, color_layer.*).  Topology is the one of configs/*.yaml: depth 8, width 256, skip 4,
 * view width 128, 96 IPE features, C <= 16 density-head channels. */
long long pnb_mlp_fused_wblob_bytes(void);
long long pnb_mlp_fused_bblob_floats(void);
int pnb_mlp_fused_act_planes(void);
int pnb_mlp_fused_bwd_planes(void);
int pnb_mlp_fused_adj_planes(void);
/* 32-bit words of the ReLU sign bit-plane buffer: per_tile = 1 -> one record per 128-sample tile of an M-sample
 * batch (what the backward kernels read), per_tile = 0 -> a per-CTA scratch (inference with normals). */
long long pnb_mlp_fused_mask_words(long long M, int per_tile);
/* static ring plan of program `prog` (0 fwd, 1 fwd+Jacobian, 2 bwd, 3 adjoint): out[0] = weight/IPE tile loads per
 * pair of 128-sample tiles, out[1] = MMA steps per pair, then {slot, is_enc, tile, blob_off} per load. */
int pnb_mlp_fused_plan(int prog, long long* out, int cap);
int pnb_mlp_fused_pack(const void* const* params_host, int C, void* wblob, float* bblob, void* stream);
/* enc: bf16 [M,96] IPE features (row stride ld_enc); row_bias: fp32 [ceil(M/S),128] per-ray view-direction term
 * (incl. the view-layer bias) - or, with vb_mod = D > 0, fp32 [D,128] indexed by ray % D (env rays share their D
 * directions, models/pano_mip_nerf.py:337-341); outputs raw_den fp32 [M,C], raw_rgb fp32 [M,3].
 * acts (nullable): bf16 [18][M][256] planes written with TMA stores for the backward pass:
 *   0..7 trunk activations h_i, 8 bottleneck, 9 view-layer activation (cols 0..127), 10..17 Jacobian rows a_0..a_7.
 * g_enc (nullable): fp32 [M,96]; when given, the density-Jacobian sweep (pano_mip_nerf.py:295-302 without
 * vmap/jacrev) runs in the same kernel and g_enc receives d raw_sigma / d enc.
 * masks: uint32 [pnb_mlp_fused_mask_words(M, masks_per_tile)] sign bits of the 9 ReLU layers; required with g_enc,
 * required per tile when the backward kernels will run, nullable otherwise. */
int pnb_mlp_fused_fwd(long long M, int S, int C, const void* enc, int ld_enc, const void* wblob, const float* bblob,
                      const float* row_bias, int vb_mod, float* raw_den, float* raw_rgb, void* acts, float* g_enc,
                      void* masks, int masks_per_tile, void* stream);
/* Inference forward with the integrated positional encoding (models/mip.py:394-428) computed INSIDE the kernel:
 * two encoder warps per CTA evaluate the 96 features of the next tile pair from means / covs [M,3] (24 B per sample)
 * into `scratch` (pnb_mlp_fused_scratch_bytes(), per-CTA double buffer that stays in L2), the [M,96] encoding array
 * of pnb_ipe_fwd + pnb_mlp_fused_fwd never exists.  Bit-identical to that two-kernel path.  g_enc (nullable) selects
 * the Jacobian sweep as in pnb_mlp_fused_fwd (masks then required).  vb_mod != 0: the per-ray view-direction term is
 * row_bias[((m / S) % vb_mod)] - env rays share their D directions (models/pano_mip_nerf.py:337-341). */
long long pnb_mlp_fused_scratch_bytes(void);
int pnb_mlp_fused_fwd_ipe(long long M, int S, int C, const float* means, const float* covs, int min_deg,
                          const void* wblob, const float* bblob, const float* row_bias, int vb_mod, float* raw_den,
                          float* raw_rgb, float* g_enc, void* masks, void* scratch, void* stream);
/* Data-gradient chain of the backward pass (the autograd of pano_mip_nerf.py:78-114 w.r.t. activations):
 * d_rgb fp32 [M,3], d_den fp32 [M,C], masks from the forward (per tile) ->
 * dz_planes bf16 [10][M][256]: 0 dz_view (cols 0..127), 1 d_bottleneck, 2..9 dz_7..dz_0 (pre-activation gradients),
 * d_enc (nullable) fp32 [M,96] gradient w.r.t. the IPE features. */
int pnb_mlp_fused_bwd(long long M, int C, const void* wblob, const float* bblob, const float* d_rgb,
                      const float* d_den, const void* masks, void* dz_planes, float* d_enc, void* stream);
/* Adjoint of the density-Jacobian sweep (second-order terms of the normals): u bf16 [M,96] = J_ipe d_v ->
 * q_planes bf16 [8][M][256], q_i = relu'(h_i) * (q_{i-1} W_i^T) (q_5 also takes u through the skip connection). */
int pnb_mlp_fused_jadj(long long M, const void* u, int ld_u, const void* wblob, const void* masks, void* q_planes,
                       void* stream);
/* out[g, n] = sum of the `group` consecutive rows of x belonging to group g (fp32 out [M/group, N]) */
int pnb_group_sum(long long M, int N, int group, const void* x, int ldx, int dtype, float* out, void* stream);
/* head gradients as tensor-core operands: src fp32 [M,C] (C <= 16) -> dst bf16 [M,64] zero-padded; colsum (nullable)
 * fp32 [C] += column sums of src (the head's bias gradient, color_layer / density_layer of pano_mip_nerf.py:56,76) */
int pnb_pad_head_grad(long long M, int C, const float* src, void* dst_bf16, float* colsum, void* stream);
/* 1 when the tcgen05 path was compiled in and the device is sm_100 */
int pnb_tc_available(void);

/* ---- render driver + validation outputs ------------------------------------
 * pnb_pack_chw: scatter n_img per-ray results src[i] = [R, channels[i]] (HOST arrays of DEVICE pointers / ints) of
 * rays [pix0, pix0+R) into the planes of one [sum(channels), H*W] buffer - the [1,C,H,W] images that
 * systems/panonerf_system.py:171-189 builds with cat + view + permute. */
int pnb_pack_chw(long long R, long long HW, long long pix0, int n_img, const void* const* src_host,
                 const int* channels_host, float* out, void* stream);
/* (pred - gt)^2 per element, optionally times row_weights[row] (WS-PSNR solid angles): utils/metrics.py:210-237,
 * 318-326.  Sum with pnb_sum (fixed order). */
int pnb_image_sqerr(long long n, int W, long long HW, const float* pred, const float* gt, const float* row_weights,
                    float* out, void* stream);
/* Scan-line payload of an uncompressed OpenEXR file, FLOAT channels B,G,R (utils/io_exr.py:30-47) and the filtered
 * scan lines of an 8-bit RGB PNG, (x*255) truncated (utils/vis.py:25-41); chw = [C,H,W] with C in {1,3}. */
long long pnb_exr_payload_bytes(int H, int W);
int pnb_exr_pack(int H, int W, int C, const float* chw, void* out, void* stream);
long long pnb_png_payload_bytes(int H, int W);
int pnb_png_pack(int H, int W, int C, const float* chw, void* out, void* stream);

/* ---- variants the upstream hot path does not call today ------------------------
 * Specular BRDF terms per (ray, light direction): kind 0 = `microfeast_brdf` (utils/surface_rendering.py:6-61),
 * kind 1 = `blinn_phong_brdf` (:64-101).  normal [R,3], roughness [R], v [R,3]; l is [D,3] (l_per_ray = 0) or
 * [R,D,3] (l_per_ray = 1); outputs spec [R,D], nol [R,D] (clamped for kind 0, raw for kind 1, as upstream).
 * Entries upstream turns into 0 with nan_to_num are 0 with zero gradient. */
int pnb_brdf_terms_fwd(int kind, int R, int D, const float* normal, const float* roughness, const float* l,
                       int l_per_ray, const float* v, float* spec, float* nol, void* stream);
int pnb_brdf_terms_bwd(int kind, int R, int D, const float* normal, const float* roughness, const float* l,
                       int l_per_ray, const float* v, const float* g_spec, const float* g_nol, float* d_normal,
                       float* d_roughness, void* stream);
/* the `roughness is not None` branch of `surface_rendering` (utils/surface_rendering.py:147-151,159):
 * diffuse = sum_d (albedo/pi) env NoL omega, specular = sum_d spec env omega, rgb = diffuse + specular. */
int pnb_shade_sum_fwd(int R, int D, const float* env_rgb, const float* albedo, const float* spec, const float* nol,
                      const float* solid_angle, float* rgb, float* diffuse, float* specular, void* stream);
int pnb_shade_sum_bwd(int R, int D, const float* env_rgb, const float* albedo, const float* spec, const float* nol,
                      const float* solid_angle, const float* g_rgb, const float* g_diffuse, const float* g_specular,
                      float* d_env, float* d_albedo, float* d_spec, float* d_nol, void* stream);
/* `RotToTarget.rot2t` (utils/vector_rotation.py:57-89): rot[R,9] = rotation taking (0,1,0) onto tvec[r]. */
int pnb_rot_to_target_fwd(int R, const float* tvec, float* rot, void* stream);
int pnb_rot_to_target_bwd(int R, const float* tvec, const float* g_rot, float* d_tvec, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PANONERF_B200_H_ */
