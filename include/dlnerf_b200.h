/* dlnerf_b200 — C ABI of the B200 (sm_100a) ray-rendering hot path.
 *
 * The reference (mertkiray/depth-lidar-nerf) has no FFI layer: its hot path is plain Python/PyTorch
 * (run_nerf.py / run_nerf_helpers.py).  This header is the drop-in boundary a maintainer binds with
 * ctypes; each entry point names the reference code it replaces.
 *
 * Conventions: every function returns 0 on success, -1 on an invalid argument (shape / alignment /
 * null pointer), or a positive cudaError_t from the launch.  All pointers are DEVICE pointers unless
 * the name ends in `_host`.  No allocation, no host synchronisation; `stream` is a cudaStream_t.
 * All float tensors are contiguous fp32 row-major; indices are int64.
 */
#ifndef DLNERF_B200_H
#define DLNERF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ----------------------------------------------------------------------------------------------
 * Sampling / encoding
 * -------------------------------------------------------------------------------------------- */

/* The packed ray batch of render(): replaces run_nerf.py:145-183 (unit view directions from the pre-warp
 * directions :147-152, ndc_rays run_nerf_helpers.py:320-337 with near plane `near_plane`, near/far columns,
 * concatenation :178-183).  rays_o / rays_d are [N,3] contiguous fp32; ray_batch is [N, use_viewdirs ? 11 : 8]. */
int dln_pack_rays(const float* rays_o, const float* rays_d, int N, int ndc, int H, int W, double focal,
                  float near_plane, float near, float far, int use_viewdirs, float* ray_batch, void* stream);

/* Stratified depths along each ray.  Replaces run_nerf.py:571-593 (t_vals linspace, near/far lerp or
 * lindisp, mids/upper/lower jitter).  rays[N, ray_stride] holds near at column 6 and far at column 7
 * (the packed ray_batch of run_nerf.py:178-183).  t_rand[N,S] may be null (perturb == 0). */
int dln_stratified_z(const float* rays, int ray_stride, const float* t_rand, float* z, int N, int S, int lindisp,
                     void* stream);

/* ----------------------------------------------------------------------------------------------
 * In-kernel random numbers.  The reference draws four tensors per render with torch.rand / torch.randn (stratified
 * jitter run_nerf.py:585, density noise run_nerf_helpers.py:565 twice, sample_pdf's u :509).  The `_rng` entry points
 * below generate those draws inside the consuming kernel with Philox4x32-10 instead of reading them from HBM:
 *   rng_state : DEVICE pointer to two uint64 {seed, base}
 *   rng_offset: by-value word added to `base`; the pair (seed, base + rng_offset) names one random tensor
 * Element e of a tensor is component (e & 3) of Philox block (e >> 2) with counter (block, base + rng_offset) and key
 * seed; uniforms are (x >> 8) * 2^-24 in [0, 1), normals are Box-Muller pairs of components (0,1) and (2,3).
 * The compositing forward and backward of one pass must be given the same pair.  `base` lives on the device so that
 * a captured CUDA graph draws fresh numbers on every replay: dln_rng_advance(state, inc) adds inc to it in-stream.
 * dln_rng_fill writes the very draws a kernel would use as a tensor (kind 0 uniform row-major, 1 normal row-major,
 * 2 uniform in the per-ray slot order of the <=64+64-sample fast path of dln_sample_pdf_rng) -- test infrastructure.
 * -------------------------------------------------------------------------------------------- */
int dln_rng_advance(unsigned long long* rng_state, unsigned long long inc, void* stream);
int dln_rng_fill(const unsigned long long* rng_state, unsigned long long rng_offset, int kind, float* out,
                 long long rows, int row_len, void* stream);
/* dln_stratified_z with the jitter t_rand ~ U[0,1) drawn in-kernel (run_nerf.py:585). */
int dln_stratified_z_rng(const float* rays, int ray_stride, const unsigned long long* rng_state,
                         unsigned long long rng_offset, float* z, int N, int S, int lindisp, void* stream);

/* Positional encoding gamma(x): [P,3] -> [P, 3+6L].  Replaces Embedder.embed, run_nerf_helpers.py:25-73. */
int dln_posenc(const float* x, float* out, long long P, int L, void* stream);

/* ----------------------------------------------------------------------------------------------
 * Alpha compositing (raw2outputs) and its backward
 * -------------------------------------------------------------------------------------------- */

/* Replaces raw2outputs, run_nerf_helpers.py:542-595.  raw[N,S,raw_ch] (raw_ch >= 4; channels 0..3 used),
 * noise[N,S] standard-normal draws or null, scaled by noise_std in-kernel.  weights may be null. */
int dln_composite_fwd(const float* raw, int raw_ch, const float* z_vals, const float* rays_d, const float* noise,
                      float noise_std, int white_bkgd, float* rgb_map, float* disp_map, float* acc_map,
                      float* weights, float* depth_map, int N, int S, void* stream);

/* Autograd backward of the above (run_nerf.py:1773 replays it through ATen).  Any upstream gradient
 * pointer may be null (= zero).  d_raw[N,S,raw_ch] is fully written (channels >= 4 get 0). */
int dln_composite_bwd(const float* raw, int raw_ch, const float* z_vals, const float* rays_d, const float* noise,
                      float noise_std, int white_bkgd, const float* g_rgb, const float* g_disp, const float* g_acc,
                      const float* g_weights, const float* g_depth, float* d_raw, int N, int S, void* stream);

/* Backward with the loss of run_nerf.py:1451-1466,1500-1536,1759-1761 fused in: rays [0,n_rgb) carry the
 * colour MSE against target_rgb (null = none), rays [n_rgb,N) the depth loss against target_depth
 * (null = none).  coef_rgb = 2/(3 n_rgb); coef_depth = 2 * depth_lambda * depth_importance / n_depth
 * (callers fold every scalar factor in).  depth_mode: 0 mse (:1524), 1 weighted (:1517),
 * 2 weighted+normalised by depth_norm=max(target) (:1520), 3 relative (:1522).
 * loss_sums[0] += sum (rgb-target)^2, loss_sums[1] += sum of depth terms (caller zeroes / normalises). */
int dln_composite_bwd_fused_loss(const float* raw, int raw_ch, const float* z_vals, const float* rays_d,
                                 const float* noise, float noise_std, int white_bkgd, const float* target_rgb,
                                 const float* target_depth, const float* ray_weights, int n_rgb, float coef_rgb,
                                 float coef_depth, int depth_mode, float depth_norm, float* loss_sums, float* d_raw,
                                 int N, int S, void* stream);

/* The three entry points above with the density noise N(0,1) drawn in-kernel (run_nerf_helpers.py:565) instead of read
 * from `noise`; dln_composite_bwd_fused_loss_dev additionally takes its three loss scalars from DEVICE memory
 * (coefs_dev = [coef_rgb, coef_depth, depth_norm], so a captured graph follows the depth_importance schedule of
 * run_nerf.py:1527-1532 and a device-side max(target_depth)) and accepts either a noise tensor, or rng_state, or
 * neither. */
int dln_composite_fwd_rng(const float* raw, int raw_ch, const float* z_vals, const float* rays_d,
                          const unsigned long long* rng_state, unsigned long long rng_offset, float noise_std,
                          int white_bkgd, float* rgb_map, float* disp_map, float* acc_map, float* weights,
                          float* depth_map, int N, int S, void* stream);
int dln_composite_bwd_rng(const float* raw, int raw_ch, const float* z_vals, const float* rays_d,
                          const unsigned long long* rng_state, unsigned long long rng_offset, float noise_std,
                          int white_bkgd, const float* g_rgb, const float* g_disp, const float* g_acc,
                          const float* g_weights, const float* g_depth, float* d_raw, int N, int S, void* stream);
int dln_composite_bwd_fused_loss_dev(const float* raw, int raw_ch, const float* z_vals, const float* rays_d,
                                     const float* noise, const unsigned long long* rng_state,
                                     unsigned long long rng_offset, float noise_std, int white_bkgd,
                                     const float* target_rgb, const float* target_depth, const float* ray_weights,
                                     int n_rgb, const float* coefs_dev, int depth_mode, float* loss_sums, float* d_raw,
                                     int N, int S, void* stream);

/* ----------------------------------------------------------------------------------------------
 * Hierarchical sampling
 * -------------------------------------------------------------------------------------------- */

/* Replaces sample_pdf, run_nerf_helpers.py:497-540 (the torch.searchsorted call at :524 included) and,
 * when z_merged != null, the sort(cat(z_vals, z_samples)) of run_nerf.py:636.
 *   bins:    row n at bins + n*bins_stride; n_bins entries, or (mid_from_z) n_bins+1 depths whose
 *            midpoints are the bins (run_nerf.py:632).
 *   weights: row n at weights + n*weights_stride; n_bins-1 entries (pass the interior slice's pointer).
 *   u[N,n_samples] uniform draws, or null for the deterministic linspace (det=True).
 *   samples[N,n_samples] out.  z_coarse[N,S] + z_merged[N,S+n_samples] optional merge.
 *   cdf_out[N,n_bins], inds_out[N,n_samples] (searchsorted right=True indices) optional. */
int dln_sample_pdf(const float* bins, int bins_stride, int mid_from_z, const float* weights, int weights_stride,
                   int n_bins, const float* u, int n_samples, float* samples, const float* z_coarse, int S,
                   float* z_merged, float* cdf_out, long long* inds_out, int N, void* stream);

/* dln_sample_pdf with u ~ U[0,1) drawn in-kernel (run_nerf_helpers.py:509). */
int dln_sample_pdf_rng(const float* bins, int bins_stride, int mid_from_z, const float* weights, int weights_stride,
                       int n_bins, const unsigned long long* rng_state, unsigned long long rng_offset, int n_samples,
                       float* samples, const float* z_coarse, int S, float* z_merged, float* cdf_out,
                       long long* inds_out, int N, void* stream);

/* raw2outputs of the coarse pass (dln_composite_fwd / _rng) followed by the hierarchical resampling of its weights
 * (dln_sample_pdf / _rng with the merge; run_nerf.py:600-636) in ONE launch, for S = 64 coarse and n_samples <= 64
 * new samples -- the same bits as the two calls.  noise / noise_rng_state and u / u_rng_state are alternatives
 * (both null: no density noise / deterministic linspace u); weights may be null. */
int dln_composite_resample_fwd(const float* raw, int raw_ch, const float* z_vals, const float* rays_d, const float* noise,
                               const unsigned long long* noise_rng_state, unsigned long long noise_rng_offset,
                               float noise_std, int white_bkgd, float* rgb_map, float* disp_map, float* acc_map,
                               float* weights, float* depth_map, const float* u, const unsigned long long* u_rng_state,
                               unsigned long long u_rng_offset, int n_samples, float* samples, float* z_merged, int N,
                               int S, void* stream);

/* Batched row-wise search with row broadcast: the contract of the vendored extension
 * torchsearchsorted/src/cuda/searchsorted_cuda_kernel.cu:110-142 (searchsorted.py:20-53). */
int dln_searchsorted(const float* a, int rows_a, int A, const float* v, int rows_v, int V, long long* out,
                     int side_right, void* stream);

/* ----------------------------------------------------------------------------------------------
 * NeRF MLP (run_nerf_helpers.py:77-145 forward; autograd backward) on tcgen05 tensor cores
 * -------------------------------------------------------------------------------------------- */

#define DLN_MAX_STEPS 12
#define DLN_MAX_KSLABS 6
#define DLN_SLAB_BYTES 16384 /* one [128 points x 64 features] bf16 panel, SWIZZLE_128B image */
#define DLN_TILE_ROWS 128

/* Epilogue kinds of a chain step. */
enum {
  DLN_EPI_RELU = 0,       /* h = relu(acc + b)                                  -> next activations   */
  DLN_EPI_RELU_SIGMA = 1, /* same, and sigma = <h, head> + head_bias            (alpha_linear)        */
  DLN_EPI_LINEAR = 2,     /* h = acc + b                                        (feature_linear)      */
  DLN_EPI_RELU_RGB = 3,   /* h = relu(acc + b); rgb = heads*h + b; write [rgb, sigma]                 */
  DLN_EPI_RELU_OUT = 4,   /* h = relu(acc + b); out = heads*h + b (output_linear, no view dirs)       */
  DLN_EPI_BWD_COPY = 8,   /* dZ = acc                                                                  */
  DLN_EPI_BWD_MASK = 9,   /* dZ = acc * relu'                                                          */
  DLN_EPI_BWD_MASK_SIGMA = 10 /* dZ = (acc + dsigma * head) * relu'                                   */
};

/* One GEMM step of the fused chain: acc[128 x n_out] = sum over K slabs of A_slab * W_stage^T. */
typedef struct {
  uint32_t w_off;     /* byte offset of this step's first weight stage in the packed bf16 blob        */
  uint32_t bias_off;  /* float offset of the bias vector in the fp32 blob                              */
  uint32_t head_off;  /* float offset of head weights [n_heads][n_out] (or the rank-1 vector)          */
  uint32_t head_bias_off; /* float offset of the head biases                                           */
  uint16_t n_out;     /* 256 or 128: output columns the step computes (a narrower layer is zero-padded up)   */
  uint8_t nk;         /* number of K slab entries                                                      */
  uint8_t epi;        /* DLN_EPI_*                                                                     */
  uint8_t kslab[DLN_MAX_KSLABS]; /* 0..3 activation slabs, 4 encoded position / direction / d_raw     */
  uint8_t kcnt[DLN_MAX_KSLABS];  /* K=16 MMA steps taken from that slab (4 = all 64 columns)           */
  int16_t stash_slot; /* first stash slot of the output slabs, -1 = not kept                           */
  int16_t mask_slot;  /* relu bit-mask slot written (fwd) / read (bwd), -1 = none                      */
  uint8_t n_heads;
  uint8_t n_valid32;  /* valid output columns / 32 (netwidth < 256: columns beyond are written as zeros, take no
                         bias and feed no head; head rows are n_valid wide); 0 = all n_out columns              */
  uint8_t pad_[2];
} DlnChainStep;

typedef struct {
  int32_t n_steps;
  int32_t backward;        /* 0 forward chain, 1 dgrad chain                                           */
  int32_t use_viewdirs;
  int32_t out_ch;          /* channels of raw / d_raw rows                                             */
  int32_t L_pts, L_dir;    /* frequencies of the two encodings                                         */
  int32_t stash_slots;     /* slabs kept per 128-point tile (0 = keep nothing)                         */
  int32_t mask_slots;
  /* backward prologue: d_raw -> dZ of the first backward layer through rgb_linear / output_linear */
  int32_t pro_head_off;    /* float offset of W_rgb[3][128] (or W_out[out_ch][256]) in the fp32 blob    */
  int32_t pro_mask_slot;   /* relu mask of that layer's forward output                                  */
  int32_t pro_slot;        /* first stash slot of the prologue's activation slabs                       */
  int32_t reload_step;     /* forward: after this step slab 4 is rewritten with the encoded direction   */
  int32_t pro_valid;       /* backward prologue: valid columns of that first dZ (= row pitch of the head
                              weights at pro_head_off); 0 = 128 with view directions, 256 without         */
  DlnChainStep steps[DLN_MAX_STEPS];
} DlnChainProgram;

/* Inputs of one chain launch.  Forward: either (rays, z) for the fused sampling+encoding prologue
 * (replaces run_nerf.py:595 + run_network :60-74) or x[P, x_ld] pre-encoded rows (NeRF.forward).
 * Backward: d_out[P,out_ch] in, dZ slabs to `stash`. */
typedef struct {
  long long P;             /* points (rows)                                                            */
  const float* rays;       /* [N, ray_stride]: o 0:3, d 3:6, unit view dir at vd_col (fused mode)      */
  int32_t ray_stride, vd_col;
  const float* z;          /* [N,S]                                                                    */
  int32_t S;
  const float* x;          /* [P, x_ld] encoded inputs (generic mode) or null                         */
  int32_t x_ld;
  const void* wblob;       /* packed bf16 weight stages                                                */
  const float* fblob;      /* fp32 biases / head vectors                                               */
  float* out;              /* fwd: raw[P,out_ch]            bwd: unused                                */
  const float* d_out;      /* bwd: d raw[P,out_ch]                                                     */
  void* stash;             /* [tiles][stash_slots][DLN_SLAB_BYTES] or null                             */
  uint32_t* masks;         /* [mask_slots][tiles][4][128][2] relu bit masks                            */
  long long* trace;        /* optional debug timeline: 64 entries [4 roles][2 steps][8 events] of %clock, CTA 0 only */
  const float* sem_g;      /* bwd, optional: G[ceil(P / sem_g_div), 256] added to dH of the last trunk layer
                              (DLN_EPI_BWD_MASK_SIGMA step) -- the semantic head's input gradient, dln_sem_head_bwd */
  int32_t sem_g_div;       /* points per row of sem_g (samples per ray; 1 = one row per point)                  */
  int32_t z_lindisp;       /* fused stratified sampling: sample linearly in inverse depth (run_nerf.py:575-578)  */
  /* Fused stratified sampling (forward, fused mode, CTA-pair kernel): when z_gen is set the tile prologue COMPUTES the
   * depths of its points (run_nerf.py:571-593: t_vals, near/far from rays columns 6/7, mids / upper / lower, jitter drawn
   * in-kernel as in dln_stratified_z_rng when z_rng_state is set, bin centres otherwise), uses them for the encoding and
   * writes them to z_gen[N,S] for the compositing kernels; `z` is then not read.  Same bits as dln_stratified_z(_rng). */
  float* z_gen;
  const unsigned long long* z_rng_state;
  unsigned long long z_rng_offset;
} DlnChainArgs;

/* Fused MLP chain (forward or dgrad).  prog_host / args_host are HOST structs passed by value to the
 * kernel.  Replaces NeRF.forward (run_nerf_helpers.py:113-145), batchify/run_network (run_nerf.py:50-74)
 * and the dgrad half of its autograd backward. */
int dln_mlp_chain(const DlnChainProgram* prog_host, const DlnChainArgs* args_host, int num_sms, void* stream);

/* One weight-gradient GEMM  dW[rows, cols] += A^T B  over all points, A/B read from the stashes. */
typedef struct {
  int32_t a_bwd_stash;     /* 1: A slabs come from the backward stash (always)                        */
  int32_t a_slot, a_nslab; /* dZ slabs: 1, 2 or 4 (64 output features each)                           */
  int32_t b_from_bwd;      /* 0: B from the forward stash                                              */
  int32_t b_slot, b_nslab; /* input-activation slabs (64 input features each)                          */
  int64_t dw_off;          /* float offset of dW[0,0] in the flat gradient buffer                      */
  int32_t ld;              /* row pitch of dW (fan_in of the layer)                                    */
  int32_t col_off;         /* first column written                                                     */
  int32_t n_cols;          /* valid columns of B (<= 64*b_nslab)                                       */
  int32_t row_off;         /* first A feature (row of the accumulator) stored                          */
  int32_t n_rows;          /* rows stored; accumulator row row_off+i -> dW row i                       */
  int64_t db_off;          /* float offset of the bias gradient, or -1                                 */
  int32_t db_col_off;      /* first A feature summed                                                   */
  int32_t db_n;            /* features summed                                                          */
} DlnWgradItem;

/* items_dev: DEVICE array of n_items; each item is split over `splits` CTAs along the points.
 * `partial`: null -> the CTAs add their shares to grads_flat with fp32 atomics (summation order varies from run to
 * run); else a scratch of n_items * splits * DLN_WGRAD_PARTIAL_FLOATS floats (16-byte aligned) -> every CTA stores its
 * share there and a second kernel adds the shares of an item in split order: bit-reproducible gradients. */
#define DLN_WGRAD_PARTIAL_FLOATS (256 * 256 + 256)
int dln_mlp_wgrad(const DlnWgradItem* items_dev, int n_items, int splits, const void* stash_fwd,
                  int fwd_slots, const void* stash_bwd, int bwd_slots, long long n_tiles, float* grads_flat,
                  float* partial, void* stream);

/* fp32 master parameters -> bf16 SWIZZLE_128B weight stages. */
typedef struct {
  int64_t src_off;     /* float offset of W[0,0] in the flat parameter buffer                          */
  int32_t ld;          /* row pitch of W (fan_in)                                                      */
  int32_t row0, col0;  /* top-left corner of the block                                                 */
  int32_t n_valid;     /* valid stage rows (output features, or input features if transposed)          */
  int32_t k_valid;     /* valid K columns (<= 64)                                                      */
  int32_t transposed;  /* 0: stage[n][k] = W[row0+n][col0+k]   1: stage[n][k] = W[row0+k][col0+n]      */
  int32_t n_rows;      /* stage rows (128 or 256)                                                      */
  uint32_t dst_off;    /* byte offset of the stage in the blob                                         */
} DlnPackJob;

int dln_mlp_pack_weights(const float* params_flat, const DlnPackJob* jobs_dev, int n_jobs, void* wblob,
                         void* stream);

/* Library / build information (arch string, e.g. "sm_100a"). */
const char* dln_build_info(void);

/* Image-aware inverse-depth smoothness on a rendered patch: InverseDepthSmoothnessLoss.forward, loss.py:87-133
 * (called as depth_inv_loss(acc_depth, acc_rgb), run_nerf.py:1249, :1646).  idepth [N,1,H,W], image [N,3,H,W], fp32
 * contiguous.  fwd ADDS sum|d_x idepth * w_x| to sums[0] and sum|d_y idepth * w_y| to sums[1] (the caller divides by
 * N*H*(W-1) and N*(H-1)*W and adds: the loss is the sum of the two means); bwd writes d loss / d idepth and
 * d loss / d image (either may be null), scaled by g_loss[0] (null = 1). */
int dln_inv_depth_smooth_fwd(const float* idepth, const float* image, int N, int H, int W, float* sums, void* stream);
int dln_inv_depth_smooth_bwd(const float* idepth, const float* image, int N, int H, int W, const float* g_loss,
                             float* g_idepth, float* g_image, void* stream);

/* feature_linear folded into views_linears (run_nerf_helpers.py:126-131: the feature layer has no activation, so
 * relu(W_v [W_f h + b_f ; dir] + b_v) = relu([W_v1 W_f] h + W_vd dir + [W_v1 b_f + b_v]); the chain then skips one
 * 256x256 layer in the forward pass, one dgrad step and one wgrad item).  All offsets are float offsets into the flat
 * buffers; `off_M` ([128 x 256]) and `off_bM` ([128]) are the derived operands behind the parameters.
 *   dln_mlp_fold         : M = W_v1 W_f, b' = W_v1 b_f + b_v            (after every weight update, before packing)
 *   dln_mlp_unfold_grads : grads_flat[off_M / off_bM] hold dM / db' of ONE wgrad call; adds dW_v1 += dM W_f^T + db' b_f^T,
 *                          dW_f += W_v1^T dM, db_f += W_v1^T db', db_v += db' to the parameters' gradient regions. */
int dln_mlp_fold(float* params_flat, long long off_views_w, int ld_views, long long off_feature_w,
                 long long off_feature_b, long long off_views_b, long long off_M, long long off_bM, void* stream);
int dln_mlp_unfold_grads(const float* params_flat, float* grads_flat, long long off_views_w, int ld_views,
                         long long off_feature_w, long long off_feature_b, long long off_views_b, long long off_M,
                         long long off_bM, void* stream);

/* ----------------------------------------------------------------------------------------------
 * Semantic head: semantic_linear = Linear(256,128) -> Linear(128,K) on `feature`
 * (run_nerf_helpers.py:107-111, :126-127), per-ray logits = UNWEIGHTED sum over the samples (:586-589), and the
 * cross-entropy of run_nerf.py:1541-1548.  No activation sits between the last trunk layer h and the logits, so
 * semantic(point) = Sw h + sc with Sw = W_s2 W_s1 W_f [K x 256] and sem_preds(ray) = Sw (sum_s h_s) + S sc: the head
 * runs on the activation slabs the forward chain keeps, never on a per-point 256 -> 256 -> 128 -> K stack.
 * -------------------------------------------------------------------------------------------- */
#define DLN_SEM_MAX_CLASSES 32

/* Float offsets into the flat parameter buffer (and, identically, the flat gradient buffer): the six parameter
 * tensors the head touches and its derived operands behind the parameters. */
typedef struct {
  int64_t w_f, b_f;    /* feature_linear.weight [256 x 256], .bias [256]                                  */
  int64_t w_s1, b_s1;  /* semantic_linear.0.weight [128 x 256], .bias [128]                               */
  int64_t w_s2, b_s2;  /* semantic_linear.1.weight [K x 128], .bias [K]                                   */
  int64_t A, a;        /* derived: A = W_s1 W_f [128 x 256], a = W_s1 b_f + b_s1 [128]  (gradient buffer: dA, da scratch) */
  int64_t Sw, sc;      /* derived: Sw = W_s2 A [K x 256], sc = W_s2 a + b_s2 [K]        (gradient buffer: dSw, dsc of one call) */
  int32_t K;           /* classes, 1..DLN_SEM_MAX_CLASSES                                                  */
  int32_t pad_;
} DlnSemOffsets;

/* A, a, Sw, sc from the parameters (after every weight update). */
int dln_sem_fold(float* params_flat, const DlnSemOffsets* off_host, void* stream);
/* grads_flat[Sw / sc] hold dSw / dsc of ONE backward call: adds the gradients of semantic_linear.{0,1}.{weight,bias}
 * and the head's share of feature_linear.{weight,bias} to their regions of grads_flat. */
int dln_sem_unfold_grads(const float* params_flat, float* grads_flat, const DlnSemOffsets* off_host, void* stream);
/* Groups of S consecutive points (a ray; S = 1: a point): hsum[g, 256] = sum of the kept last-trunk-layer activations
 * (slabs h_slot..h_slot+3 of every tile of the forward stash; may be null) and out[g*out_ld + k] = Sw hsum + n_g sc
 * (may be null).  Replaces semantic_linear's forward and `torch.sum(raw[..., 4:], -2)` (run_nerf_helpers.py:589).
 * Three kernels behind it: whole rays with S % 32 == 0 stage their 32-row blocks with bulk copies (HBM-bound, 71 % of
 * the measured copy bandwidth at 4096 x 128 points); S == 1 without hsum is the per-point packed-FMA kernel that fills
 * raw[..., 4:]; anything else uses per-lane 16-byte loads (a warp, or four for S >= 32, per group). */
int dln_sem_head_fwd(const void* stash_fwd, int fwd_slots, int h_slot, long long P, int S, const float* params_flat,
                     const DlnSemOffsets* off_host, float* hsum, float* out, int out_ld, void* stream);
/* dsem[g*dsem_ld + k] = d loss / d out of dln_sem_head_fwd.  Writes G[g, 256] = dsem Sw (hand it to the dgrad chain as
 * DlnChainArgs.sem_g with sem_g_div = S) and ADDS dSw / dsc into grads_flat[Sw / sc] (zero them before the call). */
int dln_sem_head_bwd(const float* dsem, int dsem_ld, const float* hsum, long long P, int S, const float* params_flat,
                     float* grads_flat, const DlnSemOffsets* off_host, float* G, void* stream);
/* Generic route (raw2outputs on a caller-supplied raw[N,S,C]): out[N, C-c0] = sum over the samples of raw[..., c0:]
 * (run_nerf_helpers.py:589, c0 = 4), and its backward d_raw[n,s,c] = g[n, c-c0] for c >= c0, 0 below. */
int dln_sample_sum(const float* raw, int C, int c0, int N, int S, float* out, void* stream);
int dln_sample_sum_bwd(const float* g, int C, int c0, int N, int S, float* d_raw, void* stream);
/* F.cross_entropy(logits[:n_rgb], target) with mean reduction (run_nerf.py:1542, :1546): ADDS the summed loss to
 * loss_sum[0] and writes dsem[N, K] = coef * (softmax - onehot) for the first n_rgb rays, 0 for the others. */
int dln_sem_ce_loss(const float* logits, int ld, const long long* target, int n_rgb, int N, int K, float coef,
                    float* dsem, float* loss_sum, void* stream);

/* ----------------------------------------------------------------------------------------------
 * Optimiser
 * -------------------------------------------------------------------------------------------- */

/* One Adam step on a flat fp32 buffer: replaces optimizer.step() of run_nerf.py:440 / :1774 (torch.optim.Adam,
 * no weight decay, no amsgrad) for all parameters of a network at once.  `step` is the 1-based step count used in
 * the bias corrections, `grad_scale` multiplies the gradient first (1/world after a summing all-reduce).  All four
 * buffers hold n floats and are 16-byte aligned. */
int dln_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n, double lr,
                  double beta1, double beta2, double eps, int step, float grad_scale, void* stream);

/* ----------------------------------------------------------------------------------------------
 * Ray generation: the pinhole-camera rays the reference builds on the host.
 * Arithmetic is the reference's, operation by operation and without FMA contraction (bit-identical to numpy / CPU
 * torch): dirs = ((x - W*.5)/focal, -(y - H*.5)/focal, -1), rays_d[k] = (dirs0 R[k][0] + dirs1 R[k][1]) + dirs2 R[k][2],
 * rays_o = c2w[:, 3].  `out_stride` is the distance in elements between consecutive rays of rays_o / rays_d (3 for
 * packed [N,3] outputs, 9 to write straight into the rows of the [N, 3, 3] (o, d, rgb) training bank).
 * -------------------------------------------------------------------------------------------- */

/* get_rays_np, run_nerf_helpers.py:285-300, for n_poses cameras at once (the list comprehension of
 * run_nerf.py:1126): c2w[n_poses, 3, 4] fp32 -> ray (pose, y, x) at index (pose*H + y)*W + x. */
int dln_gen_rays(const float* c2w, int n_poses, int H, int W, double focal, float* rays_o, float* rays_d,
                 long long out_stride, void* stream);
/* get_rays_by_coord_np, run_nerf_helpers.py:303-318: coords[N, 2] = fractional (x, y) pixel positions of the
 * LiDAR / COLMAP depth points (run_nerf.py:1171).  is_f64 = 0: c2w, coords and outputs fp32; 1: all fp64 (numpy
 * promotes to the coordinates' dtype, which is float64 in the reference's loaders). */
int dln_gen_rays_by_coord(const void* c2w, const void* coords, long long N, int H, int W, double focal, int is_f64,
                          void* rays_o, void* rays_d, long long out_stride, void* stream);
/* get_rays_cropped_feature_loss_new, run_nerf_helpers.py:430-494: the nH x nW crop at (start_w, start_h) in the
 * order of perm[n] (int64, a permutation of the flat crop indices row*nW + col, :466): rays_o / rays_d [n, 3] and
 * points[n, 2] = (row, col) inside the crop (int64, :461-470).  The caller splits at gradH*gradW (:468, :481). */
int dln_gen_rays_patch(const float* c2w, int H, int W, double focal, int start_w, int start_h, int nH, int nW,
                       const long long* perm, int n, float* rays_o, float* rays_d, long long* points, void* stream);

/* Host-only: sizeof of the six ABI structs, in the order ChainStep, ChainProgram, ChainArgs, WgradItem,
 * PackJob, SemOffsets, so a binding can verify its mirror of this header without touching a GPU. */
int dln_abi_sizes(int* out6_host);

#ifdef __cplusplus
}
#endif
#endif /* DLNERF_B200_H */
