/* chs.h — C ABI of the B200-native CasualHDRSplat image-formation library (libchs.so).
 *
 * Drop-in boundary (decision D1).  The reference repository defines no
 * FFI / plugin / operator interface for this path — it ships no code at all
 * (/root/reference/Readme.md:57 "Still working on....") — so each entry point below cites the
 * statement of the reference that *requires* the function instead of an interface it replaces:
 *
 *   chs_spline_*    "Trajectory control knots", "Camera motion spline", "Virtual camera pose",
 *                   "Exposure time range"        /root/reference/assets/pipeline.png (Readme.md:50)
 *   chs_project_*   "train 3DGS ... HDR scene"   /root/reference/Readme.md:54
 *   chs_bin_*       tile binning, 64-bit depth|tile keys, radix sort (BASELINE.json north_star part 3)
 *   chs_blend_*     "Virtual sharp HDR Image", "Blur from virtual sharp images", "Auto exposure
 *                   time", "implicit CRF"        /root/reference/assets/pipeline.png
 *   chs_crf_bwd     "joint estimation of camera motion, exposure time, and camera response curve"
 *                                                /root/reference/Readme.md:54
 *   chs_sh_*        view-dependent HDR colour in front of the path; the reference's
 *                   "3DGS" (Readme.md:54) evaluates spherical harmonics per camera
 *   chs_loss, chs_ssim_loss, chs_adam_step
 *                   "Jointly optimize" (assets/pipeline.png): the photometric loss on the blurred LDR frame B_i and
 *                   the optimizer step behind the path
 *   chs_rasterize_* the whole operator in two calls
 *   chs_nvls_allreduce, chs_comm_*, chs_allreduce_grads
 *                   gradient all-reduce of the frame-sharded step (BASELINE.json north_star)
 *
 * Conventions
 *  - extern "C", plain pointers and sizes only.  Every pointer is a DEVICE pointer unless its name
 *    ends in _host.  All float tensors are contiguous fp32, all index tensors int32/uint32.
 *  - The caller owns every buffer (inputs, outputs, scratch); the library never allocates device
 *    memory.  Scratch sizes come from chs_workspace_query().
 *  - Every function returns 0 on success or a negative chs_status; chs_last_error() returns a
 *    thread-local message.  Nothing throws or exits.
 *  - Functions are stream-ordered on the cudaStream_t passed as `stream` (void* to keep CUDA headers
 *    out of this file) and re-entrant.  There is no CPU fallback: without a CUDA device every compute
 *    entry point returns CHS_ERR_CUDA.
 *  - Shapes: N Gaussians, B frames, n virtual poses per frame, C = B*n cameras, camera c = i*n + k,
 *    P = width*height, tiles = ceil(W/16)*ceil(H/16), M = number of (camera, tile, Gaussian)
 *    intersections (data dependent, < 2^32 - 1 per call).
 */
#ifndef CHS_H_
#define CHS_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CHS_VERSION 200 /* 0.2.0 */

#if defined(__GNUC__)
#define CHS_API __attribute__((visibility("default")))
#else
#define CHS_API
#endif

typedef enum chs_status {
  CHS_OK = 0,
  CHS_ERR_INVALID_ARG = -1,
  CHS_ERR_CUDA = -2,
  CHS_ERR_NCCL = -3,
  CHS_ERR_WORKSPACE_TOO_SMALL = -4,
  CHS_ERR_UNSUPPORTED = -5
} chs_status;

/* Camera response F_theta, per channel, shared by all cameras.
 *   CHS_CRF_IDENTITY  F(X) = X
 *   CHS_CRF_MLP       z = ln(X + 1e-5), h = relu(w1 z + b1), y = sigmoid(w2 . h + b2)
 *   CHS_CRF_LUT       piecewise-linear table over log exposure: u = clamp((z - z_min) / (z_max - z_min), 0, 1) (L - 1),
 *                     y = lerp(v_floor(u), v_floor(u)+1, frac(u)); z_min / z_max are fixed (zero gradient) */
enum { CHS_CRF_IDENTITY = 0, CHS_CRF_MLP = 1, CHS_CRF_LUT = 2 };
enum { CHS_SPLINE_LINEAR = 0, CHS_SPLINE_CUBIC = 1 };
/* CHS_SORT_KEY64: emit 64-bit cam|tile|depth keys and radix-sort them (the literal A.4 algorithm).
 * CHS_SORT_DEPTH_PRESORT: sort the C*N Gaussians by (cam, depth) once, emit intersections in that
 * order and stable-radix-sort only the cam|tile bits.  Both produce bit-identical outputs. */
enum { CHS_SORT_KEY64 = 0, CHS_SORT_DEPTH_PRESORT = 1 };

typedef struct chs_config {
  int32_t n_gauss;            /* N */
  int32_t n_frames;           /* B (frames handled by this call / this GPU) */
  int32_t n_virtual;          /* n virtual poses per frame */
  int32_t width, height;
  int32_t tile_size;          /* must be 16 */
  float near_plane, far_plane, eps2d;
  int32_t crf_kind;           /* CHS_CRF_* */
  int32_t crf_hidden;         /* size of the learned CRF: Hd of the MLP (params [3, 3*Hd+1], Hd <= 128) or the number of
                               * knots L of the LUT (params [3, L+2] = [z_min | z_max | v_0..v_{L-1}], 2 <= L <= 1024) */
  int32_t crf_before_average; /* 0 = F(dt * mean_k H_k) (default, D0); 1 = mean_k F(dt * H_k) */
  int32_t ks_per_camera;      /* 0: Ks is [B,3,3]; 1: Ks is [C,3,3] */
  int32_t sort_mode;          /* CHS_SORT_* */
  float background[3];
  int32_t rgbo_per_camera;    /* 0: rgbo is [N] (per Gaussian); 1: rgbo is [C,N] (per camera, e.g. from chs_sh_fwd) */
  int32_t tight_bounds;       /* 0: classic square 3-sigma tile bounds (radii = r); 1: opacity-aware per-axis bounds — the
                               * axis-aligned box of the alpha >= 1/255 ellipse, intersected with the 3-sigma square; radii
                               * entries are then PACKED rx | ry << 16.  Same images and gradients, ~1/3 fewer intersections */
  int32_t pose_fused;         /* SURVEY.md 8(f) row f1.  0: one tile list per (camera, tile), the A.4 contract (default).
                               * 1: the n virtual poses of a frame share ONE tile list per (frame, tile): a Gaussian is binned once
                               * per frame into the union of its n per-pose tile rectangles and ordered by its depth at the frame's
                               * middle pose (k = n / 2); every pose still evaluates its own projection (mean2d, conic) per pixel.
                               * Binning work drops n-fold; images differ from pose_fused = 0 only where the depth order of two
                               * overlapping Gaussians differs between the middle pose and pose k (oracle: pose_fused=True).
                               * Shapes: tiles_touched [B,N] (size of the union rectangle), order / isect_offsets [B*N], list entries
                               * frame * N + g, tile_offsets [B*tiles + 1]; last_id indexes the frame's list.  A pose that does not
                               * see a Gaussian gets a record whose alpha is 0 everywhere.  Needs CHS_SORT_DEPTH_PRESORT with the
                               * default binning route and the round-2 blend kernels. */
  /* Development knobs (0 = the measured-best default everywhere).  They select between bit-/tolerance-equivalent kernel
   * instantiations and never change results beyond atomics order; see DESIGN.md section 7. */
  int32_t tune_blend_fwd;     /* 27: ungrouped pair loop; 46 / 48: grouped at 6 / 8 CTAs per SM; 26 / 28: ungrouped at 6 / 8; 8: round-1 kernel */
  int32_t tune_blend_bwd;     /* 56 / 58: default kernel at 6 / 8 CTAs per SM; 64 / 65: 16 table rows; 3: unstaged phase A; 46-48: phase B on the
                               * tensor cores; 2: round-1 tabled kernel; 1 / 22: direct kernels */
  int32_t tune_crf_bwd;       /* resident blocks per SM (2, 3, 4); + 10: the per-unit MLP kernel instead of the interval form */
  int32_t tune_bin;           /* CHS_SORT_DEPTH_PRESORT routes.  0: banded placement (default); 3: hand-written two-pass radix multisplit
                               * over emitted intersections; 1: cub::DeviceRadixSort baseline; 2: round-1 counting placement */
  int32_t tune_bin_chunk;     /* counting placement: pairs per chunk */
  int32_t tune_project_bwd;   /* resident blocks per SM of K9 (2, 3, 4) */
  int32_t reserved[1];
} chs_config;

typedef struct chs_workspace_sizes {
  uint64_t bin_count_bytes; /* scratch for chs_bin_count */
  uint64_t bin_sort_bytes;  /* scratch for chs_bin_sort at the queried M */
  uint64_t reduce_bytes;    /* fp64 accumulators for chs_crf_bwd / chs_project_bwd / chs_spline_bwd */
} chs_workspace_sizes;

CHS_API int chs_version(void);
CHS_API const char* chs_last_error(void);
/* sizeof() of the ABI structs as this library was compiled: 0 chs_config, 1 chs_workspace_sizes, 2 chs_tensors (0 for any
 * other value).  A binding checks its own struct layout against these before the first call. */
CHS_API uint64_t chs_sizeof(int32_t which);
/* Number of hand-written kernels this process has launched so far (CUB passes and memsets excluded). */
CHS_API uint64_t chs_launch_count(void);

/* Scratch sizes for a configuration and an intersection count M (pass 0 if not yet known:
 * bin_sort_bytes is then 0). n_knots is only used for reduce_bytes. */
CHS_API int chs_workspace_query(const chs_config* cfg, int64_t n_isect, int32_t n_knots, chs_workspace_sizes* out);

/* ---- K0: SE(3) spline -> virtual camera poses (fp64 arithmetic on device) ---------------------
 * knots [K,7] camera-to-world (t, q wxyz); sample times t_i + (k/(n-1) - 1/2) dt_i.
 * viewmats [C,4,4] world-to-camera, row major. */
CHS_API int chs_spline_fwd(int32_t kind, const float* knots, int32_t n_knots, double knot_t0, double knot_dt,
                   const float* frame_times, const float* exposure, int32_t n_frames, int32_t n_virtual,
                   float* viewmats, void* stream);
/* v_viewmats [C,4,4] -> v_knots [K,7], v_frame_times [B], v_exposure [B] (window path only; the
 * brightness path comes from chs_crf_bwd).  Outputs are overwritten.  workspace: reduce_bytes. */
CHS_API int chs_spline_bwd(int32_t kind, const float* knots, int32_t n_knots, double knot_t0, double knot_dt,
                   const float* frame_times, const float* exposure, int32_t n_frames, int32_t n_virtual,
                   const float* v_viewmats, float* v_knots, float* v_frame_times, float* v_exposure,
                   void* workspace, uint64_t workspace_bytes, void* stream);

/* ---- K1: projection + EWA ------------------------------------------------------------------------
 * Outputs, all [C,N] (camera major):
 *   geom          float4 (mean2d.x, mean2d.y, conic.A, conic.B)
 *   conic_c       float  conic.C
 *   depths        float  camera-space z
 *   radii         int32  ceil(3 sqrt(lambda_max)), 0 = culled
 *   tiles_touched int32
 *   rgbo [N]      float4 (r, g, b, opacity) — the per-Gaussian record the blend kernels gather */
CHS_API int chs_project_fwd(const chs_config* cfg, const float* means, const float* quats, const float* scales,
                    const float* opacities, const float* colors, const float* viewmats, const float* Ks,
                    float* geom, float* conic_c, float* depths, int32_t* radii, int32_t* tiles_touched,
                    float* rgbo, void* stream);

/* ---- K9: projection backward -------------------------------------------------------------------
 * v_geom float4 [C,N] = (v_mx, v_my, v_A, v_B); v_cogr float4 [C,N] = (v_C, v_opacity, v_r, v_g);
 * v_blue float [C,N] = v_b  (all written by chs_blend_bwd).
 * grads_flat [14*N] = sections [means 3N | quats 4N | scales 3N | opacities N | colors 3N]; the one
 * buffer that is all-reduced across GPUs.  v_viewmats [C,4,4] (last row zero).  workspace: reduce_bytes. */
CHS_API int chs_project_bwd(const chs_config* cfg, const float* means, const float* quats, const float* scales,
                    const float* viewmats, const float* Ks, const int32_t* radii, const float* v_geom,
                    const float* v_cogr, const float* v_blue, float* grads_flat, float* v_viewmats,
                    void* workspace, uint64_t workspace_bytes, void* stream);

/* ---- K2: intersection count --------------------------------------------------------------------
 * isect_offsets uint32 [C*N]: exclusive scan of tiles_touched in emission order.  Only the routes that emit intersections read
 * it (CHS_SORT_KEY64, tune_bin != 0); the default banded placement of CHS_SORT_DEPTH_PRESORT needs just M and leaves the
 * buffer untouched.
 * order int32 [C*N]: emission order (identity for CHS_SORT_KEY64; the (cam, depth)-sorted
 * permutation of c*N+g for CHS_SORT_DEPTH_PRESORT).
 * n_isect_dev: device int64 receiving M.  If n_isect_host != NULL the call synchronises the stream
 * and also stores M there (the one host sync of a step, needed to size the M-length buffers). */
CHS_API int chs_bin_count(const chs_config* cfg, const int32_t* tiles_touched, const float* depths,
                  uint32_t* isect_offsets, int32_t* order, int64_t* n_isect_dev, int64_t* n_isect_host,
                  void* workspace, uint64_t workspace_bytes, void* stream);

/* ---- K3 + K4 + K5: key generation, radix sort, per-(camera, tile) offsets ------------------------
 * keys_sorted uint64 [M] (may be NULL with CHS_SORT_DEPTH_PRESORT — only tests need it),
 * vals_sorted int32 [M] (= c*N + g), tile_offsets uint32 [C*tiles + 1] (last entry = M).
 * Bit-exact contract: identical to a stable ascending sort of
 *   key = cam << (32 + tile_bits) | tile << 32 | float_as_uint(depth). */
CHS_API int chs_bin_sort(const chs_config* cfg, int64_t n_isect, const float* geom, const int32_t* radii,
                 const float* depths, const uint32_t* isect_offsets, const int32_t* order,
                 uint64_t* keys_sorted, int32_t* vals_sorted, uint32_t* tile_offsets, void* workspace,
                 uint64_t workspace_bytes, void* stream);

/* chs_bin_sort without the host knowing M: the intersection buffers (vals_sorted, the bin_sort part of the workspace) hold
 * `isect_capacity` entries, the live count is read from n_isect_dev (what chs_bin_count wrote) ON THE DEVICE, and the call
 * never synchronises.  If M exceeds the capacity the lists are truncated consistently (no out-of-bounds access); the caller
 * detects that whenever it next reads *n_isect_dev and repeats the step with larger buffers. */
CHS_API int chs_bin_sort_dev(const chs_config* cfg, int64_t isect_capacity, const int64_t* n_isect_dev, const float* geom,
                     const int32_t* radii, const float* depths, const uint32_t* isect_offsets, const int32_t* order,
                     uint64_t* keys_sorted, int32_t* vals_sorted, uint32_t* tile_offsets, void* workspace,
                     uint64_t workspace_bytes, void* stream);

/* ---- K6: blend forward + formation epilogue -----------------------------------------------------
 * Per (frame, tile): for each virtual pose, front-to-back alpha blending in linear HDR; then
 * mean over poses, x exposure, CRF.  Outputs: ldr [B,H,W,3], alpha [B,H,W], hdr_mean [B,H,W,3]
 * (saved for the backward / return_hdr), final_T [C,H,W], last_id [C,H,W] int32 (1-based index
 * into the pixel's tile list of the last accumulated Gaussian, 0 = none).
 * With cfg->crf_before_average = 1 (the order assets/pipeline.png draws: B = mean_k F(dt H_k)) the
 * `hdr_mean` buffer instead receives every pose's HDR image and must be [C,H,W,3]; chs_crf_bwd then
 * reads it per pose and writes v_hdr [C,H,W,3], which chs_blend_bwd reads per camera. */
CHS_API int chs_blend_fwd(const chs_config* cfg, const float* geom, const float* conic_c, const float* rgbo,
                  const int32_t* vals_sorted, const uint32_t* tile_offsets, const float* exposure,
                  const float* crf_params, float* ldr, float* alpha, float* hdr_mean, float* final_T,
                  int32_t* last_id, void* stream);

/* ---- K7: CRF / exposure backward -----------------------------------------------------------------
 * v_ldr [B,H,W,3] -> v_hdr [B,H,W,3] (gradient w.r.t. each pose's HDR image, = dt_i/n * v_X),
 * v_crf_params (same shape as crf_params), v_exposure [B] (brightness path).  Outputs overwritten.
 * workspace: reduce_bytes. */
CHS_API int chs_crf_bwd(const chs_config* cfg, const float* hdr_mean, const float* exposure, const float* crf_params,
                const float* v_ldr, float* v_hdr, float* v_crf_params, float* v_exposure, void* workspace,
                uint64_t workspace_bytes, void* stream);

/* ---- K8: blend backward ------------------------------------------------------------------------
 * v_alpha [B,H,W] may be NULL.  v_geom / v_cogr / v_blue are zeroed by the call, then accumulated
 * with warp-reduced vector atomics (see chs_project_bwd for their layout). */
CHS_API int chs_blend_bwd(const chs_config* cfg, const float* geom, const float* conic_c, const float* rgbo,
                  const int32_t* vals_sorted, const uint32_t* tile_offsets, const float* final_T,
                  const int32_t* last_id, const float* v_hdr, const float* v_alpha, float* v_geom,
                  float* v_cogr, float* v_blue, void* stream);

/* ---- test / verification helper: unpack 64-bit keys in emission order (CHS_SORT_KEY64 layout) ---- */
CHS_API int chs_bin_emit_keys(const chs_config* cfg, int64_t n_isect, const float* geom, const int32_t* radii,
                      const float* depths, const uint32_t* isect_offsets, const int32_t* order,
                      uint64_t* keys, int32_t* vals, void* stream);

/* ---- test / verification helper: the hand-written radix passes of the binning stage on caller-supplied keys ----
 * Stable ascending sort of the low `bits` bits of keys_in [n_seg * seg_len] INSIDE each of the n_seg segments of seg_len items
 * (segments never mix: this is how cameras stay apart in the depth presort).  vals_out receives each sorted item's original
 * index.  workspace: 8 * n + 1024 * (n_seg * ceil(seg_len / 4096) + n_seg) + 4096 bytes. */
CHS_API int chs_radix_sort_pairs(const uint32_t* keys_in, int32_t n_seg, uint32_t seg_len, int32_t bits, uint32_t* keys_out,
                         int32_t* vals_out, void* workspace, uint64_t workspace_bytes, void* stream);

/* ---- K10: gradient all-reduce over NCCL (one process per GPU) ----------------------------------
 * The NCCL library already loaded in the process is used (dlopen of libnccl.so.2).  unique_id is
 * the 128-byte ncclUniqueId created by rank 0 with chs_comm_unique_id() and distributed by the
 * caller. */
typedef struct chs_comm chs_comm;
CHS_API int chs_comm_unique_id(void* unique_id_host_128);
CHS_API int chs_comm_init(const void* unique_id_host_128, int32_t rank, int32_t world, chs_comm** out);
CHS_API int chs_allreduce_grads(chs_comm* comm, float* buf, uint64_t count, void* stream);
CHS_API int chs_comm_destroy(chs_comm* comm);

/* ---- one-shot entry points: chain K0..K6 / K7..K9(+K0 bwd) over caller-owned buffers --------------------
 * Every pointer is a device pointer with the shape documented at the staged entry point of the same
 * name.  `viewmats` is an input when spline_kind < 0 and is written by K0 otherwise.  The intersection
 * buffers (vals_sorted and the bin_sort part of `workspace`) must hold `isect_capacity` entries; if the
 * frame needs more, chs_rasterize_fwd returns CHS_ERR_WORKSPACE_TOO_SMALL with *n_isect_out set to the
 * required count (the call synchronises the stream once to learn M, exactly like chs_bin_count).
 * workspace_bytes >= max(bin_count_bytes, bin_sort_bytes(isect_capacity), reduce_bytes). */
typedef struct chs_tensors {
  /* Gaussians, cameras, formation parameters (inputs) */
  const float *means, *quats, *scales, *opacities, *colors, *Ks, *exposure, *crf_params;
  float* viewmats;            /* [C,4,4] */
  int32_t spline_kind;        /* < 0: explicit viewmats; else CHS_SPLINE_* */
  int32_t n_knots;
  const float *knots, *frame_times;
  double knot_t0, knot_dt;
  /* forward stage buffers */
  float *geom, *conic_c, *depths, *rgbo;
  int32_t *radii, *tiles_touched, *order, *vals_sorted, *last_id;
  uint32_t *isect_offsets, *tile_offsets;
  int64_t* n_isect_dev;
  int64_t isect_capacity;
  float *ldr, *alpha, *hdr_mean, *final_T;
  /* backward: upstream gradients in, parameter gradients out */
  const float *v_ldr, *v_alpha;
  float *v_hdr, *v_geom, *v_cogr, *v_blue, *grads_flat, *v_viewmats, *v_crf_params, *v_exposure;
  float *v_knots, *v_frame_times, *v_exposure_window; /* spline only */
  void* workspace;
  uint64_t workspace_bytes;
} chs_tensors;
CHS_API int chs_rasterize_fwd(const chs_config* cfg, const chs_tensors* t, int64_t* n_isect_out, void* stream);
/* n_isect: the count chs_rasterize_fwd reported.  v_exposure receives brightness + window paths. */
CHS_API int chs_rasterize_bwd(const chs_config* cfg, const chs_tensors* t, int64_t n_isect, void* stream);

/* ---- view-dependent colour from spherical harmonics ----------------------
 * colour_ch = max(0, 0.5 + sum_k sh[k][ch] Y_k(dir)), dir = normalize(mean - camera centre), degree 0..3,
 * sh_coeffs [N, (deg+1)^2, 3].  chs_sh_fwd writes the per-camera records rgbo_c float4 [C,N] =
 * (r, g, b, opacity); pass them as `rgbo` to chs_blend_fwd / chs_blend_bwd with cfg->rgbo_per_camera = 1.
 * chs_sh_bwd turns the per-(camera, Gaussian) colour gradients chs_blend_bwd wrote (v_cogr.z/.w, v_blue)
 * into v_sh [N,K,3], the extra mean gradient v_means [N,3] and the extra pose gradient v_viewmats [C,4,4]
 * (view direction path); all three outputs are overwritten, the caller adds the latter two to the
 * outputs of chs_project_bwd.  workspace: reduce_bytes. */
CHS_API int chs_sh_fwd(const chs_config* cfg, int32_t sh_degree, const float* sh_coeffs, const float* means,
               const float* opacities, const float* viewmats, float* rgbo_c, void* stream);
CHS_API int chs_sh_bwd(const chs_config* cfg, int32_t sh_degree, const float* sh_coeffs, const float* means,
               const float* viewmats, const int32_t* radii, const float* v_cogr, const float* v_blue, float* v_sh,
               float* v_means, float* v_viewmats, void* workspace, uint64_t workspace_bytes, void* stream);

/* ---- the steps either side of the path in a trainer ----------------------
 * chs_loss: photometric loss between the blurred LDR frames and the captured frames, emitting dL/dB in
 * the same pass.  kind 0: L = 0.5 * scale * sum d^2, v = scale * d;  kind 1: L = scale * sum |d|,
 * v = scale * sign(d);  d = ldr - target.  *loss_acc (device fp64) is ADDED to (zero it first). */
CHS_API int chs_loss(int32_t kind, const float* ldr, const float* target, uint64_t count, float scale, float* v_ldr,
             double* loss_acc, void* stream);
/* chs_ssim_loss: the D-SSIM loss of 3DGS trainers with its gradient, for frames laid out [n_img, H, W, 3]:
 *   L = l1_weight * mean|d| + ssim_weight * (1 - mean SSIM(ldr, target)),   v_ldr = dL/d ldr
 * SSIM per channel with the 11x11 Gaussian window (sigma 1.5), zero padding, C1 = 0.01^2, C2 = 0.03^2; means over
 * all n_img*H*W*3 values.  (l1_weight, ssim_weight) = (0.8, 0.2) is the usual 3DGS setting.  *loss_acc (device fp64)
 * is ADDED to.  workspace: 3 * n_img*H*W*3 floats (the three derivative maps kept between the two kernels). */
CHS_API int chs_ssim_loss(const float* ldr, const float* target, int32_t n_img, int32_t height, int32_t width,
                  float l1_weight, float ssim_weight, float* v_ldr, double* loss_acc, void* workspace,
                  uint64_t workspace_bytes, void* stream);
/* Adam (bias-corrected) on `count` parameters given their section of the flat gradient buffer.
 * step is 1-based; grad_scale multiplies the gradient first (e.g. 1 / global batch). */
CHS_API int chs_adam_step(float* param, const float* grad, float* m, float* v, uint64_t count, float lr, float beta1,
                  float beta2, float eps, int32_t step, float grad_scale, void* stream);

/* ---- K10 (NVLS variant): hand-written one-shot all-reduce through the NVSwitch multicast address ----
 * mc_ptr is the MULTICAST device pointer of a symmetric buffer (same offset on every rank, e.g. from
 * torch.distributed._symmetric_memory) that already holds each rank's partial sums.  Rank r reduces its
 * slice in the switch (multimem.ld_reduce.add.v4.f32) and broadcasts the sum to every rank
 * (multimem.st), so all ranks end with bit-identical buffers.  The caller must place a cross-rank
 * barrier before (all partials written) and after (all slices broadcast) this call. count in floats. */
CHS_API int chs_nvls_allreduce(float* mc_ptr, uint64_t count, int32_t rank, int32_t world, void* stream);
/* Fan-out of the floats [begin, begin + count) of this rank's copy (local_ptr, the rank's own mapping of the symmetric buffer)
 * to the same offsets on every rank (multimem.st).  Parameter replication of the frame-sharded step: each rank uploads 1/G of
 * the parameters from its host and broadcasts its slice.  Barriers before / after as for chs_nvls_allreduce. */
CHS_API int chs_nvls_broadcast(float* mc_ptr, const float* local_ptr, uint64_t begin, uint64_t count, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CHS_H_ */
