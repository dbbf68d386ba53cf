/*
 * nmr.h  --  C ABI of libnmr.so, the B200-native NeRFMatch scene-feature renderer.
 *
 * The reference (nv-dvl/nerfmatch) has no FFI of its own: its boundary for this path is the
 * Python class nerfmatch/nerf/renderer.py:26 `NerfRenderer` plus a few free functions in
 * nerfmatch/nerf/render_utils.py.  Each entry point below names the reference interface it
 * replaces; nerfmatch_b200/renderer.py binds them with ctypes and re-creates that class on
 * top (INTEGRATION.md shows the binding a reference maintainer would add).
 *
 * Conventions
 *   - plain C types only; no torch / C++ types cross the boundary.
 *   - every function returns NMR_OK (0) or a negative nmr_status; nmr_last_error() gives text.
 *   - "dev" pointers are CUDA device pointers owned by the caller (torch tensors in the shim),
 *     row-major, fp32 unless stated; "host" pointers are ordinary (ideally pinned) host memory.
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream); all work is
 *     enqueued on it and the call returns without synchronising unless stated.
 *   - a handle is bound to the CUDA device current at nmr_create() and is not thread-safe.
 *   - there is NO CPU fallback: every entry point fails (NMR_ERR_CUDA) without a device.
 */
#ifndef NMR_H_
#define NMR_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NMR_ABI_VERSION 2
#if defined(__GNUC__)
#define NMR_API __attribute__((visibility("default")))
#else
#define NMR_API
#endif

typedef enum nmr_status {
  NMR_OK = 0,
  NMR_ERR_INVALID = -1,     /* bad argument / unsupported configuration                      */
  NMR_ERR_CUDA = -2,        /* CUDA runtime error (no device, launch failure, OOM)           */
  NMR_ERR_STATE = -3,       /* e.g. render before weights were set                           */
  NMR_ERR_UNSUPPORTED = -4  /* configuration valid for the reference but not built here      */
} nmr_status;

enum { NMR_EMBED_NORMAL = 0, NMR_EMBED_MIP = 1 };          /* config.embedding.type, renderer.py:40 */
enum { NMR_NET_COARSE = 0, NMR_NET_FINE = 1 };             /* nerf_coarse / nerf_fine               */
enum { NMR_PREC_FP32 = 0, NMR_PREC_BF16 = 1, NMR_PREC_FP16 = 2 };  /* SIMT fp32 (parity) | tcgen05 with bf16 | fp16 operands
                                                              (same kernel and rate; fp16: 8x smaller operand rounding,
                                                              activations saturate at 65504, weights must fit fp16)   */
enum { NMR_RNG_NONE = 0, NMR_RNG_PHILOX = 1 };             /* draws when t_rand/u_jitter are NULL   */

/* per-call flags: the attributes callers poke on the reference object */
#define NMR_FLAG_RET_PFEAT   (1u << 0)  /* NerfRenderer.ret_pfeat  (renderer.py:110, 224)            */
#define NMR_FLAG_FEAT_MAX    (1u << 1)  /* NerfRenderer.feat_comb == "max" (renderer.py:257-260)     */
#define NMR_FLAG_WHITE_BG    (1u << 2)  /* config.render.white_bg (renderer.py:36, render_utils:224) */
#define NMR_FLAG_SKIP_COARSE_FEAT (1u << 3) /* do not composite feat_coarse (no caller reads it)     */
#define NMR_FLAG_PERTURB     (1u << 4)  /* classic mode, rng_mode = PHILOX: draw the training-mode sample perturbation
                                           in the kernels (render_utils.py:118-126, 246-247)        */

/* Static configuration; mirrors what NerfRenderer.__init__ reads (renderer.py:27-114). */
typedef struct nmr_config {
  int32_t abi_version;        /* = NMR_ABI_VERSION                                              */
  int32_t embed_type;         /* NMR_EMBED_*                                                    */
  int32_t xyz_num_freqs;      /* config.embedding.xyz_num_freqs (15)                            */
  int32_t dirs_num_freqs;     /* config.embedding.dirs_num_freqs (4)                            */
  int32_t num_pts_coarse;     /* config.coarse_nerf.num_pts                                     */
  int32_t num_pts_fine;       /* config.fine_nerf.num_pts (ignored by the mip fine pass, as in
                                 render_utils.py:592-597)                                       */
  int32_t hid_dim;            /* 256                                                            */
  int32_t layer_num;          /* 8                                                              */
  int32_t skip_layer;         /* 4   (config.*_nerf.skips == [4])                               */
  int32_t app_dim;            /* 0, or 16 with config.embedding.appearance_embed                */
  int32_t stop_layer_coarse;  /* -1 (renderer.py:99 sets stop_layer on the fine net only)       */
  int32_t stop_layer_fine;    /* NerfRenderer(stop_layer=...), models/nerf.py:103-104           */
  int32_t use_disp;           /* config.render.use_disp (normal mode, render_utils.py:111-113)  */
  int32_t precision;          /* NMR_PREC_*                                                     */
  float mip_var_scale;        /* config.embedding.mip_var_scale (<=0: off), render_utils:311    */
  float resample_padding;     /* 0.01, render_utils.py:277                                      */
  int64_t max_chunk_points;   /* workspace bound: MLP evaluations in flight (0 = default)       */
} nmr_config;

/* fp32 nn.Linear parameters of one NeRF (models/nerf.py:44-64), weight = [out, in] row-major.
 * Pointers may be host or device memory (copied with cudaMemcpyDefault). */
typedef struct nmr_weight_ptrs {
  const float* pts_w[8];   /* pts_linears.{i}.weight: [256,xyz] | [256,256] | i==skip+1: [256,xyz+256] */
  const float* pts_b[8];   /* pts_linears.{i}.bias   [256]                                     */
  const float* views_w;    /* views_linears.0.weight [128, 256+dirs+app]                       */
  const float* views_b;    /* [128]                                                            */
  const float* feature_w;  /* feature_linear.weight  [256,256]                                 */
  const float* feature_b;  /* [256]                                                            */
  const float* alpha_w;    /* alpha_linear.weight    [1,256]                                   */
  const float* alpha_b;    /* [1]                                                              */
  const float* rgb_w;      /* rgb_linear.weight      [3,128]                                   */
  const float* rgb_b;      /* [3]                                                              */
} nmr_weight_ptrs;

/* One render call == NerfRenderer.render_rays(validation=True) (renderer.py:182-295) followed,
 * if `unnorm` is given, by unnormaliz_pts (utils/geometry.py:76-85) on the fine points.
 * Output pointers may be NULL (that output is skipped). */
typedef struct nmr_render_args {
  const float* rays;        /* [n, ncols]: o(3) d(3) near far viewdirs(3) [radius]             */
  int64_t n;
  int32_t ncols;            /* 12 mip, 11 normal; 8..10 -> viewdirs = rays_d (renderer.py:188) */
  uint32_t flags;           /* NMR_FLAG_*                                                      */
  const int64_t* ray_id;    /* [n] appearance ids or NULL (-> 1 for every ray, renderer.py:299) */
  const float* t_rand;      /* mip: coarse jitter [n, Sc+1] in [0,1) (render_utils.py:444), drawn at inference too;
                               classic: training-mode perturbation [n, Sc] (perturb=True, :118-126) or NULL */
  const float* u_jitter;    /* mip: fine jitter [n, Sc+1] in [0, 1/(Sc+1)-eps) (:483-485);
                               classic: the u draws of sample_pdf(det=False) [n, Nf] in [0,1) (:246-247), NULL = det */
  int32_t rng_mode;         /* NMR_RNG_* used when the two pointers above are NULL             */
  uint64_t seed;            /* Philox seed for NMR_RNG_PHILOX                                  */
  const float* unnorm;      /* HOST 3x4 row-major affine (top rows of unnorm_scene) or NULL    */

  float* feat_coarse;       /* [n,256]  preds["feat_coarse"]                                   */
  float* feat_fine;         /* [n,256]  preds["feat_fine"]  -> pt_feat                         */
  float* pts_coarse;        /* [n,3]                                                           */
  float* pts_fine;          /* [n,3]    normalised scene units                                 */
  float* rgb_coarse;        /* [n,3]                                                           */
  float* rgb_fine;          /* [n,3]    -> im_pred                                             */
  float* depth_coarse;      /* [n]                                                             */
  float* depth_fine;        /* [n]                                                             */
  float* pt3d;              /* [n,3]    unnorm @ [pts_fine;1]  (needs `unnorm`)                */
  /* intermediates for parity tests (all optional) */
  float* z_coarse;          /* mip [n,Sc+1] | normal [n,Sc]                                    */
  float* z_fine;            /* mip [n,Sc+1] | normal [n,Sc+Nf]                                 */
  float* weights_coarse;    /* [n,Sc]                                                          */
  float* weights_fine;      /* mip [n,Sc] | normal [n,Sc+Nf]                                   */
  float* raw_coarse;        /* [n,Sc,4]  (rgb after sigmoid, raw sigma)                        */
  float* raw_fine;
  int32_t* idx_fine;        /* mip [n,Sc+1] interval index | normal [n,Nf] searchsorted result */
  /* training-mode forward (renderer.py:217, 242, 283-286; nerf_trainer.py:140-147) - all optional, NULL at inference */
  const float* sigma_noise_coarse; /* [n,Sc]  randn * noise_std, added to the raw density before the ReLU (render_utils.py:191-195) */
  const float* sigma_noise_fine;   /* [n,Sf]                                                    */
  float* s_fine;            /* [n,zf]  preds["s_fine"] = t_to_s(z_fine, min z_fine, max z_fine) over this call's rays (:283-285) */
} nmr_render_args;

typedef struct nmr_handle nmr_handle;

/* ---- lifetime ------------------------------------------------------------------------- */
/* NerfRenderer.__init__ (renderer.py:27-114).  *out is NULL on failure. */
NMR_API int nmr_create(const nmr_config* cfg, nmr_handle** out);
NMR_API void nmr_destroy(nmr_handle* h);
/* Text of the last error on this handle (or the last nmr_create failure when h == NULL). */
NMR_API const char* nmr_last_error(const nmr_handle* h);
NMR_API int nmr_abi_version(void);

/* NerfRenderer.load_state_dict / .to(device) (nerf_evaluator.py:132-140): copies + repacks the
 * fp32 nn.Linear parameters of one net (fp32 K-padded panels; bf16 UMMA-major tiles for
 * NMR_PREC_BF16).  Synchronous. */
NMR_API int nmr_set_weights(nmr_handle* h, int which, const nmr_weight_ptrs* w);
/* embedding_a.weight [num_frames, app_dim] (renderer.py:104-107); host or device pointer. */
NMR_API int nmr_set_appearance(nmr_handle* h, const float* table, int32_t num_frames);
/* Switch between NMR_PREC_FP32 / NMR_PREC_BF16 / NMR_PREC_FP16 on a live handle (weights are kept in all forms). */
NMR_API int nmr_set_precision(nmr_handle* h, int precision);

/* ---- the hot path --------------------------------------------------------------------- */
/* sample_nerf_rays (render_utils.py:56-78) = get_ray_dirs + get_rays_c2w + rays_intersect_sphere
 * + prepare_rays_data + [ds/2::ds] subsampling, fused in one kernel.  K (3x3) and c2w (3x4,
 * already normalised: renderer.py:321-322) are HOST row-major.  rays_dev: [(H/ds')*(W/ds'), 12|11].
 * *far_fallback (host, optional) is set to 1 when a ray misses the unit sphere and far falls back
 * to 1 for the whole view, as the reference's try/except does (render_utils.py:62-68); reading it
 * synchronises the stream. */
NMR_API int nmr_raygen(nmr_handle* h, const float* K, const float* c2w, int32_t H, int32_t W, int32_t ds,
               float* rays_dev, int32_t* far_fallback, void* stream);

/* NerfRenderer.render_rays / forward / predict (renderer.py:182-313), device buffers. */
NMR_API int nmr_render(nmr_handle* h, const nmr_render_args* args, void* stream);
/* Same call with every pointer in `args` a HOST buffer: stages rays host->device, renders, and
 * copies the requested outputs device->host on `stream`, then synchronises.  This is the
 * end-to-end entry bench.py times as `e2e`. */
NMR_API int nmr_render_host(nmr_handle* h, const nmr_render_args* args, void* stream);
/* After nmr_render_host the device-side staging copy of every requested output is still resident: copy output
 * `output_index` (position among the output pointers of nmr_render_args: 0 feat_coarse, 1 feat_fine, 2 pts_coarse,
 * 3 pts_fine, 4 rgb_coarse, 5 rgb_fine, 6 depth_coarse, 7 depth_fine, 8 pt3d, ...) device-to-device into dst, e.g. to feed
 * the NCCL gather of the rendered feature map without re-uploading it.  Valid until the next nmr_render_host. */
NMR_API int nmr_copy_host_staging(nmr_handle* h, int32_t output_index, void* dst_dev, size_t bytes, void* stream);
/* NerfRenderer.render_novel_view (renderer.py:315-333) for a HOST pose: raygen + render +
 * un-normalisation; outputs in `args` are device pointers, args->rays is ignored. */
NMR_API int nmr_render_view(nmr_handle* h, const float* K, const float* c2w, int32_t H, int32_t W, int32_t ds,
                    const nmr_render_args* args, void* stream);

/* ---- per-stage entry points (unit parity, ncu, and the reference's free functions) ------ */

/* get_ray_dirs (nerf/render_utils.py:23-34): dirs[H,W,3] = [x, y, 1] @ inv(K)^T at integer pixel coordinates (optionally
 * times (1,-1,-1)).  K: 9 host floats.  get_rays_c2w (:37-41): rays_d = dirs @ R^T, rays_o = t, viewdirs = rays_d / |rays_d|
 * for n directions; c2w: 12 (or 16) host floats, row major.  prepare_rays_data (:81-104): rays[H*W, 8 (+3) (+1)] =
 * cat(o, d, near, far[, viewdirs][, radii]); near / far are per-ray device arrays or, when null, the scalars; radii
 * (comp_radii) need the [H,W] shape of rays_d.  All arrays are device pointers except K and c2w. */
NMR_API int nmr_ray_dirs(nmr_handle* h, const float* K, int32_t H, int32_t W, int32_t flipped_yz, float* dirs, void* stream);
NMR_API int nmr_rays_c2w(nmr_handle* h, const float* dirs, int64_t n, const float* c2w, float* rays_o, float* rays_d, float* viewdirs,
                         void* stream);
NMR_API int nmr_prepare_rays(nmr_handle* h, const float* rays_o, const float* rays_d, const float* viewdirs, const float* near,
                             float near_s, const float* far, float far_s, int32_t H, int32_t W, int32_t comp_radii, float* rays,
                             void* stream);
/* sample_smth_along_rays, coarse branch (render_utils.py:267-323 -> 405-450 | 107-129).
 * mip:    z [n,S+1], pts = means [n,S,3], var [n,S,3];   normal: z [n,S], pts [n,S,3], var NULL. */
NMR_API int nmr_stage_sample_coarse(nmr_handle* h, const float* rays, int64_t n, int32_t ncols, const float* t_rand,
                            float* z, float* pts, float* var, void* stream);
/* sample_smth_along_rays, fine branch (render_utils.py:555-615 | 132-148, 233-264).
 * mip:    z_in [n,S+1], weights [n,S] -> z_out [n,S+1], idx [n,S+1];
 * normal: z_in [n,S],   weights [n,S] -> z_out [n,S+Nf] (sorted), idx [n,Nf]. */
NMR_API int nmr_stage_resample(nmr_handle* h, const float* rays, int64_t n, int32_t ncols, const float* z_in,
                       const float* weights, const float* u_jitter, float* z_out, int32_t* idx,
                       float* pts, float* var, void* stream);
/* xyz_encoder / dirs_encoder (embedding.py:35-46, 66-84).  kind: 0 = xyz (IPE when var != NULL in
 * mip mode, Fourier in normal mode), 1 = dirs.  out [P, dim]. */
NMR_API int nmr_stage_encode(nmr_handle* h, int kind, const float* x, const float* var, int64_t P, float* out,
                     void* stream);
/* NeRF.forward (models/nerf.py:94-144): x [P, xyz+dirs+app] in the reference column order ->
 * raw [P,4] (sigmoid rgb, raw sigma), feat [P,256] (tap at stop_layer, <0: last trunk layer). */
NMR_API int nmr_stage_mlp(nmr_handle* h, int which, const float* x, int64_t P, int32_t stop_layer, float* raw,
                  float* feat, void* stream);
/* volume_render_radiance_field (render_utils.py:176-230) + feature / point compositing
 * (renderer.py:250-281).  feats / pts may be NULL.  z: mip [n,S+1] | normal [n,S]. */
NMR_API int nmr_stage_composite(nmr_handle* h, const float* raw, const float* z, const float* rays_d, int64_t n,
                        int32_t S, uint32_t flags, const float* feats, const float* pts, float* rgb,
                        float* depth, float* acc, float* disp, float* weights, float* feat_out,
                        float* pts_out, void* stream);

/* ---- multi-GPU exchange: peer-memory outputs ----------------------------------------------------------------
 * The one exchange of the path is "rendered feature map + 3D points -> the matcher's rank" (dist.gather in the reference's
 * terms, BASELINE config 4).  Every output pointer of nmr_render_args may be a PEER pointer - memory of another GPU of the
 * same box mapped over NVLink / NVSwitch - so the compositing kernel's own stores deliver the rows while later chunks are
 * still rendering.  nmr_peer_alloc / nmr_peer_open create and map such buffers across processes (CUDA IPC; one process per
 * GPU), nmr_peer_signal / nmr_peer_wait order a consumer stream on the matcher rank behind the producers' stores.
 *
 * Allocate `bytes` of zeroed device memory on this handle's device and export it (64-byte IPC handle, to be sent to the other
 * ranks with any byte transport, e.g. torch.distributed.broadcast). */
typedef struct nmr_peer_handle { unsigned char bytes[64]; } nmr_peer_handle;
NMR_API int nmr_peer_alloc(nmr_handle* h, size_t bytes, void** dev_ptr, nmr_peer_handle* out);
NMR_API int nmr_peer_free(nmr_handle* h, void* dev_ptr);
/* Map a buffer exported by ANOTHER process into this process; the returned pointer is valid in kernels and copies on this
 * handle's device (peer access is enabled on first use). */
NMR_API int nmr_peer_open(nmr_handle* h, const nmr_peer_handle* in, void** peer_ptr);
NMR_API int nmr_peer_close(nmr_handle* h, void* peer_ptr);
/* Producer: after all work enqueued on `stream` so far, publish `value` (a step counter) in the uint32 at `flag` (normally a
 * slot of a peer buffer on the matcher's rank).  Consumer: make `stream` wait until each of the `count` uint32 flags at
 * `flags` (local memory) has reached `value`; a producer that does not arrive within ~2 s sets *late (host, optional, read
 * after synchronising the stream) to 1 + its index instead of hanging the device. */
NMR_API int nmr_peer_signal(nmr_handle* h, uint32_t* flag, uint32_t value, void* stream);
NMR_API int nmr_peer_wait(nmr_handle* h, const uint32_t* flags, int32_t count, uint32_t value, void* stream);
NMR_API int nmr_peer_wait_status(nmr_handle* h);   /* 0, or 1 + index of the first late producer since the last call */

/* ---- next row: the differentiable render of the iNeRF pose refinement -----------------------------------------
 * NeRFMatchEvaluator.inerf_refinement (nerfmatch_evaluator.py:288-500) optimises the camera pose with Adam through
 *   gen_rays (:232-285, torch) -> xyz_encoder / dirs_encoder (:377-383) -> nerf_fine.forward (:395-404)
 *   -> volume_render_radiance_field (:410-420) -> MSE (:424-426) -> loss.backward() (:450),
 * with the sample positions and Gaussian variances constants (rays.detach(), :347).  These are the input gradients
 * (vector-Jacobian products) of those three stages; the shim wraps each forward / backward pair in a
 * torch.autograd.Function so the reference loop runs unchanged.  fp32 throughout.
 *
 * d enc -> d x for xyz_encoder (kind 0; IPE when var != NULL in mip mode, var is a constant) / dirs_encoder (kind 1).
 * x [P,3], grad_out [P,dim] (dim as nmr_stage_encode), grad_x [P,3]. */
NMR_API int nmr_stage_encode_backward(nmr_handle* h, int kind, const float* x, const float* var, int64_t P,
                                      const float* grad_out, float* grad_x, void* stream);
/* NeRF.forward backwards (models/nerf.py:94-144): x as in nmr_stage_mlp, grad_raw [P,4] w.r.t. (sigmoid rgb, raw sigma),
 * grad_feat [P,256] w.r.t. the tapped activations (or NULL) -> grad_x [P, xyz+dirs+app].  Either the forward activations are
 * recomputed chunk by chunk from x (raw_fwd = masks = NULL: nothing is kept between the forward and the backward call), or the
 * forward was nmr_stage_mlp_masks and its outputs raw_fwd [P,4] and masks are handed back (x may then be NULL). */
NMR_API int nmr_stage_mlp_backward(nmr_handle* h, int which, const float* x, int64_t P, int32_t stop_layer,
                                   const float* grad_raw, const float* grad_feat, const float* raw_fwd, const uint32_t* masks,
                                   float* grad_x, void* stream);
/* nmr_stage_mlp on the tcgen05 path (NMR_PREC_BF16 / NMR_PREC_FP16) that also emits the ReLU masks of the nine layers for
 * nmr_stage_mlp_backward: masks [ceil(P/128) * 9216] uint32 (per 128-point tile: [layer 0..8][32-column chunk][row], one bit per
 * activation; 288 bytes per point instead of the 8.5 KB of fp32 activations). */
NMR_API int nmr_stage_mlp_masks(nmr_handle* h, int which, const float* x, int64_t P, int32_t stop_layer, float* raw,
                                float* feat, uint32_t* masks, void* stream);
/* volume_render_radiance_field backwards (render_utils.py:176-230): inputs as nmr_stage_composite; incoming gradients of
 * rgb_map [n,3], depth_map [n], acc_map [n], weights [n,S] (each optional) -> grad_raw [n,S,4] and, optionally,
 * grad_rays_d [n,3] (through delta = dz * |rays_d|).  z is a constant. */
NMR_API int nmr_stage_composite_backward(nmr_handle* h, const float* raw, const float* z, const float* rays_d, int64_t n,
                                         int32_t S, uint32_t flags, const float* grad_rgb, const float* grad_depth,
                                         const float* grad_acc, const float* grad_weights, float* grad_raw,
                                         float* grad_rays_d, void* stream);

/* ---- next row: NeRFMatch-Mini coarse matching on the rendered features, batch size 1 ------------
 * NeRFMatcherCoarse.coarse_matching (nerfmatch_coarse_trainer.py:223-234) + extract_mutual_matches
 * (modules/extract_matches.py:13-40, inference branch).  im_feat [m,d], pt_feat [n,d]: device fp32, d <= 512;
 * im_mask [m] / pt_mask [n]: optional device uint8 (1 = valid); temp_div: 0 -> sim * temperature ("mul"),
 * 1 -> sim / temperature ("div").  Outputs (device): conf [m,n] (optional), i_ids / j_ids [m] int64 and
 * mconf [m] fp32 (first *count entries valid, increasing i like torch.where), count: int32[1]. */
NMR_API int nmr_coarse_match(nmr_handle* h, const float* im_feat, int32_t m, const float* pt_feat, int32_t n, int32_t d,
                             const uint8_t* im_mask, const uint8_t* pt_mask, float temperature, int32_t temp_div,
                             int32_t mutual, float threshold, float* conf, int64_t* i_ids, int64_t* j_ids,
                             float* mconf, int32_t* count, void* stream);

/* Test hooks for the tcgen05 path (NMR_PREC_BF16 only).  nmr_stage_mlp_probe evaluates the net on
 * x [P, xyz+dirs+app] and returns the post-activation output of MLP step `step` as fp32 [P,256]:
 * steps 0..7 = pts_linears.0..7 (+ReLU), 8 = views_linears.0 with feature_linear folded in (+ReLU,
 * first 128 columns).  nmr_debug_trap_code returns the code a bounded in-kernel barrier wait recorded before
 * trapping (0 = none); it lives in mapped host memory so it survives a failed launch. */
NMR_API int nmr_stage_mlp_probe(nmr_handle* h, int which, const float* x, int64_t P, int32_t step, float* out,
                                void* stream);
NMR_API int nmr_debug_trap_code(void);
/* With NMR_TC_TIMELINE=1 in the environment the tcgen05 kernel's CTA 0 records clock64 stamps for its first
 * 8 iterations: [iteration][step][4] = {MMA issue start, MMA issue end, epilogue start, epilogue end}. */
NMR_API int nmr_debug_timeline(long long* host, int n);
/* Tuning knob of the tcgen05 path (process-wide): 1 = every CTA issues its own MMAs, 2 = CTA pairs share every MMA
 * (tcgen05 cta_group::2: each SM streams only half of every weight chunk).  Default 1, or the NMR_TC_CG environment
 * variable; NMR_ERR_INVALID for other values or when the device cannot co-schedule pairs. */
NMR_API int nmr_set_tc_cluster(int cluster_size);

/* bf16 path, linear feat_comb, >= 43 samples per ray: composite the tapped features tile by tile INSIDE the MLP kernel
 * (one 256-vector per ray segment of a 128-point tile leaves the SM instead of 128 x 512 B of per-sample features; the
 * compositor scales each segment by the transmittance in front of it).  On by default (NMR_TC_FUSE=0 turns it off at
 * nmr_create); off: the MLP kernel writes tiled bf16 features and the compositor streams them, as the fp32 path and
 * feat_comb="max" always do. */
NMR_API int nmr_set_fuse_composite(nmr_handle* h, int enable);

/* Number of kernels this handle has launched since creation (bench.py's gpu_launches). */
NMR_API int64_t nmr_launch_count(const nmr_handle* h);
/* Milliseconds spent in the MLP kernels of the last nmr_render* call on this handle, measured with
 * CUDA events on `stream` (0 if timing is disabled).  nmr_set_timing(h,1) enables event recording. */
NMR_API int nmr_set_timing(nmr_handle* h, int enable);
NMR_API int nmr_last_stage_ms(nmr_handle* h, float* sample_ms, float* mlp_ms, float* composite_ms);

#ifdef __cplusplus
}
#endif
#endif /* NMR_H_ */
