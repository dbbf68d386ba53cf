// Internal declarations shared by the libnmr.so translation units (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <string>

#include "../../include/nmr.h"

namespace nmr {

constexpr int kHid = 256;        // NeRF trunk width (models/nerf.py:18)
constexpr int kHalf = 128;       // views layer width
constexpr int kLayers = 8;
constexpr int kMaxSamples = 256; // per-ray MLP evaluations a warp handles in the compositor
constexpr float kHalfPi = 1.57079637050628662109375f;  // fp32(0.5*pi), embedding.py:69
constexpr float kEps32 = 1.1920928955078125e-07f;      // torch.finfo(float32).eps

inline __host__ __device__ int round_up(int x, int m) { return (x + m - 1) / m * m; }

// Derived dimensions of one configuration.
struct Dims {
  int mip;        // 1: IPE + conical frustums, 0: classic NeRF
  int xyz;        // encoded xyz width (90 mip, 93 normal)
  int xyzp;       // padded to 16 (96)
  int dirs;       // encoded dirs width (27)
  int app;        // appearance width (0 | 16)
  int dap;        // dirs+app padded to 16 (32 | 48)
  int kx;         // xyzp + dap: row stride of the padded MLP input
  int in_dim;     // xyz + dirs + app: reference input width
  int fx, fd;     // number of frequencies
  int sc;         // coarse evaluations per ray
  int sf;         // fine evaluations per ray (mip: sc, normal: sc + num_pts_fine)
  int nf;         // normal mode: extra fine samples
  int zc, zf;     // z entries per ray (mip: S+1 fence posts)
};

// fp32 panels of one net, K padded with zero columns so every GEMM has K % 16 == 0.
struct NetF32 {
  float* W[kLayers];   // [256, Kp]; layer 0: Kp = xyzp; layer skip+1: the h-part [256,256]
  float* Wskip_x;      // [256, xyzp]  xyz-part of layer skip+1 (reference cat order is [xyz, h])
  float* b[kLayers];
  float* Wfeat; float* bfeat;       // [256,256]
  float* Walpha; float* balpha;     // [1,256]
  float* Wview_f;                   // [128,256] feature part of views_linears.0
  float* Wview_d;                   // [128,dap] dirs+app part
  float* bview;
  float* Wrgb; float* brgb;         // [3,128]
  // transposed panels for the input-gradient (dgrad) chain: g_in = g_out * W is the same "A * W'^T" GEMM with W' = W^T
  float* Wt[kLayers];  // [Kp, 256] (layer 0: [xyzp, 256], zero rows beyond xyz; layer skip+1: the h-part)
  float* Wtskip_x;     // [xyzp, 256]
  float* Wtfeat;       // [256, 256]
  float* Wtview_f;     // [256, 128]
  float* Wtview_d;     // [dap, 128]
  // the same five kinds of dgrad panels pre-split into bf16 hi + lo chunk streams for the tensor-core GEMM (device, may be null)
  const uint8_t* Pt[kLayers]; const uint8_t* Ptskip_x; const uint8_t* Ptfeat; const uint8_t* Ptview_f; const uint8_t* Ptview_d;
  bool set;
};

// ---- kernels' launch wrappers (defined in the .cu files) ---------------------------------
struct RayGenParams {
  float kinv[9]; float rot[9]; float org[3];
  int H, W, ds, h, w, mip;
  float* rays; int* miss_flag;
};
void launch_raygen(const RayGenParams& p, cudaStream_t s);
// preds["s_fine"] (renderer.py:283-285): s = t_to_s(z, min z, max z) in place over count values; scratch: 2 uint32 on the device
void launch_t_to_s(float* z, int64_t count, uint32_t* scratch, cudaStream_t s);
void launch_raygen_miss(const RayGenParams& p, cudaStream_t s);   // full-grid sphere-miss test (camera outside the unit sphere)
void launch_far_fallback(float* rays, int64_t n, int ncols, const int* miss_flag, cudaStream_t s);

struct RayOpsParams {     // get_ray_dirs / get_rays_c2w / prepare_rays_data as separate ops
  float kinv[9]; float rot[9]; float org[3];
  int H, W, flip, comp_radii, ncols; int64_t n;
  const float* dirs; const float* near; const float* far; float near_s, far_s;
  float* o; float* d; float* v;   // op 1: outputs; op 2: inputs (v may be null)
  float* out;
};
void launch_ray_op(int op, const RayOpsParams& p, cudaStream_t s);   // 0 dirs, 1 c2w, 2 prepare

struct SampleParams {
  const float* rays; int64_t n; int ncols; int64_t ray0;  // ray0: global index of first ray (Philox)
  int S; int mip; int use_disp;
  int perturb;           // classic mode: training-mode perturbation (from t_rand [n,S], or Philox when rng_mode says so)
  const float* t_rand;   // mip [n, S+1] | classic [n, S], or null
  int rng_mode; uint64_t seed;
  float* z;              // mip [n,S+1] | normal [n,S]
};
void launch_sample_coarse(const SampleParams& p, cudaStream_t s);

struct PointsParams {    // z -> pts (+var): sample_smth_along_rays outputs for the stage API
  const float* rays; int64_t n; int ncols; int S; int mip; float var_scale;
  const float* z; float* pts; float* var;
};
void launch_points(const PointsParams& p, cudaStream_t s);

struct EncodeParams {    // z -> padded fp32 MLP input rows X[P, kx]
  const float* rays; int64_t n; int ncols; int S; Dims d; float var_scale;
  const float* z;
  const int64_t* ray_id; const float* app_table; int num_frames;
  float* X;
};
void launch_encode(const EncodeParams& p, cudaStream_t s);
void launch_encode_stage(int kind, const Dims& d, const float* x, const float* var, int64_t P, float* out,
                         cudaStream_t s);
// reference-layout rows [P,in_dim] -> padded rows [P,kx]
void launch_repack_input(const Dims& d, const float* x, int64_t P, float* X, cudaStream_t s);

struct CompositeParams {
  const float* raw;        // [n,S,4]
  const float* z;          // mip [n,S+1] | normal [n,S]
  const float* rays; int ncols;   // when rays == null: rays_d given separately, no point recompute
  const float* rays_d;     // [n,3] (stage API)
  int64_t n; int S; int mip; uint32_t flags;
  const void* feats; int feats_bf16;   // [n,S,256] fp32 rows (0) | tiled 16-bit features of the tcgen05 kernel (1: bf16, 2: fp16), or null
  const float* partials;               // or: per 128-point tile and ray segment, locally composited features [tiles][4][256] (nmr_tc.cuh)
  const float* pts_in;     // [n,S,3] explicit points (stage API) or null -> recomputed from rays/z
  float* rgb; float* depth; float* acc; float* disp; float* weights;
  float* feat_out; float* pts_out; float* pt3d;
  float unnorm[12]; int has_unnorm;
  // fused hierarchical resampling for the next pass (null z_next: off)
  float* z_next; int32_t* idx_out; int nf;
  const float* u_jitter; int rng_mode; uint64_t seed; int64_t ray0; float resample_padding;
  const float* weights_in; // stage API: resample from given weights instead of computing them
  const float* sigma_noise; // [n,S] added to the raw density before the ReLU (training mode), or null
  int perturb;             // classic mode: u_jitter [n,nf] holds the u draws of sample_pdf(det=False) (or Philox)
};
void launch_composite(const CompositeParams& p, cudaStream_t s);

struct GemmParams {        // C = act(A1*W1^T + A2*W2^T + bias [+ r1a (x) r1b] [+ add]) [* (mask > 0)]
  const float* A1; int lda1; const float* W1; int K1;
  const float* A2; int lda2; const float* W2; int K2;   // optional second segment
  const float* bias; float* C; int ldc; int64_t M; int N; int act;  // 0 none, 1 relu, 2 sigmoid
  // backward-pass epilogue terms (all optional): rank-1 update r1a[row * ldr1] * r1b[col] (the sigma head's input gradient),
  // an added matrix (gradient arriving at a tapped layer / a second gradient path), and the ReLU mask of the forward
  // activation the result is the gradient of
  const float* r1a; int ldr1; const float* r1b;
  const float* add; int ldadd;
  const float* mask; int ldm;
  // ... or that mask as the bits the tcgen05 forward kernel emitted (nmr_tc.cuh: kTcMaskWordsPerTile words per 128-point tile):
  // output row r is point mask_row0 + r, output column c is bit (c & 31) of word [tile][mask_step][c >> 5][row]
  const uint32_t* mask_bits; int64_t mask_row0; int mask_step;
  // tensor-core kernel only: W1 / W2 already split into bf16 hi + lo and laid out chunk by chunk in the UMMA K-major
  // SWIZZLE_64B form (pack_split_panel): fetched with one bulk copy per K-chunk instead of being converted by every CTA
  const uint8_t* W1p; const uint8_t* W2p;
};
// Host-side packing of an fp32 panel W [N, K] (row-major, ld = K; N <= 256, K % 8 == 0) for GemmParams::W1p / W2p:
// ceil(K / 32) chunks of 32 KB, each [256 x 32] bf16 hi block (16 KB) followed by the lo block, rows >= N and columns >= K zero.
size_t packed_panel_bytes(int K);
void pack_split_panel(const float* W, int N, int K, uint8_t* out);
__host__ __device__ inline int64_t mask_word_index(int64_t pt, int step, int chunk) {
  return ((pt >> 7) * 9 + step) * 1024 + chunk * 128 + (pt & 127);
}
void launch_sgemm(const GemmParams& p, cudaStream_t s);
// same contract on the tensor cores (nmr_gemm_tc.cu: fp32 operands split into bf16 hi + lo, three tcgen05 MMAs per K step)
bool gemm_tc_supported(const GemmParams& p);
int launch_gemm_tc(const GemmParams& p, cudaStream_t s);

// ---- backward pieces of the differentiable render (nmr_backward.cu) ----
struct CompositeBwdParams {
  const float* raw; const float* z; const float* rays_d; int64_t n; int S; int mip; uint32_t flags;
  const float* grad_rgb; const float* grad_depth; const float* grad_acc; const float* grad_weights;   // any may be null
  float* grad_raw;        // [n,S,4]
  float* grad_rays_d;     // [n,3] or null
};
void launch_composite_backward(const CompositeBwdParams& p, cudaStream_t s);
void launch_encode_backward(int kind, const Dims& d, const float* x, const float* var, const float* gout, int64_t P, float* gx,
                            cudaStream_t s);
// V: views hidden activations (mask = V > 0) or null -> mask_bits (step 8 of the forward kernel's masks, rows from row0)
void launch_rgb_head_backward(const float* grad_raw, const float* raw, const float* V, const uint32_t* mask_bits, int64_t row0,
                              const float* wrgb, int64_t P, float* Gv, cudaStream_t s);
void launch_unpack_grad(const Dims& d, const float* Gx, const float* Gd, int ldgd, int64_t P, float* out, cudaStream_t s);

// peer-memory exchange (nmr_peer.cu)
void launch_peer_signal(uint32_t* flag, uint32_t value, cudaStream_t s);
void launch_peer_wait(const uint32_t* flags, int count, uint32_t value, int* status, cudaStream_t s);

// coarse matching (nmr_match.cu)
struct MatchParams {
  const float* sim; int m, n;
  const uint8_t* im_mask; const uint8_t* pt_mask;
  float temperature; int temp_div;
  float* rmax; float* rsum;          // [m] softmax over n (dim=2)
  float* cpmax; float* cpsum;        // [R][n] partial softmax stats over m (dim=1)
  float* cmax; float* csum;          // [n]
  float* conf;                       // [m,n] or null
  float* rcmax; float* ccpmax; float* ccmax;   // max of conf per row [m], partial per column [R][n], per column [n]
  int R;
  int mutual; float threshold;
  int* flag; int* jsel; float* csel; // per row
  int64_t* i_ids; int64_t* j_ids; float* mconf; int* count;
};

void launch_normalize_rows(const float* x, int rows, int d, int dp, float* out, cudaStream_t s);
int launch_match_passes(const MatchParams& p, cudaStream_t s);

// Philox4x32-10 -> uniform [0,1) with 24 random bits (same granularity as torch.rand fp32)
__device__ __forceinline__ uint32_t mulhilo32(uint32_t a, uint32_t b, uint32_t* hi) {
  uint64_t p = (uint64_t)a * b;
  *hi = (uint32_t)(p >> 32);
  return (uint32_t)p;
}
__device__ __forceinline__ float philox_uniform(uint64_t seed, uint64_t index, uint32_t stream) {
  uint32_t c0 = (uint32_t)index, c1 = (uint32_t)(index >> 32), c2 = stream, c3 = 0x6e6d7221u;
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0, hi1;
    uint32_t lo0 = mulhilo32(0xD2511F53u, c0, &hi0);
    uint32_t lo1 = mulhilo32(0xCD9E8D57u, c2, &hi1);
    uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return (float)(c0 >> 8) * (1.0f / 16777216.0f);
}

// torch.linspace(0, end, steps)[i] as the CPU kernel computes it (probed: lower half i*step,
// upper half fma(-step, steps-1-i, end)); used for tau (render_utils.py:110, 434) and u (:244, 493).
__device__ __forceinline__ float torch_linspace(float end, int steps, int i) {
  if (steps <= 1) return 0.0f;
  float step = __fdiv_rn(end, (float)(steps - 1));
  return (i < steps / 2) ? __fmul_rn((float)i, step) : __fmaf_rn(-step, (float)(steps - 1 - i), end);
}

}  // namespace nmr
