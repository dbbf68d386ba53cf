// Interface of the bf16 tcgen05 / TMEM MLP path (nmr_mlp_tc.cu), used by nmr_api.cu.
#pragma once
#include <string>
#include <vector>

#include "nmr_internal.cuh"

namespace nmr {

// Unpadded fp32 copies of one net's nn.Linear parameters (host), input of the bf16 packer.
struct HostNet {
  std::vector<float> pts_w[kLayers], pts_b[kLayers];
  std::vector<float> feature_w, feature_b, alpha_w, alpha_b, views_w, views_b, rgb_w, rgb_b;
};

// Device-resident bf16 weight chunks (UMMA K-major, 64B-swizzled, in consumption order) and the
// fp32 vectors the epilogues need.
struct TcNet {
  void* pool = nullptr;
  size_t bytes = 0;
  bool set = false;
  int fmt = 0;                       // operand format of the chunk stream and of the tapped features: 0 = bf16, 1 = fp16
  const uint8_t* chunks = nullptr;   // weight chunk stream (device)
  const float* bias = nullptr;       // [8][256] pts_linears bias (fp32, device)
  const float* walpha = nullptr;     // [256]
  const float* wrgb = nullptr;       // [3][128]
  const float* wview_d = nullptr;    // [128][dap]  dirs+app part of views_linears.0 (fp32, device)
  const float* bview = nullptr;      // [128] views bias with the folded feature_linear bias (device)
  float balpha = 0.f;
  float brgb[3] = {0.f, 0.f, 0.f};
};

struct TcRun {
  const float* rays; int64_t n; int ncols; int S;
  const float* z;               // mip [n,S+1] | normal [n,S]
  const int64_t* ray_id; const float* app_table; int num_frames;
  float var_scale;
  int stop_layer; bool want_feat;
  bool fuse_composite;          // tile-local feature compositing inside the kernel (needs S >= 43, linear feat_comb)
  void* ws;                     // workspace from tc_ensure_workspace(d, ws_rays, ..., ws_with_feat): the layout depends on both
  int64_t ws_rays; bool ws_with_feat;
};

int tc_pack_net(const Dims& d, int skip_layer, const HostNet& hn, TcNet* out, std::string* err, int fmt = 0);
void tc_free_net(TcNet* n);
int tc_ensure_workspace(const Dims& d, int64_t chunk_rays, void** ws, size_t* cap, std::string* err, bool with_feat = true);
// Sample rows of one pass (encode + MLP fused): writes raw [n*S,4] fp32 and either tiled bf16 features or
// (fuse_composite) per-tile, per-ray-segment partial composites [tiles][4][256] fp32 into the workspace and
// returns their addresses (*partials_out is null when not fused).
int tc_run(const Dims& d, const TcNet& net, const TcRun& r, float** raw_out, const void** feat_out, const float** partials_out,
           int* launches, cudaStream_t s, std::string* err);
constexpr int kTcSegs = 4;      // partial vectors per 128-point tile
constexpr int kTcMinFuseS = 43; // at most 4 rays touch a tile when S >= 43
// Explicit padded fp32 rows X[P,kx] -> raw [P,4], feat fp32 [P,256] (stage API).  debug_steps > 0 also
// dumps the post-activation output of step (debug_steps-1) into debug_out [P,256] (unit tests).
int tc_run_rows(const Dims& d, const TcNet& net, const float* X, int64_t P, int stop_layer, float* raw, float* feat,
                void** ws, size_t* cap, int* launches, cudaStream_t s, std::string* err, int debug_steps = 0,
                float* debug_out = nullptr, uint32_t* masks = nullptr);
// ReLU masks the kernel can emit for the backward pass (nmr_stage_mlp_backward): per 128-point tile kTcMaskWordsPerTile words,
// [step 0..8][32-column chunk 0..7][row 0..127], bit c of a word = column 32 * chunk + c of that step's output is non-zero
// (steps 0..7: pts_linears.i after ReLU; step 8: views_linears.0 after ReLU, chunks 0..3).
constexpr int kTcMaskWordsPerTile = 9 * 8 * 128;

// code recorded by a bounded barrier wait before trapping (mapped host memory), 0 = none
int tc_trap_code();
int tc_read_timeline(long long* host, int n);   // clock64 trace of CTA 0 (NMR_TC_TIMELINE=1)
int tc_set_cluster(int cluster_size);   // 1 or 2

}  // namespace nmr
