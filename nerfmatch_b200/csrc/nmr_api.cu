// C ABI of libnmr.so (include/nmr.h): handle, weight packing, workspace and the orchestration of
// NerfRenderer.render_rays (nerfmatch/nerf/renderer.py:182-295) over ray chunks.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "nmr_internal.cuh"
#include "nmr_tc.cuh"

using namespace nmr;

namespace {
thread_local std::string g_create_error;

struct StageEvent { cudaEvent_t a, b; int stage; };

struct DevBuf {   // grow-only device buffer
  void* p = nullptr; size_t cap = 0;
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e == cudaSuccess) cap = bytes;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};
}  // namespace

struct nmr_handle {
  nmr_config cfg;
  Dims d;
  int device = 0;
  bool tc_gemm = true;          // backward pass / matcher GEMMs on the tensor cores (3 x bf16 split; NMR_GEMM_TC=0: fp32 SIMT)
  bool fuse_composite = true;   // bf16 path: composite the tapped features tile-locally inside the MLP kernel (NMR_TC_FUSE=0: off)
  NetF32 net[2];
  DevBuf wpool[2];
  DevBuf ppool[2];              // dgrad panels pre-split into bf16 hi + lo chunk streams (tensor-core GEMM, nmr_gemm_tc.cu)
  TcNet tc[2][2];               // [operand format: 0 bf16, 1 fp16][coarse, fine]
  DevBuf app;
  int num_frames = 0;
  // workspace (per ray chunk)
  int64_t ws_points = 0, ws_rays = 0;
  DevBuf X, B[4], raw, zc, zf, misc, rays_view, tcws, match, bwd;
  DevBuf stage[28];     // nmr_render_host staging
  std::string err;
  int64_t launches = 0;
  bool timing = false;
  bool auto_chunk = true;
  int* peer_status = nullptr; int* peer_status_dev = nullptr;   // mapped host int: late-producer report of nmr_peer_wait
  cudaStream_t copy_stream = nullptr;   // nmr_render_host: device->host copies overlap the next chunk's kernels
  std::vector<StageEvent> events;       // recorded, not yet read by nmr_last_stage_ms
  std::vector<StageEvent> event_pool;   // created once, re-used (no event creation inside a timed render after warm-up)
};

#define NMR_FAIL(h, code, ...)                                   \
  do {                                                           \
    char _b[512];                                                \
    snprintf(_b, sizeof(_b), __VA_ARGS__);                       \
    (h)->err = _b;                                               \
    return (code);                                               \
  } while (0)

#define NMR_CUDA(h, expr)                                                                      \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) NMR_FAIL(h, NMR_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(_e)); \
  } while (0)

static int check_launch(nmr_handle* h, const char* what, int count = 1) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) NMR_FAIL(h, NMR_ERR_CUDA, "launch %s: %s", what, cudaGetErrorString(e));
  h->launches += count;
  return NMR_OK;
}
#define NMR_LAUNCHED(h, what, ...)                          \
  do {                                                      \
    int _r = check_launch(h, what, ##__VA_ARGS__);          \
    if (_r != NMR_OK) return _r;                            \
  } while (0)

static Dims make_dims(const nmr_config& c) {
  Dims d{};
  d.mip = c.embed_type == NMR_EMBED_MIP;
  d.fx = c.xyz_num_freqs; d.fd = c.dirs_num_freqs;
  d.xyz = 6 * d.fx + (d.mip ? 0 : 3);
  d.xyzp = round_up(d.xyz, 16);
  d.dirs = 6 * d.fd + 3;
  d.app = c.app_dim;
  d.dap = round_up(d.dirs + d.app, 16);
  d.kx = d.xyzp + d.dap;
  d.in_dim = d.xyz + d.dirs + d.app;
  d.sc = c.num_pts_coarse;
  d.nf = c.num_pts_fine;
  d.sf = d.mip ? d.sc : d.sc + d.nf;
  d.zc = d.mip ? d.sc + 1 : d.sc;
  d.zf = d.mip ? d.sc + 1 : d.sf;
  return d;
}

extern "C" int nmr_abi_version(void) { return NMR_ABI_VERSION; }

extern "C" const char* nmr_last_error(const nmr_handle* h) {
  return h ? h->err.c_str() : g_create_error.c_str();
}

extern "C" int nmr_create(const nmr_config* cfg, nmr_handle** out) {
  if (out) *out = nullptr;
  if (!cfg || !out) { g_create_error = "nmr_create: null argument"; return NMR_ERR_INVALID; }
  if (cfg->abi_version != NMR_ABI_VERSION) { g_create_error = "nmr_create: ABI version mismatch"; return NMR_ERR_INVALID; }
  const nmr_config& c = *cfg;
  if (c.hid_dim != kHid || c.layer_num != kLayers || c.skip_layer != 4) {
    g_create_error = "nmr_create: only the shipped trunk (8 x 256, skips=[4]) is built";
    return NMR_ERR_UNSUPPORTED;
  }
  if (c.embed_type != NMR_EMBED_MIP && c.embed_type != NMR_EMBED_NORMAL) { g_create_error = "nmr_create: embed_type"; return NMR_ERR_INVALID; }
  if (c.xyz_num_freqs < 1 || c.xyz_num_freqs > 16 || c.dirs_num_freqs < 1 || c.dirs_num_freqs > 8) {
    g_create_error = "nmr_create: xyz_num_freqs must be 1..16 and dirs_num_freqs 1..8";
    return NMR_ERR_UNSUPPORTED;
  }
  if (c.app_dim != 0 && c.app_dim != 16) { g_create_error = "nmr_create: app_dim must be 0 or 16"; return NMR_ERR_UNSUPPORTED; }
  Dims d = make_dims(c);
  if (d.sc < 4 || d.sf > kMaxSamples || (!d.mip && d.nf < 2)) {
    g_create_error = "nmr_create: samples per ray must satisfy 4 <= coarse and (coarse + fine) <= 256";
    return NMR_ERR_UNSUPPORTED;
  }
  if (c.stop_layer_coarse >= kLayers || c.stop_layer_fine >= kLayers) { g_create_error = "nmr_create: stop_layer >= layer_num"; return NMR_ERR_INVALID; }
  if (c.precision != NMR_PREC_FP32 && c.precision != NMR_PREC_BF16 && c.precision != NMR_PREC_FP16) { g_create_error = "nmr_create: precision"; return NMR_ERR_INVALID; }
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    g_create_error = std::string("nmr_create: no CUDA device (") + cudaGetErrorString(e) + "); libnmr has no CPU fallback";
    return NMR_ERR_CUDA;
  }
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, dev);
  if (e != cudaSuccess) { g_create_error = std::string("nmr_create: ") + cudaGetErrorString(e); return NMR_ERR_CUDA; }
  if (prop.major != 10) {
    g_create_error = "nmr_create: libnmr is built for sm_100a (B200) only; found sm_" + std::to_string(prop.major) + std::to_string(prop.minor);
    return NMR_ERR_CUDA;
  }
  nmr_handle* h = new nmr_handle();
  h->cfg = c;
  if (h->cfg.resample_padding <= 0.f) h->cfg.resample_padding = 0.01f;
  h->auto_chunk = h->cfg.max_chunk_points <= 0;
  if (h->auto_chunk) h->cfg.max_chunk_points = (int64_t)1 << 20;
  h->d = d;
  h->device = dev;
  if (const char* fv = getenv("NMR_TC_FUSE")) h->fuse_composite = atoi(fv) != 0;
  if (const char* gv = getenv("NMR_GEMM_TC")) h->tc_gemm = atoi(gv) != 0;
  for (int i = 0; i < 2; ++i) { memset(&h->net[i], 0, sizeof(NetF32)); h->tc[0][i] = TcNet{}; h->tc[1][i] = TcNet{}; }
  *out = h;
  return NMR_OK;
}

extern "C" void nmr_destroy(nmr_handle* h) {
  if (!h) return;
  for (auto& ev : h->events) { cudaEventDestroy(ev.a); cudaEventDestroy(ev.b); }
  for (auto& ev : h->event_pool) { cudaEventDestroy(ev.a); cudaEventDestroy(ev.b); }
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  if (h->peer_status) cudaFreeHost(h->peer_status);
  for (int i = 0; i < 2; ++i) { h->wpool[i].release(); h->ppool[i].release(); tc_free_net(&h->tc[0][i]); tc_free_net(&h->tc[1][i]); }
  h->app.release(); h->X.release(); h->raw.release(); h->zc.release(); h->zf.release(); h->misc.release();
  h->rays_view.release(); h->tcws.release(); h->match.release(); h->bwd.release();
  for (auto& b : h->B) b.release();
  for (auto& b : h->stage) b.release();
  delete h;
}

extern "C" int64_t nmr_launch_count(const nmr_handle* h) { return h ? h->launches : 0; }

extern "C" int nmr_set_precision(nmr_handle* h, int precision) {
  if (!h) return NMR_ERR_INVALID;
  if (precision != NMR_PREC_FP32 && precision != NMR_PREC_BF16 && precision != NMR_PREC_FP16) NMR_FAIL(h, NMR_ERR_INVALID, "nmr_set_precision: %d", precision);
  h->cfg.precision = precision;
  return NMR_OK;
}

// ------------------------------------------------------------------------------------------
// weights
// ------------------------------------------------------------------------------------------
static int fetch(nmr_handle* h, std::vector<float>& dst, const float* src, size_t count, const char* name) {
  if (!src) NMR_FAIL(h, NMR_ERR_INVALID, "nmr_set_weights: %s is NULL", name);
  dst.resize(count);
  NMR_CUDA(h, cudaMemcpy(dst.data(), src, count * sizeof(float), cudaMemcpyDefault));
  return NMR_OK;
}

extern "C" int nmr_set_weights(nmr_handle* h, int which, const nmr_weight_ptrs* w) {
  if (!h || !w) return NMR_ERR_INVALID;
  if (which != NMR_NET_COARSE && which != NMR_NET_FINE) NMR_FAIL(h, NMR_ERR_INVALID, "nmr_set_weights: which=%d", which);
  NMR_CUDA(h, cudaSetDevice(h->device));
  const Dims& d = h->d;
  const int skipl = h->cfg.skip_layer + 1;
  // host-side packed pool
  std::vector<float> pool;
  auto alloc = [&](size_t count) { size_t off = pool.size(); pool.resize(off + ((count + 63) / 64) * 64, 0.f); return off; };
  size_t oW[kLayers], ob[kLayers], oWsx, oWf, obf, oWa, oba, oWvf, oWvd, obv, oWr, obr;
  std::vector<float> tmp;
  int r;
  HostNet hn;   // unpadded fp32 copies for the tensor-core packer
  for (int i = 0; i < kLayers; ++i) {
    int kin = (i == 0) ? d.xyz : (i == skipl ? d.xyz + kHid : kHid);
    if ((r = fetch(h, tmp, w->pts_w[i], (size_t)kHid * kin, "pts_w")) != NMR_OK) return r;
    hn.pts_w[i] = tmp;
    if (i == 0) {
      oW[i] = alloc((size_t)kHid * d.xyzp);
      for (int o = 0; o < kHid; ++o) memcpy(&pool[oW[i] + (size_t)o * d.xyzp], &tmp[(size_t)o * kin], d.xyz * sizeof(float));
    } else if (i == skipl) {
      oWsx = alloc((size_t)kHid * d.xyzp);
      oW[i] = alloc((size_t)kHid * kHid);
      for (int o = 0; o < kHid; ++o) {
        memcpy(&pool[oWsx + (size_t)o * d.xyzp], &tmp[(size_t)o * kin], d.xyz * sizeof(float));
        memcpy(&pool[oW[i] + (size_t)o * kHid], &tmp[(size_t)o * kin + d.xyz], kHid * sizeof(float));
      }
    } else {
      oW[i] = alloc((size_t)kHid * kHid);
      memcpy(&pool[oW[i]], tmp.data(), tmp.size() * sizeof(float));
    }
    if ((r = fetch(h, tmp, w->pts_b[i], kHid, "pts_b")) != NMR_OK) return r;
    hn.pts_b[i] = tmp;
    ob[i] = alloc(kHid);
    memcpy(&pool[ob[i]], tmp.data(), kHid * sizeof(float));
  }
  if ((r = fetch(h, tmp, w->feature_w, (size_t)kHid * kHid, "feature_w")) != NMR_OK) return r;
  hn.feature_w = tmp;
  oWf = alloc(tmp.size()); memcpy(&pool[oWf], tmp.data(), tmp.size() * sizeof(float));
  if ((r = fetch(h, tmp, w->feature_b, kHid, "feature_b")) != NMR_OK) return r;
  hn.feature_b = tmp;
  obf = alloc(kHid); memcpy(&pool[obf], tmp.data(), kHid * sizeof(float));
  if ((r = fetch(h, tmp, w->alpha_w, kHid, "alpha_w")) != NMR_OK) return r;
  hn.alpha_w = tmp;
  oWa = alloc(kHid); memcpy(&pool[oWa], tmp.data(), kHid * sizeof(float));
  if ((r = fetch(h, tmp, w->alpha_b, 1, "alpha_b")) != NMR_OK) return r;
  hn.alpha_b = tmp;
  oba = alloc(1); pool[oba] = tmp[0];
  const int vin = kHid + d.dirs + d.app;
  if ((r = fetch(h, tmp, w->views_w, (size_t)kHalf * vin, "views_w")) != NMR_OK) return r;
  hn.views_w = tmp;
  oWvf = alloc((size_t)kHalf * kHid);
  oWvd = alloc((size_t)kHalf * d.dap);
  for (int o = 0; o < kHalf; ++o) {
    memcpy(&pool[oWvf + (size_t)o * kHid], &tmp[(size_t)o * vin], kHid * sizeof(float));
    memcpy(&pool[oWvd + (size_t)o * d.dap], &tmp[(size_t)o * vin + kHid], (d.dirs + d.app) * sizeof(float));
  }
  if ((r = fetch(h, tmp, w->views_b, kHalf, "views_b")) != NMR_OK) return r;
  hn.views_b = tmp;
  obv = alloc(kHalf); memcpy(&pool[obv], tmp.data(), kHalf * sizeof(float));
  if ((r = fetch(h, tmp, w->rgb_w, 3 * kHalf, "rgb_w")) != NMR_OK) return r;
  hn.rgb_w = tmp;
  oWr = alloc(3 * kHalf); memcpy(&pool[oWr], tmp.data(), 3 * kHalf * sizeof(float));
  if ((r = fetch(h, tmp, w->rgb_b, 3, "rgb_b")) != NMR_OK) return r;
  hn.rgb_b = tmp;
  obr = alloc(4); memcpy(&pool[obr], tmp.data(), 3 * sizeof(float));
  // transposed panels for the dgrad chain (nmr_stage_mlp_backward): src [rows, cols] -> [cols, rows]
  auto transposed = [&](size_t src, int rows, int cols) {
    size_t off = alloc((size_t)rows * cols);
    for (int r2 = 0; r2 < rows; ++r2)
      for (int c2 = 0; c2 < cols; ++c2) pool[off + (size_t)c2 * rows + r2] = pool[src + (size_t)r2 * cols + c2];
    return off;
  };
  size_t oWt[kLayers];
  for (int i = 0; i < kLayers; ++i) oWt[i] = transposed(oW[i], kHid, i == 0 ? d.xyzp : kHid);
  const size_t oWtsx = transposed(oWsx, kHid, d.xyzp), oWtf = transposed(oWf, kHid, kHid),
               oWtvf = transposed(oWvf, kHalf, kHid), oWtvd = transposed(oWvd, kHalf, d.dap);

  NMR_CUDA(h, h->wpool[which].reserve(pool.size() * sizeof(float)));
  float* base = (float*)h->wpool[which].p;
  NMR_CUDA(h, cudaMemcpy(base, pool.data(), pool.size() * sizeof(float), cudaMemcpyHostToDevice));
  NetF32& n = h->net[which];
  for (int i = 0; i < kLayers; ++i) { n.W[i] = base + oW[i]; n.b[i] = base + ob[i]; }
  n.Wskip_x = base + oWsx;
  n.Wfeat = base + oWf; n.bfeat = base + obf; n.Walpha = base + oWa; n.balpha = base + oba;
  n.Wview_f = base + oWvf; n.Wview_d = base + oWvd; n.bview = base + obv; n.Wrgb = base + oWr; n.brgb = base + obr;
  for (int i = 0; i < kLayers; ++i) n.Wt[i] = base + oWt[i];
  n.Wtskip_x = base + oWtsx; n.Wtfeat = base + oWtf; n.Wtview_f = base + oWtvf; n.Wtview_d = base + oWtvd;
  {  // the same panels pre-split for the tensor-core GEMM: [N, K] fp32 -> K/32 chunks of (bf16 hi | bf16 lo) swizzled blocks
    struct Panel { size_t src; int N, K; const uint8_t** dst; };
    std::vector<Panel> panels;
    for (int i = 0; i < kLayers; ++i) panels.push_back({oWt[i], i == 0 ? d.xyzp : kHid, kHid, &n.Pt[i]});
    panels.push_back({oWtsx, d.xyzp, kHid, &n.Ptskip_x});
    panels.push_back({oWtf, kHid, kHid, &n.Ptfeat});
    panels.push_back({oWtvf, kHid, kHalf, &n.Ptview_f});
    panels.push_back({oWtvd, d.dap, kHalf, &n.Ptview_d});
    size_t total = 0;
    std::vector<size_t> offs;
    for (auto& pn : panels) { offs.push_back(total); total += packed_panel_bytes(pn.K); }
    std::vector<uint8_t> packed(total);
    for (size_t i = 0; i < panels.size(); ++i) pack_split_panel(&pool[panels[i].src], panels[i].N, panels[i].K, packed.data() + offs[i]);
    NMR_CUDA(h, h->ppool[which].reserve(total));
    NMR_CUDA(h, cudaMemcpy(h->ppool[which].p, packed.data(), total, cudaMemcpyHostToDevice));
    for (size_t i = 0; i < panels.size(); ++i) *panels[i].dst = (const uint8_t*)h->ppool[which].p + offs[i];
  }
  n.set = true;
  // 16-bit UMMA-major tiles for the tcgen05 path, in both operand formats (nmr_set_precision switches on a live handle)
  for (int fmt = 0; fmt < 2; ++fmt) {
    std::string terr;
    int tr = tc_pack_net(h->d, h->cfg.skip_layer, hn, &h->tc[fmt][which], &terr, fmt);
    if (tr < 0) NMR_FAIL(h, NMR_ERR_CUDA, "tc_pack_net: %s", terr.c_str());
    if (tr > 0 && h->cfg.precision == (fmt ? NMR_PREC_FP16 : NMR_PREC_BF16)) NMR_FAIL(h, NMR_ERR_UNSUPPORTED, "nmr_set_weights: %s", terr.c_str());
  }
  return NMR_OK;
}

extern "C" int nmr_set_appearance(nmr_handle* h, const float* table, int32_t num_frames) {
  if (!h) return NMR_ERR_INVALID;
  if (h->d.app == 0) NMR_FAIL(h, NMR_ERR_INVALID, "nmr_set_appearance: handle was created with app_dim = 0");
  if (!table || num_frames <= 0) NMR_FAIL(h, NMR_ERR_INVALID, "nmr_set_appearance: empty table");
  NMR_CUDA(h, cudaSetDevice(h->device));
  size_t bytes = (size_t)num_frames * h->d.app * sizeof(float);
  NMR_CUDA(h, h->app.reserve(bytes));
  NMR_CUDA(h, cudaMemcpy(h->app.p, table, bytes, cudaMemcpyDefault));
  h->num_frames = num_frames;
  return NMR_OK;
}

// ------------------------------------------------------------------------------------------
// workspace
// ------------------------------------------------------------------------------------------
static int ensure_workspace(nmr_handle* h, int64_t rays, int64_t points) {
  const Dims& d = h->d;
  if (points > h->ws_points) {
    NMR_CUDA(h, h->X.reserve((size_t)points * d.kx * sizeof(float)));
    for (auto& b : h->B) NMR_CUDA(h, b.reserve((size_t)points * kHid * sizeof(float)));
    NMR_CUDA(h, h->raw.reserve((size_t)points * 4 * sizeof(float)));
    h->ws_points = points;
  }
  if (rays > h->ws_rays) {
    NMR_CUDA(h, h->zc.reserve((size_t)rays * (d.zc + 1) * sizeof(float)));
    NMR_CUDA(h, h->zf.reserve((size_t)rays * (d.zf + 1) * sizeof(float)));
    h->ws_rays = rays;
  }
  NMR_CUDA(h, h->misc.reserve(256));
  return NMR_OK;
}

constexpr size_t kMaxPendingStageEvents = 4096;   // timing left on and never read: the oldest records are recycled
static void stage_begin(nmr_handle* h, int stage, cudaStream_t s) {
  if (!h->timing) return;
  if (h->events.size() >= kMaxPendingStageEvents) {
    h->event_pool.insert(h->event_pool.end(), h->events.begin(), h->events.begin() + kMaxPendingStageEvents / 2);
    h->events.erase(h->events.begin(), h->events.begin() + kMaxPendingStageEvents / 2);
  }
  StageEvent ev;
  if (!h->event_pool.empty()) {
    ev = h->event_pool.back(); h->event_pool.pop_back();
  } else if (cudaEventCreate(&ev.a) != cudaSuccess || cudaEventCreate(&ev.b) != cudaSuccess) {
    cudaGetLastError();
    return;                                        // no event, no sample: stage_end checks the marker below
  }
  ev.stage = stage;
  cudaEventRecord(ev.a, s);
  h->events.push_back(ev);
  h->events.back().stage = stage | 0x100;          // open: stage_end closes it
}
static void stage_end(nmr_handle* h, cudaStream_t s) {
  if (!h->timing || h->events.empty() || !(h->events.back().stage & 0x100)) return;
  h->events.back().stage &= 0xff;
  cudaEventRecord(h->events.back().b, s);
}

extern "C" int nmr_copy_host_staging(nmr_handle* h, int32_t output_index, void* dst_dev, size_t bytes, void* stream) {
  if (!h) return NMR_ERR_INVALID;
  if (output_index < 0 || output_index >= 17 || !dst_dev) NMR_FAIL(h, NMR_ERR_INVALID, "nmr_copy_host_staging: bad argument");
  DevBuf& b = h->stage[6 + output_index];            // slots 0..5 are the inputs of nmr_render_host
  if (!b.p || b.cap < bytes) NMR_FAIL(h, NMR_ERR_INVALID, "nmr_copy_host_staging: output %d was not produced by the last nmr_render_host (or is smaller)", output_index);
  NMR_CUDA(h, cudaSetDevice(h->device));
  NMR_CUDA(h, cudaMemcpyAsync(dst_dev, b.p, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return NMR_OK;
}

extern "C" int nmr_set_fuse_composite(nmr_handle* h, int enable) {
  if (!h) return NMR_ERR_INVALID;
  h->fuse_composite = enable != 0;
  return NMR_OK;
}

extern "C" int nmr_set_timing(nmr_handle* h, int enable) {
  if (!h) return NMR_ERR_INVALID;
  h->timing = enable != 0;
  return NMR_OK;
}

extern "C" int nmr_last_stage_ms(nmr_handle* h, float* sample_ms, float* mlp_ms, float* composite_ms) {
  if (!h) return NMR_ERR_INVALID;
  float acc[3] = {0, 0, 0};
  cudaError_t first = cudaSuccess;
  for (auto& ev : h->events) {
    float ms = 0;
    cudaError_t e = (ev.stage & 0x100) ? cudaErrorNotReady : cudaEventSynchronize(ev.b);   // (an open record has no end event)
    if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, ev.a, ev.b);
    if (e == cudaSuccess) acc[ev.stage & 3] += ms; else if (first == cudaSuccess && !(ev.stage & 0x100)) first = e;
    h->event_pool.push_back(ev);
  }
  h->events.clear();
  if (first != cudaSuccess) { cudaGetLastError(); NMR_FAIL(h, NMR_ERR_CUDA, "nmr_last_stage_ms: %s", cudaGetErrorString(first)); }
  if (sample_ms) *sample_ms = acc[0];
  if (mlp_ms) *mlp_ms = acc[1];
  if (composite_ms) *composite_ms = acc[2];
  return NMR_OK;
}

// ------------------------------------------------------------------------------------------
// the fp32 MLP: NeRF.forward (models/nerf.py:94-144) as a chain of SIMT GEMMs
// ------------------------------------------------------------------------------------------
// X: padded input rows [P,kx].  Returns the buffer that holds the tapped features in *feat.
static int mlp_fp32(nmr_handle* h, int which, const float* X, int64_t P, int stop_layer, float* raw,
                    const float** feat, cudaStream_t s) {
  const Dims& d = h->d;
  const NetF32& n = h->net[which];
  const int skipl = h->cfg.skip_layer + 1;
  float* bufs[4] = {(float*)h->B[0].p, (float*)h->B[1].p, (float*)h->B[2].p, (float*)h->B[3].p};
  const float* in = X; int lda = d.kx;
  const float* tap = nullptr;
  auto pick = [&](const float* a, const float* b) -> float* {
    for (float* c : bufs) if (c != a && c != b) return c;
    return nullptr;
  };
  for (int i = 0; i < kLayers; ++i) {
    float* out = pick(in, tap);
    GemmParams g{};
    if (i == skipl) {
      g.A1 = X; g.lda1 = d.kx; g.W1 = n.Wskip_x; g.K1 = d.xyzp;
      g.A2 = in; g.lda2 = lda; g.W2 = n.W[i]; g.K2 = kHid;
    } else {
      g.A1 = in; g.lda1 = lda; g.W1 = n.W[i]; g.K1 = (i == 0) ? d.xyzp : kHid;
    }
    g.bias = n.b[i]; g.C = out; g.ldc = kHid; g.M = P; g.N = kHid; g.act = 1;
    launch_sgemm(g, s);
    NMR_LAUNCHED(h, "sgemm trunk");
    in = out; lda = kHid;
    if (i == stop_layer) tap = out;
  }
  if (stop_layer < 0) tap = in;   // pt_feat = last trunk activation (models/nerf.py:107, 134-141)
  {  // alpha_linear -> raw[:,3]
    GemmParams g{};
    g.A1 = in; g.lda1 = kHid; g.W1 = n.Walpha; g.K1 = kHid; g.bias = n.balpha;
    g.C = raw + 3; g.ldc = 4; g.M = P; g.N = 1; g.act = 0;
    launch_sgemm(g, s);
    NMR_LAUNCHED(h, "alpha head");
  }
  float* F = pick(in, tap);
  {  // feature_linear (no activation)
    GemmParams g{};
    g.A1 = in; g.lda1 = kHid; g.W1 = n.Wfeat; g.K1 = kHid; g.bias = n.bfeat;
    g.C = F; g.ldc = kHid; g.M = P; g.N = kHid; g.act = 0;
    launch_sgemm(g, s);
    NMR_LAUNCHED(h, "feature_linear");
  }
  float* V = pick(F, tap);
  {  // views_linears.0 on [feature | dirs | app] + ReLU
    GemmParams g{};
    g.A1 = F; g.lda1 = kHid; g.W1 = n.Wview_f; g.K1 = kHid;
    g.A2 = X + d.xyzp; g.lda2 = d.kx; g.W2 = n.Wview_d; g.K2 = d.dap;
    g.bias = n.bview; g.C = V; g.ldc = kHalf; g.M = P; g.N = kHalf; g.act = 1;
    launch_sgemm(g, s);
    NMR_LAUNCHED(h, "views_linear");
  }
  {  // rgb_linear + sigmoid -> raw[:,0:3]
    GemmParams g{};
    g.A1 = V; g.lda1 = kHalf; g.W1 = n.Wrgb; g.K1 = kHalf; g.bias = n.brgb;
    g.C = raw; g.ldc = 4; g.M = P; g.N = 3; g.act = 2;
    launch_sgemm(g, s);
    NMR_LAUNCHED(h, "rgb head");
  }
  *feat = tap;
  return NMR_OK;
}

// ------------------------------------------------------------------------------------------
// render
// ------------------------------------------------------------------------------------------
static int validate_render(nmr_handle* h, const nmr_render_args* a) {
  if (!a) NMR_FAIL(h, NMR_ERR_INVALID, "nmr_render: args is NULL");
  if (a->n < 0) NMR_FAIL(h, NMR_ERR_INVALID, "nmr_render: n < 0");
  if (a->n > 0 && !a->rays) NMR_FAIL(h, NMR_ERR_INVALID, "nmr_render: rays is NULL");
  if (a->ncols < 8) NMR_FAIL(h, NMR_ERR_INVALID, "nmr_render: rays need at least 8 columns (o, d, near, far), got %d", a->ncols);
  if (h->d.mip && a->ncols < 12) NMR_FAIL(h, NMR_ERR_INVALID, "nmr_render: mip rays need 12 columns (radius in column 11), got %d", a->ncols);
  if (!h->net[0].set || !h->net[1].set) NMR_FAIL(h, NMR_ERR_STATE, "nmr_render: weights of both nets must be set first");
  if (h->cfg.precision != NMR_PREC_FP32) {
    const int fmt = h->cfg.precision == NMR_PREC_FP16;
    if (!h->tc[fmt][0].set || !h->tc[fmt][1].set)
      NMR_FAIL(h, NMR_ERR_UNSUPPORTED, "nmr_render: this configuration / these weights are not built for the tcgen05 path in this operand format; use NMR_PREC_FP32");
  }
  if (h->d.app > 0 && h->num_frames == 0) NMR_FAIL(h, NMR_ERR_STATE, "nmr_render: appearance table not set");
  if (a->pt3d && !a->unnorm) NMR_FAIL(h, NMR_ERR_INVALID, "nmr_render: pt3d requested without unnorm");
  return NMR_OK;
}

// per-ray outputs that nmr_render_host streams back chunk by chunk
struct HostSink {
  struct Item { char* host; const char* dev; size_t bytes_per_ray; };
  std::vector<Item> items;
  std::vector<cudaEvent_t> events;
  float* s_fine_host = nullptr;   // s_fine needs the whole call's min / max: copied home after the last chunk
};

static int render_device(nmr_handle* h, const nmr_render_args* a, cudaStream_t s, HostSink* sink = nullptr) {
  int r = validate_render(h, a);
  if (r != NMR_OK) return r;
  if (a->n == 0) return NMR_OK;
  NMR_CUDA(h, cudaSetDevice(h->device));
  const Dims& d = h->d;
  const nmr_config& c = h->cfg;
  const bool bf16 = c.precision != NMR_PREC_FP32;      // the tcgen05 path (bf16 or fp16 operands)
  const int tcfmt = c.precision == NMR_PREC_FP16;
  const int smax = std::max(d.sc, d.sf);
  const bool pfeat = (a->flags & NMR_FLAG_RET_PFEAT) != 0;
  // tcgen05 path: does any pass stream per-sample features to the compositor (unfused tap: 512 B per point of workspace)?
  const bool fusable = h->fuse_composite && !(a->flags & NMR_FLAG_FEAT_MAX) && std::min(d.sc, d.sf) >= kTcMinFuseS &&
                       !a->sigma_noise_coarse && !a->sigma_noise_fine;
  const bool feat_tiles = pfeat && !fusable;
  // fp32 path: ~3.7 KB of workspace per point.  tcgen05 path: 48 B per point fused (raw + per-tile partial composites), so a whole
  // 480x640 view (59 M points, 2.8 GB) goes through in ONE chunk - no per-chunk launch ramps / partial last waves; 0.6 KB per point
  // with the unfused tap.  The host entry (sink) halves the remaining rays chunk by chunk (1/2, 1/4, ... of the call): the big
  // first chunks keep the launch count low and their device->host copies ride under the rest of the render; only the small last
  // chunk's copy (~3-5 % of the outputs) is exposed.
  const int64_t max_points = (bf16 && h->auto_chunk) ? (feat_tiles ? ((int64_t)1 << 22) : ((int64_t)1 << 26)) : c.max_chunk_points;
  const int64_t cap_rays = std::max<int64_t>(1, std::min<int64_t>(a->n, max_points / smax));
  std::vector<int64_t> schedule;                       // rays per chunk
  if (sink && bf16 && h->auto_chunk && !feat_tiles) {
    const int64_t min_chunk = 8192;
    int64_t rem = a->n;
    while (rem > 2 * min_chunk) {
      const int64_t cnk = std::min(cap_rays, (rem / 2 + 1023) / 1024 * 1024);
      schedule.push_back(cnk);
      rem -= cnk;
    }
    while (rem > 0) { const int64_t cnk = std::min(cap_rays, rem); schedule.push_back(cnk); rem -= cnk; }
  } else {
    for (int64_t rem = a->n; rem > 0; rem -= std::min(cap_rays, rem)) schedule.push_back(std::min(cap_rays, rem));
  }
  const int64_t chunk_rays = *std::max_element(schedule.begin(), schedule.end());     // workspace: the largest chunk
  if ((r = ensure_workspace(h, chunk_rays, bf16 ? 0 : chunk_rays * smax)) != NMR_OK) return r;
  if (bf16) {
    std::string terr;
    if (tc_ensure_workspace(d, chunk_rays, &h->tcws.p, &h->tcws.cap, &terr, feat_tiles) != 0) NMR_FAIL(h, NMR_ERR_CUDA, "tc workspace: %s", terr.c_str());
  }
  float* zc = (float*)h->zc.p;
  float* zf = (float*)h->zf.p;
  float* X = (float*)h->X.p;
  float* raw = (float*)h->raw.p;
  const float* app_table = (const float*)h->app.p;

  int64_t r0 = 0;
  for (size_t ci = 0; ci < schedule.size(); r0 += schedule[ci], ++ci) {
    const int64_t nr = schedule[ci];
    const float* rays = a->rays + r0 * a->ncols;
    const int64_t* rid = a->ray_id ? a->ray_id + r0 : nullptr;
    for (int pass = 0; pass < 2; ++pass) {
      const int S = pass == 0 ? d.sc : d.sf;
      const int nz = pass == 0 ? d.zc : d.zf;
      float* z = pass == 0 ? zc : zf;
      const int stop = pass == 0 ? c.stop_layer_coarse : c.stop_layer_fine;
      // ---- sampling + encoding (kernel a) ----
      stage_begin(h, 0, s);
      if (pass == 0) {
        SampleParams sp{};
        sp.rays = rays; sp.n = nr; sp.ncols = a->ncols; sp.ray0 = r0; sp.S = d.sc; sp.mip = d.mip; sp.use_disp = c.use_disp;
        sp.perturb = !d.mip && (a->t_rand != nullptr || (a->flags & NMR_FLAG_PERTURB));     // classic mode: training only
        sp.t_rand = a->t_rand ? a->t_rand + r0 * d.zc : nullptr;
        sp.rng_mode = a->rng_mode; sp.seed = a->seed; sp.z = zc;
        launch_sample_coarse(sp, s);
        NMR_LAUNCHED(h, "sample_coarse");
      }
      const float* feat = nullptr; int feat_bf16 = 0;
      const float* partials = nullptr;
      float* raw_pass = raw;
      const bool want_feat = pfeat && (pass == 1 ? a->feat_fine != nullptr
                                                 : (a->feat_coarse != nullptr && !(a->flags & NMR_FLAG_SKIP_COARSE_FEAT)));
      if (!bf16) {
        EncodeParams ep{};
        ep.rays = rays; ep.n = nr; ep.ncols = a->ncols; ep.S = S; ep.d = d; ep.var_scale = c.mip_var_scale;
        ep.z = z; ep.ray_id = rid; ep.app_table = app_table; ep.num_frames = h->num_frames; ep.X = X;
        launch_encode(ep, s);
        NMR_LAUNCHED(h, "encode");
        stage_end(h, s);
        // ---- MLP (kernel b, fp32 reference-precision variant) ----
        stage_begin(h, 1, s);
        if ((r = mlp_fp32(h, pass, X, nr * S, stop, raw, &feat, s)) != NMR_OK) return r;
        stage_end(h, s);
      } else {
        stage_end(h, s);
        stage_begin(h, 1, s);
        TcRun tr{};
        tr.rays = rays; tr.n = nr; tr.ncols = a->ncols; tr.S = S; tr.z = z; tr.ray_id = rid;
        tr.app_table = app_table; tr.num_frames = h->num_frames; tr.var_scale = c.mip_var_scale;
        tr.stop_layer = stop; tr.want_feat = want_feat; tr.ws = h->tcws.p; tr.ws_rays = chunk_rays; tr.ws_with_feat = feat_tiles;
        // (the in-kernel tile compositing recomputes alpha from the raw density: not with injected sigma noise)
        tr.fuse_composite = h->fuse_composite && !(a->flags & NMR_FLAG_FEAT_MAX) && !(pass == 0 ? a->sigma_noise_coarse : a->sigma_noise_fine);
        std::string terr; int nl = 0;
        if (tc_run(d, h->tc[tcfmt][pass], tr, &raw_pass, (const void**)&feat, &partials, &nl, s, &terr) != 0) NMR_FAIL(h, NMR_ERR_CUDA, "tc_run: %s", terr.c_str());
        NMR_LAUNCHED(h, "tcgen05 mlp", nl);
        feat_bf16 = 1 + tcfmt;
        stage_end(h, s);
      }
      // ---- volume rendering + compositing (+ fused re-sampling after the coarse pass) (kernel c) ----
      stage_begin(h, 2, s);
      CompositeParams cp{};
      cp.raw = raw_pass; cp.z = z; cp.rays = rays; cp.ncols = a->ncols; cp.n = nr; cp.S = S; cp.mip = d.mip;
      cp.flags = a->flags;
      const float* noise = pass == 0 ? a->sigma_noise_coarse : a->sigma_noise_fine;
      cp.sigma_noise = noise ? noise + r0 * S : nullptr;
      cp.feats = want_feat ? feat : nullptr; cp.feats_bf16 = feat_bf16;
      cp.partials = want_feat ? partials : nullptr;
      float* fo = pass == 0 ? a->feat_coarse : a->feat_fine;
      cp.feat_out = (want_feat && fo) ? fo + r0 * kHid : nullptr;
      float* po = pass == 0 ? a->pts_coarse : a->pts_fine;
      cp.pts_out = po ? po + r0 * 3 : nullptr;
      float* ro = pass == 0 ? a->rgb_coarse : a->rgb_fine;
      cp.rgb = ro ? ro + r0 * 3 : nullptr;
      float* dp = pass == 0 ? a->depth_coarse : a->depth_fine;
      cp.depth = dp ? dp + r0 : nullptr;
      float* wo = pass == 0 ? a->weights_coarse : a->weights_fine;
      cp.weights = wo ? wo + r0 * S : nullptr;
      if (pass == 1 && a->pt3d && a->unnorm) {
        cp.pt3d = a->pt3d + r0 * 3; cp.has_unnorm = 1;
        memcpy(cp.unnorm, a->unnorm, 12 * sizeof(float));
      }
      if (pass == 0) {
        cp.z_next = zf; cp.nf = d.nf; cp.resample_padding = c.resample_padding;
        cp.idx_out = a->idx_fine ? a->idx_fine + r0 * (d.mip ? d.sc + 1 : d.nf) : nullptr;
        cp.u_jitter = a->u_jitter ? a->u_jitter + r0 * (d.mip ? d.sc + 1 : d.nf) : nullptr;
        cp.perturb = !d.mip && (a->u_jitter != nullptr || (a->flags & NMR_FLAG_PERTURB));
        cp.rng_mode = a->rng_mode; cp.seed = a->seed; cp.ray0 = r0;
      }
      launch_composite(cp, s);
      NMR_LAUNCHED(h, "composite");
      stage_end(h, s);
      // ---- optional intermediates for the parity tests ----
      float* zo = pass == 0 ? a->z_coarse : a->z_fine;
      if (zo) NMR_CUDA(h, cudaMemcpyAsync(zo + r0 * nz, z, (size_t)nr * nz * sizeof(float), cudaMemcpyDeviceToDevice, s));
      if (pass == 1 && a->s_fine)   // z_fine of every chunk first; mapped to s once the whole call's min / max are known
        NMR_CUDA(h, cudaMemcpyAsync(a->s_fine + r0 * nz, z, (size_t)nr * nz * sizeof(float), cudaMemcpyDeviceToDevice, s));
      float* rawo = pass == 0 ? a->raw_coarse : a->raw_fine;
      if (rawo) NMR_CUDA(h, cudaMemcpyAsync(rawo + r0 * S * 4, raw_pass, (size_t)nr * S * 4 * sizeof(float), cudaMemcpyDeviceToDevice, s));
    }
    if (sink) {   // this chunk's rows go home while the next chunk renders
      cudaEvent_t ev;
      NMR_CUDA(h, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
      sink->events.push_back(ev);
      NMR_CUDA(h, cudaEventRecord(ev, s));
      NMR_CUDA(h, cudaStreamWaitEvent(h->copy_stream, ev, 0));
      for (auto& it : sink->items)
        NMR_CUDA(h, cudaMemcpyAsync(it.host + (size_t)r0 * it.bytes_per_ray, it.dev + (size_t)r0 * it.bytes_per_ray,
                                    (size_t)nr * it.bytes_per_ray, cudaMemcpyDeviceToHost, h->copy_stream));
    }
  }
  if (a->s_fine) {   // preds["s_fine"] = t_to_s(z_fine, z_fine.min(), z_fine.max()) (renderer.py:283-285)
    launch_t_to_s(a->s_fine, a->n * d.zf, (uint32_t*)((char*)h->misc.p + 64), s);
    NMR_LAUNCHED(h, "t_to_s", 3);
    if (sink && sink->s_fine_host) {
      NMR_CUDA(h, cudaMemcpyAsync(sink->s_fine_host, a->s_fine, (size_t)a->n * d.zf * sizeof(float), cudaMemcpyDeviceToHost, s));
    }
  }
  return NMR_OK;
}

extern "C" int nmr_render(nmr_handle* h, const nmr_render_args* args, void* stream) {
  if (!h) return NMR_ERR_INVALID;
  return render_device(h, args, (cudaStream_t)stream);
}

static void invert3x3(const float* K, float* out) {
  double a = K[0], b = K[1], c = K[2], d = K[3], e = K[4], f = K[5], g = K[6], hh = K[7], i = K[8];
  double A = e * i - f * hh, B = -(d * i - f * g), C = d * hh - e * g;
  double det = a * A + b * B + c * C;
  double inv[9] = {A, -(b * i - c * hh), b * f - c * e, B, a * i - c * g, -(a * f - c * d), C, -(a * hh - b * g), a * e - b * d};
  for (int k = 0; k < 9; ++k) out[k] = (float)(inv[k] / det);
}

extern "C" int nmr_raygen(nmr_handle* h, const float* K, const float* c2w, int32_t H, int32_t W, int32_t ds,
                          float* rays_dev, int32_t* far_fallback, void* stream) {
  if (!h) return NMR_ERR_INVALID;
  if (!K || !c2w || !rays_dev) NMR_FAIL(h, NMR_ERR_INVALID, "nmr_raygen: null argument");
  if (H < 3 || W < 1 || ds < 1) NMR_FAIL(h, NMR_ERR_INVALID, "nmr_raygen: need H >= 3, W >= 1, ds >= 1");
  NMR_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t s = (cudaStream_t)stream;
  NMR_CUDA(h, h->misc.reserve(256));
  RayGenParams p{};
  invert3x3(K, p.kinv);
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) p.rot[3 * i + j] = c2w[4 * i + j];
    p.org[i] = c2w[4 * i + 3];
  }
  p.H = H; p.W = W; p.ds = ds; p.mip = h->d.mip;
  p.h = (H - ds / 2 + ds - 1) / ds; p.w = (W - ds / 2 + ds - 1) / ds;   // len(range(ds//2, H, ds))
  p.rays = rays_dev; p.miss_flag = (int*)h->misc.p;
  NMR_CUDA(h, cudaMemsetAsync(p.miss_flag, 0, sizeof(int), s));
  launch_raygen(p, s);
  NMR_LAUNCHED(h, "raygen");
  const float on = p.org[0] * p.org[0] + p.org[1] * p.org[1] + p.org[2] * p.org[2];
  if (ds > 1 && !(1.0f - on >= 0.0f)) {          // camera outside the sphere: pixels the subsampling skips can miss too
    launch_raygen_miss(p, s);
    NMR_LAUNCHED(h, "raygen_miss");
  }
  launch_far_fallback(rays_dev, (int64_t)p.h * p.w, h->d.mip ? 12 : 11, p.miss_flag, s);
  NMR_LAUNCHED(h, "far_fallback");
  if (far_fallback) {
    NMR_CUDA(h, cudaMemcpyAsync(far_fallback, p.miss_flag, sizeof(int), cudaMemcpyDeviceToHost, s));
    NMR_CUDA(h, cudaStreamSynchronize(s));
  }
  return NMR_OK;
}

// ---- the reference's ray helpers as separate device ops (render_utils.py:23-41, 81-104) ----
extern "C" int nmr_ray_dirs(nmr_handle* h, const float* K, int32_t H, int32_t W, int32_t flipped_yz, float* dirs_dev, void* stream) {
  if (!h) return NMR_ERR_INVALID;
  if (!K || !dirs_dev || H < 1 || W < 1) NMR_FAIL(h, NMR_ERR_INVALID, "nmr_ray_dirs: bad argument");
  NMR_CUDA(h, cudaSetDevice(h->device));
  RayOpsParams p{};
  invert3x3(K, p.kinv);
  p.H = H; p.W = W; p.flip = flipped_yz; p.out = dirs_dev;
  launch_ray_op(0, p, (cudaStream_t)stream);
  NMR_LAUNCHED(h, "ray_dirs");
  return NMR_OK;
}

extern "C" int nmr_rays_c2w(nmr_handle* h, const float* dirs_dev, int64_t n, const float* c2w, float* rays_o, float* rays_d,
                            float* viewdirs, void* stream) {
  if (!h) return NMR_ERR_INVALID;
  if (n < 0 || !c2w || (n > 0 && (!dirs_dev || !rays_o || !rays_d || !viewdirs))) NMR_FAIL(h, NMR_ERR_INVALID, "nmr_rays_c2w: bad argument");
  NMR_CUDA(h, cudaSetDevice(h->device));
  RayOpsParams p{};
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) p.rot[3 * i + j] = c2w[4 * i + j];
    p.org[i] = c2w[4 * i + 3];
  }
  p.n = n; p.dirs = dirs_dev; p.o = rays_o; p.d = rays_d; p.v = viewdirs;
  launch_ray_op(1, p, (cudaStream_t)stream);
  NMR_LAUNCHED(h, "rays_c2w");
  return NMR_OK;
}

extern "C" int nmr_prepare_rays(nmr_handle* h, const float* rays_o, const float* rays_d, const float* viewdirs, const float* near_dev,
                                float near_s, const float* far_dev, float far_s, int32_t H, int32_t W, int32_t comp_radii,
                                float* rays_out, void* stream) {
  if (!h) return NMR_ERR_INVALID;
  if (H < 1 || W < 1 || !rays_o || !rays_d || !rays_out) NMR_FAIL(h, NMR_ERR_INVALID, "nmr_prepare_rays: bad argument");
  if (comp_radii && H < 3) NMR_FAIL(h, NMR_ERR_INVALID, "nmr_prepare_rays: radii need at least 3 image rows");
  NMR_CUDA(h, cudaSetDevice(h->device));
  RayOpsParams p{};
  p.H = H; p.W = W; p.n = (int64_t)H * W; p.comp_radii = comp_radii;
  p.o = const_cast<float*>(rays_o); p.d = const_cast<float*>(rays_d); p.v = const_cast<float*>(viewdirs);
  p.near = near_dev; p.near_s = near_s; p.far = far_dev; p.far_s = far_s;
  p.ncols = 8 + (viewdirs ? 3 : 0) + (comp_radii ? 1 : 0);
  p.out = rays_out;
  launch_ray_op(2, p, (cudaStream_t)stream);
  NMR_LAUNCHED(h, "prepare_rays");
  return NMR_OK;
}

extern "C" int nmr_render_view(nmr_handle* h, const float* K, const float* c2w, int32_t H, int32_t W, int32_t ds,
                               const nmr_render_args* args, void* stream) {
  if (!h) return NMR_ERR_INVALID;
  if (!args) NMR_FAIL(h, NMR_ERR_INVALID, "nmr_render_view: args is NULL");
  if (ds < 1) NMR_FAIL(h, NMR_ERR_INVALID, "nmr_render_view: ds < 1");
  const int hh = (H - ds / 2 + ds - 1) / ds, ww = (W - ds / 2 + ds - 1) / ds;
  const int ncols = h->d.mip ? 12 : 11;
  NMR_CUDA(h, cudaSetDevice(h->device));
  NMR_CUDA(h, h->rays_view.reserve((size_t)hh * ww * ncols * sizeof(float)));
  int r = nmr_raygen(h, K, c2w, H, W, ds, (float*)h->rays_view.p, nullptr, stream);
  if (r != NMR_OK) return r;
  nmr_render_args a = *args;
  a.rays = (const float*)h->rays_view.p; a.n = (int64_t)hh * ww; a.ncols = ncols;
  return render_device(h, &a, (cudaStream_t)stream);
}

// host-buffer entry: H2D rays (+ optional draws / ids), render, D2H outputs, synchronise.
extern "C" int nmr_render_host(nmr_handle* h, const nmr_render_args* args, void* stream) {
  if (!h) return NMR_ERR_INVALID;
  if (!args) NMR_FAIL(h, NMR_ERR_INVALID, "nmr_render_host: args is NULL");
  if (args->n < 0 || (args->n > 0 && !args->rays)) NMR_FAIL(h, NMR_ERR_INVALID, "nmr_render_host: bad rays");
  NMR_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t s = (cudaStream_t)stream;
  const Dims& d = h->d;
  const int64_t n = args->n;
  nmr_render_args a = *args;
  int slot = 0;
  if (!h->copy_stream) NMR_CUDA(h, cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
  HostSink sink;
  cudaError_t stage_err = cudaSuccess;          // first staging failure (allocation or copy) of any buffer
  const char* stage_what = "";
  auto in = [&](const void* host, size_t bytes, const char* what) -> void* {
    DevBuf& b = h->stage[slot++];
    if (!host || bytes == 0) return nullptr;
    cudaError_t e = b.reserve(bytes);
    if (e == cudaSuccess) e = cudaMemcpyAsync(b.p, host, bytes, cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess) { if (stage_err == cudaSuccess) { stage_err = e; stage_what = what; } return nullptr; }
    return b.p;
  };
  auto out = [&](void* host, size_t bytes, const char* what) -> void* {
    DevBuf& b = h->stage[slot++];
    if (!host || bytes == 0 || n == 0) return nullptr;
    cudaError_t e = b.reserve(bytes);
    if (e != cudaSuccess) { if (stage_err == cudaSuccess) { stage_err = e; stage_what = what; } return nullptr; }
    sink.items.push_back({(char*)host, (const char*)b.p, bytes / (size_t)n});
    return b.p;
  };
  const size_t f = sizeof(float);
  a.rays = (const float*)in(args->rays, (size_t)n * args->ncols * f, "rays");
  a.ray_id = (const int64_t*)in(args->ray_id, (size_t)n * sizeof(int64_t), "ray_id");
  a.t_rand = (const float*)in(args->t_rand, (size_t)n * d.zc * f, "t_rand");
  a.u_jitter = (const float*)in(args->u_jitter, (size_t)n * (d.mip ? d.sc + 1 : d.nf) * f, "u_jitter");
  a.sigma_noise_coarse = (const float*)in(args->sigma_noise_coarse, (size_t)n * d.sc * f, "sigma_noise_coarse");
  a.sigma_noise_fine = (const float*)in(args->sigma_noise_fine, (size_t)n * d.sf * f, "sigma_noise_fine");
  a.feat_coarse = (float*)out(args->feat_coarse, (size_t)n * kHid * f, "feat_coarse");
  a.feat_fine = (float*)out(args->feat_fine, (size_t)n * kHid * f, "feat_fine");
  a.pts_coarse = (float*)out(args->pts_coarse, (size_t)n * 3 * f, "pts_coarse");
  a.pts_fine = (float*)out(args->pts_fine, (size_t)n * 3 * f, "pts_fine");
  a.rgb_coarse = (float*)out(args->rgb_coarse, (size_t)n * 3 * f, "rgb_coarse");
  a.rgb_fine = (float*)out(args->rgb_fine, (size_t)n * 3 * f, "rgb_fine");
  a.depth_coarse = (float*)out(args->depth_coarse, (size_t)n * f, "depth_coarse");
  a.depth_fine = (float*)out(args->depth_fine, (size_t)n * f, "depth_fine");
  a.pt3d = (float*)out(args->pt3d, (size_t)n * 3 * f, "pt3d");
  a.z_coarse = (float*)out(args->z_coarse, (size_t)n * d.zc * f, "z_coarse");
  a.z_fine = (float*)out(args->z_fine, (size_t)n * d.zf * f, "z_fine");
  a.weights_coarse = (float*)out(args->weights_coarse, (size_t)n * d.sc * f, "weights_coarse");
  a.weights_fine = (float*)out(args->weights_fine, (size_t)n * d.sf * f, "weights_fine");
  a.raw_coarse = (float*)out(args->raw_coarse, (size_t)n * d.sc * 4 * f, "raw_coarse");
  a.raw_fine = (float*)out(args->raw_fine, (size_t)n * d.sf * 4 * f, "raw_fine");
  a.idx_fine = (int32_t*)out(args->idx_fine, (size_t)n * (d.mip ? d.sc + 1 : d.nf) * sizeof(int32_t), "idx_fine");
  a.s_fine = (float*)out(args->s_fine, (size_t)n * d.zf * f, "s_fine");
  if (a.s_fine) { sink.items.pop_back(); sink.s_fine_host = args->s_fine; }      // not per chunk: see HostSink
  if (stage_err != cudaSuccess) {
    cudaGetLastError();
    NMR_FAIL(h, NMR_ERR_CUDA, "nmr_render_host: staging %s failed: %s", stage_what, cudaGetErrorString(stage_err));
  }
  int r = render_device(h, &a, s, &sink);
  cudaError_t e1 = cudaStreamSynchronize(s), e2 = cudaStreamSynchronize(h->copy_stream);
  for (auto ev : sink.events) cudaEventDestroy(ev);
  if (r != NMR_OK) return r;
  NMR_CUDA(h, e1);
  NMR_CUDA(h, e2);
  return NMR_OK;
}

// ------------------------------------------------------------------------------------------
// stage entry points
// ------------------------------------------------------------------------------------------
extern "C" int nmr_stage_sample_coarse(nmr_handle* h, const float* rays, int64_t n, int32_t ncols, const float* t_rand,
                                       float* z, float* pts, float* var, void* stream) {
  if (!h) return NMR_ERR_INVALID;
  if (n < 0 || (n > 0 && (!rays || !z))) NMR_FAIL(h, NMR_ERR_INVALID, "nmr_stage_sample_coarse: null argument");
  if (ncols < 8 || (h->d.mip && ncols < 12)) NMR_FAIL(h, NMR_ERR_INVALID, "nmr_stage_sample_coarse: ncols=%d", ncols);
  if (n == 0) return NMR_OK;
  NMR_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t s = (cudaStream_t)stream;
  const Dims& d = h->d;
  SampleParams sp{};
  sp.rays = rays; sp.n = n; sp.ncols = ncols; sp.S = d.sc; sp.mip = d.mip; sp.use_disp = h->cfg.use_disp;
  sp.t_rand = t_rand; sp.rng_mode = NMR_RNG_NONE; sp.z = z;
  launch_sample_coarse(sp, s);
  NMR_LAUNCHED(h, "sample_coarse");
  if (pts || var) {
    PointsParams pp{};
    pp.rays = rays; pp.n = n; pp.ncols = ncols; pp.S = d.sc; pp.mip = d.mip; pp.var_scale = h->cfg.mip_var_scale;
    pp.z = z; pp.pts = pts; pp.var = var;
    launch_points(pp, s);
    NMR_LAUNCHED(h, "points");
  }
  return NMR_OK;
}

extern "C" int nmr_stage_resample(nmr_handle* h, const float* rays, int64_t n, int32_t ncols, const float* z_in,
                                  const float* weights, const float* u_jitter, float* z_out, int32_t* idx,
                                  float* pts, float* var, void* stream) {
  if (!h) return NMR_ERR_INVALID;
  if (n < 0 || (n > 0 && (!rays || !z_in || !weights || !z_out))) NMR_FAIL(h, NMR_ERR_INVALID, "nmr_stage_resample: null argument");
  if (ncols < 8 || (h->d.mip && ncols < 12)) NMR_FAIL(h, NMR_ERR_INVALID, "nmr_stage_resample: ncols=%d", ncols);
  if (n == 0) return NMR_OK;
  NMR_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t s = (cudaStream_t)stream;
  const Dims& d = h->d;
  CompositeParams cp{};
  cp.z = z_in; cp.rays = rays; cp.ncols = ncols; cp.n = n; cp.S = d.sc; cp.mip = d.mip;
  cp.weights_in = weights; cp.z_next = z_out; cp.idx_out = idx; cp.nf = d.nf;
  cp.u_jitter = u_jitter; cp.rng_mode = NMR_RNG_NONE; cp.resample_padding = h->cfg.resample_padding;
  launch_composite(cp, s);
  NMR_LAUNCHED(h, "resample");
  if (pts || var) {
    PointsParams pp{};
    pp.rays = rays; pp.n = n; pp.ncols = ncols; pp.S = d.sf; pp.mip = d.mip; pp.var_scale = h->cfg.mip_var_scale;
    pp.z = z_out; pp.pts = pts; pp.var = var;
    launch_points(pp, s);
    NMR_LAUNCHED(h, "points");
  }
  return NMR_OK;
}

extern "C" int nmr_stage_encode(nmr_handle* h, int kind, const float* x, const float* var, int64_t P, float* out,
                                void* stream) {
  if (!h) return NMR_ERR_INVALID;
  if (P < 0 || (P > 0 && (!x || !out)) || (kind != 0 && kind != 1)) NMR_FAIL(h, NMR_ERR_INVALID, "nmr_stage_encode: bad argument");
  if (P == 0) return NMR_OK;
  NMR_CUDA(h, cudaSetDevice(h->device));
  launch_encode_stage(kind, h->d, x, var, P, out, (cudaStream_t)stream);
  NMR_LAUNCHED(h, "encode_stage");
  return NMR_OK;
}

static int stage_mlp_impl(nmr_handle* h, int which, const float* x, int64_t P, int32_t stop_layer, float* raw, float* feat,
                          uint32_t* masks, void* stream);

extern "C" int nmr_stage_mlp(nmr_handle* h, int which, const float* x, int64_t P, int32_t stop_layer, float* raw,
                             float* feat, void* stream) {
  return stage_mlp_impl(h, which, x, P, stop_layer, raw, feat, nullptr, stream);
}

extern "C" int nmr_stage_mlp_masks(nmr_handle* h, int which, const float* x, int64_t P, int32_t stop_layer, float* raw,
                                   float* feat, uint32_t* masks, void* stream) {
  if (!h) return NMR_ERR_INVALID;
  if (!masks) NMR_FAIL(h, NMR_ERR_INVALID, "nmr_stage_mlp_masks: masks is NULL");
  if (h->cfg.precision == NMR_PREC_FP32) NMR_FAIL(h, NMR_ERR_UNSUPPORTED, "nmr_stage_mlp_masks: the fp32 path keeps no masks (nmr_stage_mlp_backward recomputes the activations)");
  return stage_mlp_impl(h, which, x, P, stop_layer, raw, feat, masks, stream);
}

static int stage_mlp_impl(nmr_handle* h, int which, const float* x, int64_t P, int32_t stop_layer, float* raw, float* feat,
                          uint32_t* masks, void* stream) {
  if (!h) return NMR_ERR_INVALID;
  if (which != 0 && which != 1) NMR_FAIL(h, NMR_ERR_INVALID, "nmr_stage_mlp: which=%d", which);
  if (P < 0 || (P > 0 && (!x || !raw))) NMR_FAIL(h, NMR_ERR_INVALID, "nmr_stage_mlp: null argument");
  if (stop_layer >= kLayers) NMR_FAIL(h, NMR_ERR_INVALID, "nmr_stage_mlp: stop_layer=%d", stop_layer);
  if (!h->net[which].set) NMR_FAIL(h, NMR_ERR_STATE, "nmr_stage_mlp: weights not set");
  if (P == 0) return NMR_OK;
  NMR_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t s = (cudaStream_t)stream;
  const Dims& d = h->d;
  const int64_t chunk = std::min<int64_t>(P, std::max<int64_t>(128, h->cfg.max_chunk_points / 128 * 128));   // whole tiles per chunk
  int r;
  if ((r = ensure_workspace(h, 1, chunk)) != NMR_OK) return r;
  for (int64_t p0 = 0; p0 < P; p0 += chunk) {
    const int64_t np = std::min<int64_t>(chunk, P - p0);
    launch_repack_input(d, x + p0 * d.in_dim, np, (float*)h->X.p, s);
    NMR_LAUNCHED(h, "repack_input");
    const float* tap = nullptr;
    if (h->cfg.precision != NMR_PREC_FP32) {
      const TcNet& tcn = h->tc[h->cfg.precision == NMR_PREC_FP16][which];
      if (!tcn.set) NMR_FAIL(h, NMR_ERR_UNSUPPORTED, "nmr_stage_mlp: this configuration / these weights are not built for the tcgen05 path");
      std::string terr; int nl = 0;
      if (tc_run_rows(d, tcn, (const float*)h->X.p, np, stop_layer, raw + p0 * 4, feat ? feat + p0 * kHid : nullptr,
                      &h->tcws.p, &h->tcws.cap, &nl, s, &terr, 0, nullptr, masks ? masks + (p0 / 128) * kTcMaskWordsPerTile : nullptr) != 0)
        NMR_FAIL(h, NMR_ERR_CUDA, "tc_run_rows: %s", terr.c_str());
      NMR_LAUNCHED(h, "tcgen05 mlp rows", nl);
    } else {
      if ((r = mlp_fp32(h, which, (const float*)h->X.p, np, stop_layer, (float*)h->raw.p, &tap, s)) != NMR_OK) return r;
      NMR_CUDA(h, cudaMemcpyAsync(raw + p0 * 4, h->raw.p, (size_t)np * 4 * sizeof(float), cudaMemcpyDeviceToDevice, s));
      if (feat) NMR_CUDA(h, cudaMemcpyAsync(feat + p0 * kHid, tap, (size_t)np * kHid * sizeof(float), cudaMemcpyDeviceToDevice, s));
    }
  }
  return NMR_OK;
}


// ------------------------------------------------------------------------------------------
// peer-memory exchange
// ------------------------------------------------------------------------------------------
extern "C" int nmr_peer_alloc(nmr_handle* h, size_t bytes, void** dev_ptr, nmr_peer_handle* out) {
  if (!h) return NMR_ERR_INVALID;
  if (!dev_ptr || !out || bytes == 0) NMR_FAIL(h, NMR_ERR_INVALID, "nmr_peer_alloc: bad argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == sizeof(nmr_peer_handle), "IPC handle size");
  NMR_CUDA(h, cudaSetDevice(h->device));
  void* p = nullptr;
  NMR_CUDA(h, cudaMalloc(&p, bytes));
  cudaError_t e = cudaMemset(p, 0, bytes);
  cudaIpcMemHandle_t ih;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&ih, p);
  if (e != cudaSuccess) { cudaFree(p); cudaGetLastError(); NMR_FAIL(h, NMR_ERR_CUDA, "nmr_peer_alloc: %s", cudaGetErrorString(e)); }
  memcpy(out->bytes, &ih, sizeof(ih));
  *dev_ptr = p;
  return NMR_OK;
}
extern "C" int nmr_peer_free(nmr_handle* h, void* dev_ptr) {
  if (!h) return NMR_ERR_INVALID;
  NMR_CUDA(h, cudaSetDevice(h->device));
  NMR_CUDA(h, cudaFree(dev_ptr));
  return NMR_OK;
}
extern "C" int nmr_peer_open(nmr_handle* h, const nmr_peer_handle* in, void** peer_ptr) {
  if (!h) return NMR_ERR_INVALID;
  if (!in || !peer_ptr) NMR_FAIL(h, NMR_ERR_INVALID, "nmr_peer_open: null argument");
  NMR_CUDA(h, cudaSetDevice(h->device));
  cudaIpcMemHandle_t ih;
  memcpy(&ih, in->bytes, sizeof(ih));
  NMR_CUDA(h, cudaIpcOpenMemHandle(peer_ptr, ih, cudaIpcMemLazyEnablePeerAccess));
  return NMR_OK;
}
extern "C" int nmr_peer_close(nmr_handle* h, void* peer_ptr) {
  if (!h) return NMR_ERR_INVALID;
  NMR_CUDA(h, cudaSetDevice(h->device));
  NMR_CUDA(h, cudaIpcCloseMemHandle(peer_ptr));
  return NMR_OK;
}
extern "C" int nmr_peer_signal(nmr_handle* h, uint32_t* flag, uint32_t value, void* stream) {
  if (!h) return NMR_ERR_INVALID;
  if (!flag) NMR_FAIL(h, NMR_ERR_INVALID, "nmr_peer_signal: null flag");
  NMR_CUDA(h, cudaSetDevice(h->device));
  launch_peer_signal(flag, value, (cudaStream_t)stream);
  NMR_LAUNCHED(h, "peer_signal");
  return NMR_OK;
}
extern "C" int nmr_peer_wait(nmr_handle* h, const uint32_t* flags, int32_t count, uint32_t value, void* stream) {
  if (!h) return NMR_ERR_INVALID;
  if (!flags || count < 1 || count > 64) NMR_FAIL(h, NMR_ERR_INVALID, "nmr_peer_wait: need 1..64 flags");
  NMR_CUDA(h, cudaSetDevice(h->device));
  if (!h->peer_status) {
    NMR_CUDA(h, cudaHostAlloc((void**)&h->peer_status, sizeof(int), cudaHostAllocMapped));
    *h->peer_status = 0;
    NMR_CUDA(h, cudaHostGetDevicePointer((void**)&h->peer_status_dev, h->peer_status, 0));
  }
  launch_peer_wait(flags, count, value, h->peer_status_dev, (cudaStream_t)stream);
  NMR_LAUNCHED(h, "peer_wait");
  return NMR_OK;
}
extern "C" int nmr_peer_wait_status(nmr_handle* h) {
  if (!h || !h->peer_status) return 0;
  const int v = *(volatile int*)h->peer_status;
  *(volatile int*)h->peer_status = 0;
  return v;
}

// ------------------------------------------------------------------------------------------
// differentiable path: input gradients of the three stages the iNeRF refinement
// (nerfmatch_evaluator.py:288-500) differentiates through
// ------------------------------------------------------------------------------------------
extern "C" int nmr_stage_encode_backward(nmr_handle* h, int kind, const float* x, const float* var, int64_t P,
                                         const float* grad_out, float* grad_x, void* stream) {
  if (!h) return NMR_ERR_INVALID;
  if (P < 0 || (P > 0 && (!x || !grad_out || !grad_x)) || (kind != 0 && kind != 1)) NMR_FAIL(h, NMR_ERR_INVALID, "nmr_stage_encode_backward: bad argument");
  if (P == 0) return NMR_OK;
  NMR_CUDA(h, cudaSetDevice(h->device));
  launch_encode_backward(kind, h->d, x, var, grad_out, P, grad_x, (cudaStream_t)stream);
  NMR_LAUNCHED(h, "encode_backward");
  return NMR_OK;
}

extern "C" int nmr_stage_composite_backward(nmr_handle* h, const float* raw, const float* z, const float* rays_d, int64_t n, int32_t S,
                                            uint32_t flags, const float* grad_rgb, const float* grad_depth, const float* grad_acc,
                                            const float* grad_weights, float* grad_raw, float* grad_rays_d, void* stream) {
  if (!h) return NMR_ERR_INVALID;
  if (n < 0 || (n > 0 && (!raw || !z || !rays_d || !grad_raw))) NMR_FAIL(h, NMR_ERR_INVALID, "nmr_stage_composite_backward: null argument");
  if (S < 2 || S > kMaxSamples) NMR_FAIL(h, NMR_ERR_INVALID, "nmr_stage_composite_backward: S=%d (2..%d)", S, kMaxSamples);
  if (n == 0) return NMR_OK;
  NMR_CUDA(h, cudaSetDevice(h->device));
  CompositeBwdParams p{};
  p.raw = raw; p.z = z; p.rays_d = rays_d; p.n = n; p.S = S; p.mip = h->d.mip; p.flags = flags;
  p.grad_rgb = grad_rgb; p.grad_depth = grad_depth; p.grad_acc = grad_acc; p.grad_weights = grad_weights;
  p.grad_raw = grad_raw; p.grad_rays_d = grad_rays_d;
  launch_composite_backward(p, (cudaStream_t)stream);
  NMR_LAUNCHED(h, "composite_backward");
  return NMR_OK;
}

// NeRF.forward backwards (models/nerf.py:94-144): d raw [P,4] (+ d feat [P,256] at the tapped layer) -> d x [P,in_dim].
// fp32 gradients.  The ReLU masks come either from the tcgen05 forward kernel (nmr_stage_mlp_masks: one bit per activation, no
// re-forward) or from re-evaluating each chunk of points forward with every activation kept; the dgrad chain
// g_in = (g_out * W) [* mask of the layer below] then runs on the GEMM kernels with transposed panels (tensor cores:
// nmr_gemm_tc.cu with pre-split panels; NMR_GEMM_TC=0: fp32 SIMT).
extern "C" int nmr_stage_mlp_backward(nmr_handle* h, int which, const float* x, int64_t P, int32_t stop_layer, const float* grad_raw,
                                      const float* grad_feat, const float* raw_fwd, const uint32_t* masks, float* grad_x, void* stream) {
  if (!h) return NMR_ERR_INVALID;
  if (which != 0 && which != 1) NMR_FAIL(h, NMR_ERR_INVALID, "nmr_stage_mlp_backward: which=%d", which);
  if (P < 0 || (P > 0 && (!grad_raw || !grad_x))) NMR_FAIL(h, NMR_ERR_INVALID, "nmr_stage_mlp_backward: null argument");
  const bool have_masks = masks != nullptr && raw_fwd != nullptr;      // from nmr_stage_mlp_masks: no re-forward needed
  if (P > 0 && !have_masks && !x) NMR_FAIL(h, NMR_ERR_INVALID, "nmr_stage_mlp_backward: x is needed to recompute the activations");
  if (stop_layer >= kLayers) NMR_FAIL(h, NMR_ERR_INVALID, "nmr_stage_mlp_backward: stop_layer=%d", stop_layer);
  if (!h->net[which].set) NMR_FAIL(h, NMR_ERR_STATE, "nmr_stage_mlp_backward: weights not set");
  if (P == 0) return NMR_OK;
  NMR_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t s = (cudaStream_t)stream;
  const Dims& d = h->d;
  const NetF32& n = h->net[which];
  const int skipl = h->cfg.skip_layer + 1;
  const int tap = stop_layer < 0 ? kLayers - 1 : stop_layer;
  // points per pass: a whole number of rounds of 128-row GEMM tiles over the SMs (the persistent GEMM deals tiles round-robin)
  int sms = 0;
  NMR_CUDA(h, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device));
  const int64_t C = std::min<int64_t>(P, (int64_t)std::max(sms, 1) * 128 * 4);
  // workspace layout (floats per point): X kx | H 8 x 256 | F 256 | V 128 | raw 4 | Gv 128 | Ga 256 | Gb 256 | Gxs xyzp | Gx0 xyzp | Gd dap
  const size_t per_pt = (size_t)d.kx + 8 * kHid + kHid + kHalf + 4 + kHalf + 2 * kHid + 2 * d.xyzp + d.dap;
  NMR_CUDA(h, h->bwd.reserve(per_pt * C * sizeof(float)));
  float* wsp = (float*)h->bwd.p;
  auto take = [&](size_t per) { float* q = wsp; wsp += per * C; return q; };
  float* X = take(d.kx);
  float* H[kLayers];
  for (int i = 0; i < kLayers; ++i) H[i] = take(kHid);
  float* F = take(kHid); float* V = take(kHalf); float* rawc = take(4);
  float* Gv = take(kHalf); float* Ga = take(kHid); float* Gb = take(kHid);
  float* Gxs = take(d.xyzp); float* Gx0 = take(d.xyzp); float* Gd = take(d.dap);
  auto gemm = [&](const GemmParams& g, const char* what) -> int {
    if (h->tc_gemm && gemm_tc_supported(g)) {
      if (launch_gemm_tc(g, s) != 0) NMR_FAIL(h, NMR_ERR_CUDA, "launch_gemm_tc setup failed (%s)", what);
    } else {
      launch_sgemm(g, s);
    }
    return check_launch(h, what);
  };
  int r;
  for (int64_t p0 = 0; p0 < P; p0 += C) {
    const int64_t np = std::min<int64_t>(C, P - p0);
    const float* gr = grad_raw + p0 * 4;
    const float* gf = grad_feat ? grad_feat + p0 * kHid : nullptr;
    auto set_mask = [&](GemmParams& g, int layer) {                    // ReLU mask of trunk layer `layer`'s output
      if (have_masks) { g.mask_bits = masks; g.mask_row0 = p0; g.mask_step = layer; } else { g.mask = H[layer]; g.ldm = kHid; }
    };
    if (!have_masks) {
    // ---- forward, every activation kept ----
    launch_repack_input(d, x + p0 * d.in_dim, np, X, s);
    NMR_LAUNCHED(h, "repack_input");
    const float* in = X; int lda = d.kx;
    for (int i = 0; i < kLayers; ++i) {
      GemmParams g{};
      if (i == skipl) {
        g.A1 = X; g.lda1 = d.kx; g.W1 = n.Wskip_x; g.K1 = d.xyzp;
        g.A2 = in; g.lda2 = lda; g.W2 = n.W[i]; g.K2 = kHid;
      } else {
        g.A1 = in; g.lda1 = lda; g.W1 = n.W[i]; g.K1 = (i == 0) ? d.xyzp : kHid;
      }
      g.bias = n.b[i]; g.C = H[i]; g.ldc = kHid; g.M = np; g.N = kHid; g.act = 1;
      if ((r = gemm(g, "bwd: trunk forward")) != NMR_OK) return r;
      in = H[i]; lda = kHid;
    }
    {
      GemmParams g{};
      g.A1 = H[7]; g.lda1 = kHid; g.W1 = n.Wfeat; g.K1 = kHid; g.bias = n.bfeat; g.C = F; g.ldc = kHid; g.M = np; g.N = kHid; g.act = 0;
      if ((r = gemm(g, "bwd: feature_linear forward")) != NMR_OK) return r;
      GemmParams v{};
      v.A1 = F; v.lda1 = kHid; v.W1 = n.Wview_f; v.K1 = kHid; v.A2 = X + d.xyzp; v.lda2 = d.kx; v.W2 = n.Wview_d; v.K2 = d.dap;
      v.bias = n.bview; v.C = V; v.ldc = kHalf; v.M = np; v.N = kHalf; v.act = 1;
      if ((r = gemm(v, "bwd: views forward")) != NMR_OK) return r;
      GemmParams c{};
      c.A1 = V; c.lda1 = kHalf; c.W1 = n.Wrgb; c.K1 = kHalf; c.bias = n.brgb; c.C = rawc; c.ldc = 4; c.M = np; c.N = 3; c.act = 2;
      if ((r = gemm(c, "bwd: rgb forward")) != NMR_OK) return r;
    }
    }
    // ---- backward ----
    launch_rgb_head_backward(gr, have_masks ? raw_fwd + p0 * 4 : rawc, have_masks ? nullptr : V, masks, p0, n.Wrgb, np, Gv, s);
    NMR_LAUNCHED(h, "rgb_head_backward");
    {
      GemmParams g{};   // d feature = d views_pre * Wview[:, :256]
      g.A1 = Gv; g.lda1 = kHalf; g.W1 = n.Wtview_f; g.W1p = n.Ptview_f; g.K1 = kHalf; g.C = Ga; g.ldc = kHid; g.M = np; g.N = kHid;
      if ((r = gemm(g, "bwd: d feature")) != NMR_OK) return r;
      GemmParams q{};   // d [dirs | app] = d views_pre * Wview[:, 256:]
      q.A1 = Gv; q.lda1 = kHalf; q.W1 = n.Wtview_d; q.W1p = n.Ptview_d; q.K1 = kHalf; q.C = Gd; q.ldc = d.dap; q.M = np; q.N = d.dap;
      if ((r = gemm(q, "bwd: d dirs")) != NMR_OK) return r;
      GemmParams t{};   // d h7 = d feature * Wfeat + d sigma (x) w_alpha (+ d feat if h7 is the tapped layer), masked by h7 > 0
      t.A1 = Ga; t.lda1 = kHid; t.W1 = n.Wtfeat; t.W1p = n.Ptfeat; t.K1 = kHid; t.C = Gb; t.ldc = kHid; t.M = np; t.N = kHid;
      t.r1a = gr + 3; t.ldr1 = 4; t.r1b = n.Walpha;
      if (gf && tap == kLayers - 1) { t.add = gf; t.ldadd = kHid; }
      set_mask(t, 7);
      if ((r = gemm(t, "bwd: d h7")) != NMR_OK) return r;
    }
    float* gin = Gb; float* gout = Ga;
    for (int i = kLayers - 1; i >= 1; --i) {
      if (i == skipl) {   // the skip connection's xyz part: d x += d pre_5 * W5[:, :xyz]
        GemmParams g{};
        g.A1 = gin; g.lda1 = kHid; g.W1 = n.Wtskip_x; g.W1p = n.Ptskip_x; g.K1 = kHid; g.C = Gxs; g.ldc = d.xyzp; g.M = np; g.N = d.xyzp;
        if ((r = gemm(g, "bwd: d x (skip)")) != NMR_OK) return r;
      }
      GemmParams g{};
      g.A1 = gin; g.lda1 = kHid; g.W1 = n.Wt[i]; g.W1p = n.Pt[i]; g.K1 = kHid; g.C = gout; g.ldc = kHid; g.M = np; g.N = kHid;
      if (gf && tap == i - 1) { g.add = gf; g.ldadd = kHid; }
      set_mask(g, i - 1);
      if ((r = gemm(g, "bwd: d h")) != NMR_OK) return r;
      std::swap(gin, gout);
    }
    {
      GemmParams g{};
      g.A1 = gin; g.lda1 = kHid; g.W1 = n.Wt[0]; g.W1p = n.Pt[0]; g.K1 = kHid; g.C = Gx0; g.ldc = d.xyzp; g.M = np; g.N = d.xyzp;
      g.add = Gxs; g.ldadd = d.xyzp;
      if ((r = gemm(g, "bwd: d x")) != NMR_OK) return r;
    }
    launch_unpack_grad(d, Gx0, Gd, d.dap, np, grad_x + p0 * d.in_dim, s);
    NMR_LAUNCHED(h, "unpack_grad");
  }
  return NMR_OK;
}

extern "C" int nmr_coarse_match(nmr_handle* h, const float* im_feat, int32_t m, const float* pt_feat, int32_t n, int32_t d,
                                const uint8_t* im_mask, const uint8_t* pt_mask, float temperature, int32_t temp_div,
                                int32_t mutual, float threshold, float* conf, int64_t* i_ids, int64_t* j_ids,
                                float* mconf, int32_t* count, void* stream) {
  if (!h) return NMR_ERR_INVALID;
  if (!im_feat || !pt_feat || !i_ids || !j_ids || !mconf || !count) NMR_FAIL(h, NMR_ERR_INVALID, "nmr_coarse_match: null argument");
  if (m <= 0 || n <= 0 || d <= 0 || d > 512) NMR_FAIL(h, NMR_ERR_INVALID, "nmr_coarse_match: need m, n > 0 and 0 < d <= 512");
  if (temperature == 0.f) NMR_FAIL(h, NMR_ERR_INVALID, "nmr_coarse_match: temperature == 0");
  NMR_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t s = (cudaStream_t)stream;
  const int dp = round_up(d, 16), R = 32;
  auto al = [](size_t b) { return (b + 255) / 256 * 256; };
  const size_t o_im = 0, o_pt = o_im + al((size_t)m * dp * 4), o_sim = o_pt + al((size_t)n * dp * 4),
               o_r = o_sim + al((size_t)m * n * 4), o_c = o_r + al((size_t)m * 4) * 6, o_cp = o_c + al((size_t)n * 4) * 3,
               total = o_cp + al((size_t)R * n * 4) * 3;
  NMR_CUDA(h, h->match.reserve(total));
  char* b = (char*)h->match.p;
  float* imn = (float*)(b + o_im); float* ptn = (float*)(b + o_pt); float* sim = (float*)(b + o_sim);
  launch_normalize_rows(im_feat, m, d, dp, imn, s);
  launch_normalize_rows(pt_feat, n, d, dp, ptn, s);
  NMR_LAUNCHED(h, "normalize_rows", 2);
  GemmParams g{};
  g.A1 = imn; g.lda1 = dp; g.W1 = ptn; g.K1 = dp; g.C = sim; g.ldc = n; g.M = m; g.N = n; g.act = 0;
  if (h->tc_gemm && gemm_tc_supported(g)) {        // the one contraction of the matcher: tensor cores, fp32-grade (3 x bf16 split)
    if (launch_gemm_tc(g, s) != 0) NMR_FAIL(h, NMR_ERR_CUDA, "launch_gemm_tc setup failed (similarity)");
  } else {
    launch_sgemm(g, s);
  }
  NMR_LAUNCHED(h, "similarity gemm");
  MatchParams p{};
  p.sim = sim; p.m = m; p.n = n; p.im_mask = im_mask; p.pt_mask = pt_mask; p.temperature = temperature; p.temp_div = temp_div;
  const size_t rs = al((size_t)m * 4), cs = al((size_t)n * 4), ps = al((size_t)R * n * 4);
  p.rmax = (float*)(b + o_r); p.rsum = (float*)(b + o_r + rs); p.rcmax = (float*)(b + o_r + 2 * rs);
  p.flag = (int*)(b + o_r + 3 * rs); p.jsel = (int*)(b + o_r + 4 * rs); p.csel = (float*)(b + o_r + 5 * rs);
  p.cmax = (float*)(b + o_c); p.csum = (float*)(b + o_c + cs); p.ccmax = (float*)(b + o_c + 2 * cs);
  p.cpmax = (float*)(b + o_cp); p.cpsum = (float*)(b + o_cp + ps); p.ccpmax = (float*)(b + o_cp + 2 * ps);
  p.conf = conf; p.R = R; p.mutual = mutual; p.threshold = threshold;
  p.i_ids = i_ids; p.j_ids = j_ids; p.mconf = mconf; p.count = count;
  const int nl = launch_match_passes(p, s);
  NMR_LAUNCHED(h, "coarse match passes", nl);
  return NMR_OK;
}

extern "C" int nmr_stage_mlp_probe(nmr_handle* h, int which, const float* x, int64_t P, int32_t step, float* out,
                                   void* stream) {
  if (!h) return NMR_ERR_INVALID;
  if (which != 0 && which != 1) NMR_FAIL(h, NMR_ERR_INVALID, "nmr_stage_mlp_probe: which=%d", which);
  if (h->cfg.precision == NMR_PREC_FP32) NMR_FAIL(h, NMR_ERR_INVALID, "nmr_stage_mlp_probe: tcgen05 path only (NMR_PREC_BF16 / NMR_PREC_FP16)");
  if (step < 0 || step > 8 || P <= 0 || !x || !out) NMR_FAIL(h, NMR_ERR_INVALID, "nmr_stage_mlp_probe: bad argument");
  if (!h->net[which].set) NMR_FAIL(h, NMR_ERR_STATE, "nmr_stage_mlp_probe: weights not set");
  NMR_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t s = (cudaStream_t)stream;
  const Dims& d = h->d;
  int r;
  if ((r = ensure_workspace(h, 1, P)) != NMR_OK) return r;
  launch_repack_input(d, x, P, (float*)h->X.p, s);
  NMR_LAUNCHED(h, "repack_input");
  std::string terr; int nl = 0;
  if (tc_run_rows(d, h->tc[h->cfg.precision == NMR_PREC_FP16][which], (const float*)h->X.p, P, 3, (float*)h->raw.p, nullptr, &h->tcws.p, &h->tcws.cap, &nl, s,
                  &terr, step + 1, out) != 0)
    NMR_FAIL(h, NMR_ERR_CUDA, "tc_run_rows: %s", terr.c_str());
  NMR_LAUNCHED(h, "tcgen05 mlp probe", nl);
  return NMR_OK;
}

extern "C" int nmr_debug_trap_code(void) { return tc_trap_code(); }
extern "C" int nmr_debug_timeline(long long* host, int n) { return tc_read_timeline(host, n); }
extern "C" int nmr_set_tc_cluster(int cluster_size) { return tc_set_cluster(cluster_size) == 0 ? NMR_OK : NMR_ERR_INVALID; }

extern "C" int nmr_stage_composite(nmr_handle* h, const float* raw, const float* z, const float* rays_d, int64_t n,
                                   int32_t S, uint32_t flags, const float* feats, const float* pts, float* rgb,
                                   float* depth, float* acc, float* disp, float* weights, float* feat_out,
                                   float* pts_out, void* stream) {
  if (!h) return NMR_ERR_INVALID;
  if (n < 0 || (n > 0 && (!raw || !z || !rays_d))) NMR_FAIL(h, NMR_ERR_INVALID, "nmr_stage_composite: null argument");
  if (S < 2 || S > kMaxSamples) NMR_FAIL(h, NMR_ERR_INVALID, "nmr_stage_composite: S=%d (2..%d)", S, kMaxSamples);
  if (pts_out && !pts) NMR_FAIL(h, NMR_ERR_INVALID, "nmr_stage_composite: pts_out needs pts");
  if (n == 0) return NMR_OK;
  NMR_CUDA(h, cudaSetDevice(h->device));
  CompositeParams cp{};
  cp.raw = raw; cp.z = z; cp.rays = nullptr; cp.rays_d = rays_d; cp.n = n; cp.S = S; cp.mip = h->d.mip; cp.flags = flags;
  cp.feats = feats; cp.feats_bf16 = 0; cp.pts_in = pts;
  cp.rgb = rgb; cp.depth = depth; cp.acc = acc; cp.disp = disp; cp.weights = weights;
  cp.feat_out = feats ? feat_out : nullptr; cp.pts_out = pts_out;
  launch_composite(cp, (cudaStream_t)stream);
  NMR_LAUNCHED(h, "composite");
  return NMR_OK;
}
