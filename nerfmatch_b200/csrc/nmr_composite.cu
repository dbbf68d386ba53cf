// Kernel (c): warp-per-ray volume rendering, feature / depth / point compositing and the fused
// hierarchical re-sampling (CDF inversion) for the next pass.  Compiled with -fmad=false.
//
// Reference behaviour restated here:
//   nerfmatch/nerf/render_utils.py:176-230   volume_render_radiance_field
//   nerfmatch/nerf/renderer.py:250-281       feature / point compositing ("lin" | "max")
//   nerfmatch/nerf/render_utils.py:453-552, 555-615  mip resampling (blur, pdf, cdf, find_interval)
//   nerfmatch/nerf/render_utils.py:132-148, 233-264  classic sample_pdf + sort
//   nerfmatch/utils/geometry.py:76-85        unnormaliz_pts
//
// One warp owns one ray.  The S samples of the ray live in shared memory; the feature rows
// ([S,256] fp32 or bf16, the only large read) are streamed with 16-byte loads, 8 channels per
// lane, several rows in flight.  Scans that feed the CDF are done the way torch's CPU kernels do
// them (sequential, float64 accumulator, each prefix rounded to fp32) so that interval indices are
// bit-identical to the oracle; the 0.1 us this costs per ray is invisible next to the MLP.
#include "nmr_internal.cuh"

namespace nmr {

constexpr int kWarpsPerBlock = 4;
constexpr int kSmemPerWarp = 4 * (kMaxSamples + 8);   // floats: w, z, cdf/aux, tmp

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Sum and running sums of the n fp32 values x[0..n) in float64, spread over the warp (lane l owns a contiguous segment;
// the lane totals are combined by a shuffle scan).  The oracle (and torch's CPU kernels) accumulate sequentially in
// float64; the association here differs, but the result does not: every value summed in this file is an fp32 number
// within a 2^21 dynamic range of the others (weights carry a +1e-5 or +0.01 floor and are <= ~1), so every partial sum
// of <= 256 of them is exactly representable in a 53-bit significand - float64 addition is EXACT here and therefore
// associative.  Returns this lane's exclusive prefix (sum of the segments before it) and the total.
__device__ __forceinline__ void warp_prefix_f64(const float* x, int n, int lane, int& s0, int& s1, double& excl, double& total) {
  const int seg = (n + 31) / 32;
  s0 = min(n, lane * seg); s1 = min(n, s0 + seg);
  double local = 0.0;
  for (int s = s0; s < s1; ++s) local += (double)x[s];
  double incl = local;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const double up = __shfl_up_sync(0xffffffffu, incl, off);
    if (lane >= off) incl += up;
  }
  total = __shfl_sync(0xffffffffu, incl, 31);
  excl = incl - local;
}

__device__ __forceinline__ void frustum_tmean(float t0, float t1, float* t_mean) {
  float mu = (t0 + t1) / 2.0f, hw = (t1 - t0) / 2.0f;
  float mu2 = mu * mu, hw2 = hw * hw;
  float denom = fmaxf(kEps32, 3.0f * mu2 + hw2);
  *t_mean = mu + (2.0f * mu * hw2) / denom;
}

// number of entries of the non-decreasing table a[0..n) that are <= u  (searchsorted right=True).
// Fixed-step, branch-free form (the warp's lanes search different keys: a data-dependent loop runs every lane for the
// longest search and pays a branch per step): lo grows by 2^k whenever the entry that step would skip is still <= u.
// For a sorted table this is the partition point, i.e. exactly what the bisection returns.
__device__ __forceinline__ int upper_bound(const float* a, int n, float u) {
  int lo = 0;
  for (int step = 1 << (31 - __clz(max(n, 1))); step > 0; step >>= 1) {   // n is warp-uniform: no divergence
    const int t = lo + step;
    if (t <= n && a[t - 1] <= u) lo = t;
  }
  return lo;
}
// number of entries of the non-decreasing table a[0..n) that are < v  (searchsorted right=False)
__device__ __forceinline__ int lower_bound(const float* a, int n, float v) {
  int lo = 0;
  for (int step = 1 << (31 - __clz(max(n, 1))); step > 0; step >>= 1) {
    const int t = lo + step;
    if (t <= n && a[t - 1] < v) lo = t;
  }
  return lo;
}

// eight packed 16-bit features (fmt 1: bf16, 2: fp16 - the tcgen05 kernel's operand format) times w
__device__ __forceinline__ void accum8_16(float acc[8], uint4 v, float w, int fmt) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
  const __half2* hh = reinterpret_cast<const __half2*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = fmt == 2 ? __half22float2(hh[i]) : __bfloat1622float2(h[i]);
    acc[2 * i] = fmaf(w, f.x, acc[2 * i]);
    acc[2 * i + 1] = fmaf(w, f.y, acc[2 * i + 1]);
  }
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32) composite_kernel(CompositeParams p) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t ray = (int64_t)blockIdx.x * kWarpsPerBlock + warp;
  if (ray >= p.n) return;
  float* sw = smem + warp * kSmemPerWarp;        // weights
  float* sz = sw + (kMaxSamples + 8);            // z / t fence posts
  float* sa = sz + (kMaxSamples + 8);            // alpha, then pdf / cdf
  float* st = sa + (kMaxSamples + 8);            // transmittance terms, then scratch
  const int S = p.S;
  const int nz = p.mip ? S + 1 : S;

  const float* r = p.rays ? p.rays + ray * p.ncols : nullptr;
  float o[3] = {0.f, 0.f, 0.f}, d[3];
  if (r) {
    o[0] = r[0]; o[1] = r[1]; o[2] = r[2];
    d[0] = r[3]; d[1] = r[4]; d[2] = r[5];
  } else {
    d[0] = p.rays_d[ray * 3]; d[1] = p.rays_d[ray * 3 + 1]; d[2] = p.rays_d[ray * 3 + 2];
  }
  // This lane's raw rows (sigmoid rgb, raw sigma) of samples s = lane + 32 i, fetched ONCE as 16-byte loads and kept in
  // registers for both the alpha and the colour loop; issued together with the z loads so that the warp pays one memory
  // latency at its start instead of three dependent ones (ncu: the z / sigma / rgb loads were 52 % of the stall samples).
  constexpr int kPerLane = kMaxSamples / 32;
  float4 rv[kPerLane];
  if (p.weights_in == nullptr) {
    const float4* rawv = reinterpret_cast<const float4*>(p.raw) + ray * S;
#pragma unroll
    for (int i = 0; i < kPerLane; ++i) {
      const int s = lane + 32 * i;
      rv[i] = s < S ? __ldcs(rawv + s) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  for (int j = lane; j < nz; j += 32) sz[j] = p.z[ray * nz + j];
  __syncwarp();

  if (p.weights_in == nullptr) {
    // ---- alpha and the exclusive transmittance product ------------------------------------
    const float dnorm = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
#pragma unroll
    for (int i = 0; i < kPerLane; ++i) {
      const int s = lane + 32 * i;
      if (s >= S) break;
      float sraw = rv[i].w;
      if (p.sigma_noise) sraw = sraw + p.sigma_noise[ray * S + s];   // training mode: relu(raw + noise) (render_utils.py:191-195)
      float sigma = fmaxf(sraw, 0.0f);
      float dist = (p.mip || s < S - 1) ? (sz[s + 1] - sz[s]) : 1e10f;
      dist = dist * dnorm;
      float alpha = 1.0f - expf(-sigma * dist);
      sa[s] = alpha;
      st[s] = 1.0f - alpha + 1e-10f;
    }
    __syncwarp();
    // cumprod as torch's CPU kernel: float64 running product, each prefix rounded to fp32.
    const int seg = (S + 31) / 32;
    const int s0 = min(S, lane * seg), s1 = min(S, s0 + seg);
    double prod = 1.0;
    for (int s = s0; s < s1; ++s) prod *= (double)st[s];
    double incl = prod;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      double up = __shfl_up_sync(0xffffffffu, incl, off);
      if (lane >= off) incl *= up;
    }
    double run = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) run = 1.0;
    for (int s = s0; s < s1; ++s) {
      const float T = (float)run;                  // exclusive transmittance in front of sample s
      const double f = (double)st[s];
      sw[s] = sa[s] * T;
      st[s] = T;                                   // kept for the segment-wise feature compositing below
      run *= f;
    }
    __syncwarp();
  } else {
    for (int s = lane; s < S; s += 32) sw[s] = p.weights_in[ray * S + s];
    __syncwarp();
  }

  const bool render = (p.weights_in == nullptr);
  if (render) {
    // ---- rgb / depth / acc / points ----------------------------------------------------------
    float acc = 0.f, dep = 0.f, c0 = 0.f, c1 = 0.f, c2 = 0.f, p0 = 0.f, p1 = 0.f, p2 = 0.f;
    float best = -1.0f; int besti = 0x7fffffff;
#pragma unroll
    for (int i = 0; i < kPerLane; ++i) {
      const int s = lane + 32 * i;
      if (s >= S) break;
      float w = sw[s];
      float zmid = p.mip ? 0.5f * (sz[s] + sz[s + 1]) : sz[s];
      acc += w; dep += w * zmid;
      c0 += w * rv[i].x; c1 += w * rv[i].y; c2 += w * rv[i].z;
      if (p.pts_out || p.pt3d) {
        float q[3];
        if (p.pts_in) {
          const float* pi = p.pts_in + (ray * S + s) * 3;
          q[0] = pi[0]; q[1] = pi[1]; q[2] = pi[2];
        } else {
          float t;
          if (p.mip) frustum_tmean(sz[s], sz[s + 1], &t); else t = sz[s];
          q[0] = p.mip ? d[0] * t + o[0] : o[0] + d[0] * t;
          q[1] = p.mip ? d[1] * t + o[1] : o[1] + d[1] * t;
          q[2] = p.mip ? d[2] * t + o[2] : o[2] + d[2] * t;
        }
        p0 += w * q[0]; p1 += w * q[1]; p2 += w * q[2];
      }
      if (w > best) { best = w; besti = s; }
    }
    acc = warp_sum(acc); dep = warp_sum(dep);
    c0 = warp_sum(c0); c1 = warp_sum(c1); c2 = warp_sum(c2);
    p0 = warp_sum(p0); p1 = warp_sum(p1); p2 = warp_sum(p2);
    const bool fmax = (p.flags & NMR_FLAG_FEAT_MAX) != 0;
    if (fmax) {
      // first index of the maximum weight (torch.max(dim) on CPU returns the first occurrence)
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        float ob = __shfl_xor_sync(0xffffffffu, best, off);
        int oi = __shfl_xor_sync(0xffffffffu, besti, off);
        if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; }
      }
      if (besti >= S) besti = 0;
      if (p.pts_out || p.pt3d) {
        if (p.pts_in) {
          const float* pi = p.pts_in + (ray * S + besti) * 3;
          p0 = pi[0]; p1 = pi[1]; p2 = pi[2];
        } else {
          float t;
          if (p.mip) frustum_tmean(sz[besti], sz[besti + 1], &t); else t = sz[besti];
          p0 = p.mip ? d[0] * t + o[0] : o[0] + d[0] * t;
          p1 = p.mip ? d[1] * t + o[1] : o[1] + d[1] * t;
          p2 = p.mip ? d[2] * t + o[2] : o[2] + d[2] * t;
        }
      }
    }
    if (lane == 0) {
      if (p.flags & NMR_FLAG_WHITE_BG) { float bg = 1.0f - acc; c0 += bg; c1 += bg; c2 += bg; }
      if (p.rgb) { p.rgb[ray * 3] = c0; p.rgb[ray * 3 + 1] = c1; p.rgb[ray * 3 + 2] = c2; }
      if (p.depth) p.depth[ray] = dep;
      if (p.acc) p.acc[ray] = acc;
      if (p.disp) { float q = dep / acc; p.disp[ray] = (q != q) ? q : 1.0f / fmaxf(1e-10f, q); }
      if (p.pts_out) { p.pts_out[ray * 3] = p0; p.pts_out[ray * 3 + 1] = p1; p.pts_out[ray * 3 + 2] = p2; }
      if (p.pt3d && p.has_unnorm) {
        const float* u = p.unnorm;
#pragma unroll
        for (int i = 0; i < 3; ++i)
          p.pt3d[ray * 3 + i] = fmaf(u[4 * i], p0, fmaf(u[4 * i + 1], p1, fmaf(u[4 * i + 2], p2, u[4 * i + 3])));
      }
    }
    if (p.weights) for (int s = lane; s < S; s += 32) p.weights[ray * S + s] = sw[s];

    // ---- feature compositing ---------------------------------------------------------------------
    if (p.partials && p.feat_out) {
      // The tcgen05 kernel already reduced the tapped features inside each 128-point tile with tile-local weights
      // (one 256-vector per ray segment of the tile): scale every segment of this ray by the transmittance in front of
      // its first sample and add.  ~1 KB per segment instead of S x 512 B of per-sample features.
      const int64_t p0 = ray * S;
      float a[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = 0.f;
      for (int64_t t = p0 >> 7; t <= (p0 + S - 1) >> 7; ++t) {
        const int64_t tstart = t << 7;
        const int sfirst = tstart > p0 ? (int)(tstart - p0) : 0;
        const int seg = (int)(ray - tstart / S);
        const float T = st[sfirst];
        const float4* src = reinterpret_cast<const float4*>(p.partials + ((size_t)t * 4 + seg) * kHid) + lane * 2;
        const float4 x = __ldcs(src), y = __ldcs(src + 1);
        a[0] = fmaf(T, x.x, a[0]); a[1] = fmaf(T, x.y, a[1]); a[2] = fmaf(T, x.z, a[2]); a[3] = fmaf(T, x.w, a[3]);
        a[4] = fmaf(T, y.x, a[4]); a[5] = fmaf(T, y.y, a[5]); a[6] = fmaf(T, y.z, a[6]); a[7] = fmaf(T, y.w, a[7]);
      }
      float4* out = reinterpret_cast<float4*>(p.feat_out + ray * kHid) + lane * 2;
      out[0] = make_float4(a[0], a[1], a[2], a[3]);
      out[1] = make_float4(a[4], a[5], a[6], a[7]);
    } else if (p.feats && p.feat_out) {
      float a[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = 0.f;
      if (p.feats_bf16) {
        // tiled layout written by the tcgen05 kernel's tap warps (fusion off): 64 KB per 128 points, eight
        // [128 x 32] blocks, rows of 64 B with the UMMA 64B swizzle.  Lane l reads the 16-byte chunk
        // holding channels [8l, 8l+8) of the sample's row.
        const uint8_t* fb = reinterpret_cast<const uint8_t*>(p.feats);
        const int jb = (lane >> 2) * 8192, cc = lane & 3;
        auto addr = [&](int s) -> const uint4* {
          const int64_t pt = ray * S + s;
          const int rr = (int)(pt & 127);
          return reinterpret_cast<const uint4*>(fb + (pt >> 7) * 65536 + jb + ((rr >> 3) << 9) + ((rr & 7) << 6) +
                                                ((cc ^ ((rr & 7) >> 1)) << 4));
        };
        if (fmax) {
          accum8_16(a, __ldg(addr(besti)), 1.0f, p.feats_bf16);
        } else {
          int s = 0;
          for (; s + 8 <= S; s += 8) {
            uint4 v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = __ldcs(addr(s + k));
#pragma unroll
            for (int k = 0; k < 8; ++k) accum8_16(a, v[k], sw[s + k], p.feats_bf16);
          }
          for (; s < S; ++s) accum8_16(a, __ldcs(addr(s)), sw[s], p.feats_bf16);
        }
      } else {
        // lane owns channels [4*lane, +4) and [128 + 4*lane, +4): every load instruction is a
        // fully used, contiguous 512-byte warp request.
        const float4* base = reinterpret_cast<const float4*>(p.feats) + (ray * S) * 64 + lane;
        if (fmax) {
          float4 x = __ldg(base + (int64_t)besti * 64), y = __ldg(base + (int64_t)besti * 64 + 32);
          a[0] = x.x; a[1] = x.y; a[2] = x.z; a[3] = x.w; a[4] = y.x; a[5] = y.y; a[6] = y.z; a[7] = y.w;
        } else {
          int s = 0;
          for (; s + 4 <= S; s += 4) {
            float4 x[4], y[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              x[k] = __ldcs(base + (int64_t)(s + k) * 64);
              y[k] = __ldcs(base + (int64_t)(s + k) * 64 + 32);
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              float w = sw[s + k];
              a[0] = fmaf(w, x[k].x, a[0]); a[1] = fmaf(w, x[k].y, a[1]);
              a[2] = fmaf(w, x[k].z, a[2]); a[3] = fmaf(w, x[k].w, a[3]);
              a[4] = fmaf(w, y[k].x, a[4]); a[5] = fmaf(w, y[k].y, a[5]);
              a[6] = fmaf(w, y[k].z, a[6]); a[7] = fmaf(w, y[k].w, a[7]);
            }
          }
          for (; s < S; ++s) {
            float w = sw[s];
            float4 x = __ldcs(base + (int64_t)s * 64), y = __ldcs(base + (int64_t)s * 64 + 32);
            a[0] = fmaf(w, x.x, a[0]); a[1] = fmaf(w, x.y, a[1]); a[2] = fmaf(w, x.z, a[2]); a[3] = fmaf(w, x.w, a[3]);
            a[4] = fmaf(w, y.x, a[4]); a[5] = fmaf(w, y.y, a[5]); a[6] = fmaf(w, y.z, a[6]); a[7] = fmaf(w, y.w, a[7]);
          }
        }
      }
      if (p.feats_bf16) {   // lane owns channels [8*lane, +8)
        float4* out = reinterpret_cast<float4*>(p.feat_out + ray * kHid) + lane * 2;
        out[0] = make_float4(a[0], a[1], a[2], a[3]);
        out[1] = make_float4(a[4], a[5], a[6], a[7]);
      } else {
        float4* out = reinterpret_cast<float4*>(p.feat_out + ray * kHid) + lane;
        out[0] = make_float4(a[0], a[1], a[2], a[3]);
        out[32] = make_float4(a[4], a[5], a[6], a[7]);
      }
    }
  }

  if (p.z_next == nullptr) return;
  __syncwarp();

  // ---- fused hierarchical re-sampling ---------------------------------------------------------
  if (p.mip) {
    // blur: edge pad, max of neighbours, mean of neighbours, + resample_padding (:583-590)
    for (int s = lane; s < S; s += 32) {
      float wm = sw[max(s - 1, 0)], w0 = sw[s], wp = sw[min(s + 1, S - 1)];
      st[s] = 0.5f * (fmaxf(wm, w0) + fmaxf(w0, wp)) + p.resample_padding;
    }
    __syncwarp();
    int q0, q1; double excl, total;
    warp_prefix_f64(st, S, lane, q0, q1, excl, total);
    const float wsum = (float)total;
    float padding = fmaxf(0.0f, 1e-5f - wsum);
    float wtot = wsum + padding;
    float padw = padding / (float)S;
    __syncwarp();
    for (int s = lane; s < S; s += 32) st[s] = (st[s] + padw) / wtot;     // pdf
    __syncwarp();
    warp_prefix_f64(st, S, lane, q0, q1, excl, total);                     // cdf: float64 running sum, each prefix rounded to fp32
    for (int s = q0; s < q1; ++s) { excl += (double)st[s]; if (s < S - 1) sa[s + 1] = fminf(1.0f, (float)excl); }
    if (lane == 0) { sa[0] = 0.0f; sa[S] = 1.0f; }
    __syncwarp();
    const int S1 = S + 1;
    const float sstep = (float)(1.0 / (double)S1);
    for (int j = lane; j < S1; j += 32) {
      float u;
      if (p.u_jitter != nullptr || p.rng_mode == NMR_RNG_PHILOX) {
        float jit = p.u_jitter ? p.u_jitter[ray * S1 + j]
                               : philox_uniform(p.seed, (uint64_t)(p.ray0 + ray) * 1024u + j, 1u) *
                                     ((float)(1.0 / (double)S1 - (double)kEps32));
        float uu = (float)j * sstep;
        u = fminf((uu + uu) + jit, 1.0f - kEps32);                       // u + u + jitter (:480-490)
      } else {
        u = torch_linspace(1.0f - kEps32, S1, j);                         // randomized=False (:493-496)
      }
      int idx = upper_bound(sa, S1, u) - 1;
      idx = min(max(idx, 0), S - 1);
      float x0 = sz[idx], x1 = sz[idx + 1], y0 = sa[idx], y1 = sa[idx + 1];
      float t = (u - y0) / (y1 - y0);
      if (t != t) t = 0.0f;                                               // nan_to_num(.., 0)
      t = fminf(fmaxf(t, 0.0f), 1.0f);
      p.z_next[ray * S1 + j] = x0 + t * (x1 - x0);
      if (p.idx_out) p.idx_out[ray * S1 + j] = idx;
    }
  } else {
    // sample_pdf(bins = z mid-points [S-1], weights[1:-1] + 1e-5 [S-2], N = nf, det=True)
    const int nb = S - 1, nw = S - 2, N = p.nf;
    for (int i = lane; i < nw; i += 32) st[i] = sw[i + 1] + 1e-5f;
    __syncwarp();
    int q0, q1; double excl, total;
    warp_prefix_f64(st, nw, lane, q0, q1, excl, total);
    const float wsum = (float)total;
    __syncwarp();
    for (int i = lane; i < nw; i += 32) st[i] = st[i] / wsum;            // pdf
    __syncwarp();
    warp_prefix_f64(st, nw, lane, q0, q1, excl, total);                   // cdf
    for (int i = q0; i < q1; ++i) { excl += (double)st[i]; sa[i + 1] = (float)excl; }
    if (lane == 0) sa[0] = 0.0f;
    __syncwarp();
    // st is free again: hold the combined [z | samples] list there (S + N <= 2 * kMaxSamples)
    float* comb = st;   // S + N <= kMaxSamples floats (checked in nmr_create)
    for (int j = lane; j < S; j += 32) comb[j] = sz[j];
    for (int j = lane; j < N; j += 32) {
      float u;
      if (p.perturb && (p.u_jitter != nullptr || p.rng_mode == NMR_RNG_PHILOX))       // sample_pdf(det=False): u = torch.rand (:246-247)
        u = p.u_jitter ? p.u_jitter[ray * N + j] : philox_uniform(p.seed, (uint64_t)(p.ray0 + ray) * 1024u + j, 1u);
      else
        u = torch_linspace(1.0f, N, j);
      int inds = upper_bound(sa, nb, u);
      int below = max(0, inds - 1), above = min(nb - 1, inds);
      float cb = sa[below], ca = sa[above];
      float bb = 0.5f * (sz[below + 1] + sz[below]), ba = 0.5f * (sz[above + 1] + sz[above]);
      float denom = ca - cb;
      if (denom < 1e-5f) denom = 1.0f;
      float t = (u - cb) / denom;
      comb[S + j] = bb + t * (ba - bb);
      if (p.idx_out) p.idx_out[ray * N + j] = inds;
    }
    __syncwarp();
    // sort(cat(z, samples)).  Both lists are (normally) already sorted, so the stable rank of an element
    // is its own index plus a binary search in the other list: z_i -> i + #{samples < z_i},
    // sample_j -> j + #{z <= sample_j} (ties keep the concatenation order, like a stable sort).  If fp
    // round-off ever breaks monotonicity the warp falls back to the O(T^2) counting rank, which needs no
    // assumption; torch.sort returns the same multiset either way.
    const int T = S + N;
    bool mono = true;
    for (int j = lane; j < T; j += 32)
      if (j != 0 && j != S && comb[j] < comb[j - 1]) mono = false;
    if (__all_sync(0xffffffffu, mono)) {
      for (int e = lane; e < T; e += 32) {
        float v = comb[e];
        int rank;
        if (e < S) {            // #{samples < v}
          rank = e + lower_bound(comb + S, N, v);
        } else {
          rank = (e - S) + upper_bound(comb, S, v);
        }
        p.z_next[ray * T + rank] = v;
      }
    } else {
      for (int e = lane; e < T; e += 32) {
        float v = comb[e];
        int rank = 0;
        for (int f = 0; f < T; ++f) {
          float q = comb[f];
          rank += (q < v) || (q == v && f < e);
        }
        p.z_next[ray * T + rank] = v;
      }
    }
  }
}

void launch_composite(const CompositeParams& p, cudaStream_t s) {
  if (p.n == 0) return;
  size_t smem = kWarpsPerBlock * kSmemPerWarp * sizeof(float);
  unsigned blocks = (unsigned)((p.n + kWarpsPerBlock - 1) / kWarpsPerBlock);
  composite_kernel<<<blocks, kWarpsPerBlock * 32, smem, s>>>(p);
}

}  // namespace nmr
