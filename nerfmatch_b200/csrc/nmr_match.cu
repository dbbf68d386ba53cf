// SURVEY.md 8f row 2 - NeRFMatch-Mini coarse matching on the device, consuming the rendered 3D features:
//   NeRFMatcherCoarse.coarse_matching   nerfmatch/nerfmatch_coarse_trainer.py:223-234
//     normalise -> sim = im . pt^T -> temperature -> mask -> softmax(dim=1) * softmax(dim=2)
//   extract_mutual_matches              nerfmatch/modules/extract_matches.py:13-40 (inference branch)
//     conf > thr, row-max (and column-max when mutual), first such j per image cell, compacted in row order.
// fp32 throughout (the confidence is a product of two softmaxes at temperature 10: bf16 would move matches).
// The similarity GEMM re-uses the fp32 SIMT GEMM (7.9 GFLOP at m=4800, n=6420, d=256: ~0.3 ms); everything
// else is a handful of streaming passes over the [m,n] matrix.
#include "nmr_internal.cuh"

namespace nmr {

__global__ void normalize_rows_kernel(const float* __restrict__ x, int rows, int d, int dp, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= rows) return;
  const float* r = x + (int64_t)row * d;
  float ss = 0.f;
  for (int k = lane; k < d; k += 32) ss = fmaf(r[k], r[k], ss);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float inv = sqrtf(ss) + 1e-6f;           // feat / (feat.norm(dim=-1, keepdim=True) + 1e-6)
  for (int k = lane; k < dp; k += 32) out[(int64_t)row * dp + k] = k < d ? r[k] / inv : 0.f;
}

__device__ __forceinline__ float sim_value(const MatchParams& p, int i, int j) {
  // explicit round-to-nearest ops (no FMA contraction): the selection passes re-evaluate this expression and compare
  // the result for EQUALITY with maxima found by other kernels, so every evaluation must give the same bits
  float v = p.sim[(int64_t)i * p.n + j];
  v = p.temp_div ? __fdiv_rn(v, p.temperature) : __fmul_rn(v, p.temperature);
  const bool ok = (!p.im_mask || p.im_mask[i]) && (!p.pt_mask || p.pt_mask[j]);
  return ok ? v : -1e9f;
}
__device__ __forceinline__ float conf_value(const MatchParams& p, int i, int j) {
  const float v = sim_value(p, i, j);
  return __fmul_rn(__fdiv_rn(expf(__fsub_rn(v, p.cmax[j])), p.csum[j]), __fdiv_rn(expf(__fsub_rn(v, p.rmax[i])), p.rsum[i]));
}
__device__ __forceinline__ void sm_merge(float& m, float& s, float m2, float s2) {
  if (m2 == -INFINITY) return;                       // empty operand (a thread / slab that saw no element): -inf - -inf = NaN
  if (m == -INFINITY) { m = m2; s = s2; return; }
  const float mm = fmaxf(m, m2);
  s = s * expf(m - mm) + s2 * expf(m2 - mm);
  m = mm;
}

// block per row: softmax statistics over n
__global__ void __launch_bounds__(256) row_stats_kernel(MatchParams p) {
  __shared__ float sm[8], ss[8];
  const int i = blockIdx.x;
  float mx = -INFINITY, su = 0.f;
  for (int j = threadIdx.x; j < p.n; j += 256) sm_merge(mx, su, sim_value(p, i, j), 1.f);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sm_merge(mx, su, __shfl_xor_sync(0xffffffffu, mx, o), __shfl_xor_sync(0xffffffffu, su, o));
  if ((threadIdx.x & 31) == 0) { sm[threadIdx.x >> 5] = mx; ss[threadIdx.x >> 5] = su; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) sm_merge(mx, su, sm[w], ss[w]);
    p.rmax[i] = mx; p.rsum[i] = su;
  }
}
// thread per column, rows split in R slabs: partial softmax statistics over m
__global__ void __launch_bounds__(256) col_stats_partial_kernel(MatchParams p) {
  const int j = blockIdx.x * 256 + threadIdx.x;
  if (j >= p.n) return;
  const int rows = (p.m + p.R - 1) / p.R, i0 = blockIdx.y * rows, i1 = min(p.m, i0 + rows);
  float mx = -INFINITY, su = 0.f;
  for (int i = i0; i < i1; ++i) sm_merge(mx, su, sim_value(p, i, j), 1.f);
  p.cpmax[(int64_t)blockIdx.y * p.n + j] = mx; p.cpsum[(int64_t)blockIdx.y * p.n + j] = su;
}
__global__ void col_stats_final_kernel(MatchParams p) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= p.n) return;
  float mx = -INFINITY, su = 0.f;
  for (int r = 0; r < p.R; ++r) {
    const float m2 = p.cpmax[(int64_t)r * p.n + j];
    if (m2 != -INFINITY) sm_merge(mx, su, m2, p.cpsum[(int64_t)r * p.n + j]);
  }
  p.cmax[j] = mx; p.csum[j] = su;
}
// block per row: conf (optionally stored) and its row maximum
__global__ void __launch_bounds__(256) conf_row_kernel(MatchParams p) {
  __shared__ float sm[8];
  const int i = blockIdx.x;
  float mx = 0.f;
  for (int j = threadIdx.x; j < p.n; j += 256) {
    const float c = conf_value(p, i, j);
    if (p.conf) p.conf[(int64_t)i * p.n + j] = c;
    mx = fmaxf(mx, c);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x == 0) { for (int w = 1; w < 8; ++w) mx = fmaxf(mx, sm[w]); p.rcmax[i] = mx; }
}
__global__ void __launch_bounds__(256) conf_col_partial_kernel(MatchParams p) {
  const int j = blockIdx.x * 256 + threadIdx.x;
  if (j >= p.n) return;
  const int rows = (p.m + p.R - 1) / p.R, i0 = blockIdx.y * rows, i1 = min(p.m, i0 + rows);
  float mx = 0.f;
  for (int i = i0; i < i1; ++i) mx = fmaxf(mx, conf_value(p, i, j));
  p.ccpmax[(int64_t)blockIdx.y * p.n + j] = mx;
}
__global__ void conf_col_final_kernel(MatchParams p) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= p.n) return;
  float mx = 0.f;
  for (int r = 0; r < p.R; ++r) mx = fmaxf(mx, p.ccpmax[(int64_t)r * p.n + j]);
  p.ccmax[j] = mx;
}
// block per row: first j that passes the threshold / row-max / (column-max) tests
__global__ void __launch_bounds__(256) select_kernel(MatchParams p) {
  __shared__ int sj[8];
  const int i = blockIdx.x;
  const float rm = p.rcmax[i];
  int best = 0x7fffffff;
  for (int j = threadIdx.x; j < p.n; j += 256) {
    const float c = conf_value(p, i, j);
    if (c > p.threshold && c == rm && (!p.mutual || c == p.ccmax[j])) { best = min(best, j); }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
  if ((threadIdx.x & 31) == 0) sj[threadIdx.x >> 5] = best;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) best = min(best, sj[w]);
    const bool hit = best != 0x7fffffff;
    p.flag[i] = hit ? 1 : 0;
    p.jsel[i] = hit ? best : 0;
    p.csel[i] = hit ? conf_value(p, i, best) : 0.f;
  }
}
// single block: ordered compaction (torch.where order = increasing i)
__global__ void __launch_bounds__(1024) compact_kernel(MatchParams p) {
  __shared__ int warp_tot[32];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < p.m; base += 1024) {
    const int i = base + threadIdx.x;
    const int f = (i < p.m) ? p.flag[i] : 0;
    int x = f;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if ((threadIdx.x & 31) >= o) x += y; }
    if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x < 32) {
      int t = warp_tot[threadIdx.x];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, t, o); if (threadIdx.x >= o) t += y; }
      warp_tot[threadIdx.x] = t;
    }
    __syncthreads();
    const int before = carry + ((threadIdx.x >> 5) ? warp_tot[(threadIdx.x >> 5) - 1] : 0) + x - f;
    if (f) { p.i_ids[before] = i; p.j_ids[before] = p.jsel[i]; p.mconf[before] = p.csel[i]; }
    __syncthreads();
    if (threadIdx.x == 0) carry += warp_tot[31];
    __syncthreads();
  }
  if (threadIdx.x == 0) *p.count = carry;
}

void launch_normalize_rows(const float* x, int rows, int d, int dp, float* out, cudaStream_t s) {
  if (rows == 0) return;
  normalize_rows_kernel<<<(rows * 32 + 255) / 256, 256, 0, s>>>(x, rows, d, dp, out);
}

int launch_match_passes(const MatchParams& p, cudaStream_t s) {
  const dim3 cg((p.n + 255) / 256, p.R);
  row_stats_kernel<<<p.m, 256, 0, s>>>(p);
  col_stats_partial_kernel<<<cg, 256, 0, s>>>(p);
  col_stats_final_kernel<<<(p.n + 255) / 256, 256, 0, s>>>(p);
  conf_row_kernel<<<p.m, 256, 0, s>>>(p);
  int launches = 4;
  if (p.mutual) {
    conf_col_partial_kernel<<<cg, 256, 0, s>>>(p);
    conf_col_final_kernel<<<(p.n + 255) / 256, 256, 0, s>>>(p);
    launches += 2;
  }
  select_kernel<<<p.m, 256, 0, s>>>(p);
  compact_kernel<<<1, 1024, 0, s>>>(p);
  return launches + 2;
}

}  // namespace nmr
