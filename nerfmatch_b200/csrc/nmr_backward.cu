// Backward kernels of the differentiable render used by the iNeRF pose refinement:
//   nerfmatch/nerfmatch_evaluator.py:288-500 differentiates, w.r.t. the camera pose,
//     positional encodings of pose-dependent points / directions   (embedding.py:35-46, 66-84)
//     -> fine NeRF MLP                                              (models/nerf.py:94-144)
//     -> volume_render_radiance_field                               (render_utils.py:176-230)
//   with the sample positions z and the Gaussian variances held constant (taken from detached rays).
// Here: the input gradients of those three stages.  The MLP's dgrad GEMM chain itself runs on the GEMM kernels (tensor
// cores: nmr_gemm_tc.cu; fp32 SIMT: nmr_mlp_simt.cu) with the mask / rank-1 / add epilogue; this file holds the non-GEMM pieces.
// Compiled with -fmad=false like the forward sampling / compositing kernels.
#include "nmr_internal.cuh"
#include "nmr_encode.cuh"

namespace nmr {

// ---- rgb head: d raw[:, 0:3] (after sigmoid) -> d (views hidden, pre-ReLU)  (models/nerf.py:114-117 backwards) -------------
// Gv[p, j] = (sum_c g_c * rgb_c * (1 - rgb_c) * Wrgb[c, j]) * [V[p, j] > 0]
__global__ void rgb_head_backward_kernel(const float* __restrict__ grad_raw, const float* __restrict__ raw, const float* __restrict__ V,
                                         const uint32_t* __restrict__ mask_bits, int64_t row0, const float* __restrict__ wrgb, int64_t P,
                                         float* __restrict__ Gv) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P * kHalf) return;
  const int64_t p = i >> 7;
  const int j = (int)(i & (kHalf - 1));
  const float4 g = *reinterpret_cast<const float4*>(grad_raw + p * 4);
  const float4 r = *reinterpret_cast<const float4*>(raw + p * 4);
  const float a0 = g.x * (r.x * (1.0f - r.x)), a1 = g.y * (r.y * (1.0f - r.y)), a2 = g.z * (r.z * (1.0f - r.z));
  const float v = a0 * wrgb[j] + a1 * wrgb[kHalf + j] + a2 * wrgb[2 * kHalf + j];
  const bool on = V ? V[i] > 0.0f : ((mask_bits[mask_word_index(row0 + p, 8, j >> 5)] >> (j & 31)) & 1u) != 0;
  Gv[i] = on ? v : 0.0f;
}
void launch_rgb_head_backward(const float* grad_raw, const float* raw, const float* V, const uint32_t* mask_bits, int64_t row0,
                              const float* wrgb, int64_t P, float* Gv, cudaStream_t s) {
  if (P == 0) return;
  rgb_head_backward_kernel<<<(unsigned)((P * kHalf + 255) / 256), 256, 0, s>>>(grad_raw, raw, V, mask_bits, row0, wrgb, P, Gv);
}

// ---- padded gradient rows [P, kx] -> reference column order [P, in_dim] --------------------------------------------------------
__global__ void unpack_grad_kernel(Dims d, const float* __restrict__ Gx, const float* __restrict__ Gd, int ldgd, int64_t P,
                                   float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P * d.in_dim) return;
  const int64_t p = i / d.in_dim;
  const int j = (int)(i - p * d.in_dim);
  out[i] = j < d.xyz ? Gx[p * d.xyzp + j] : Gd[p * ldgd + (j - d.xyz)];
}
void launch_unpack_grad(const Dims& d, const float* Gx, const float* Gd, int ldgd, int64_t P, float* out, cudaStream_t s) {
  if (P == 0) return;
  unpack_grad_kernel<<<(unsigned)((P * d.in_dim + 255) / 256), 256, 0, s>>>(d, Gx, Gd, ldgd, P, out);
}

// ---- positional encodings: d enc -> d x (the variances of the IPE are constants of the refinement) -----------------------------
// mode 0: IPE  enc = [e_k sin(2^k x) | e_k sin(2^k x + pi/2)],  e_k = exp(-0.5 var 4^k)          (embedding.py:66-79)
// mode 1: PE   enc = [sin(2^k x) | sin(2^k x + pi/2) | x]                                        (embedding.py:80-84)
// mode 2: Fourier enc = [x, sin(2^0 x), cos(2^0 x), sin(2^1 x), ...]                             (embedding.py:35-46)
__global__ void encode_backward_kernel(int mode, int F, int dim, const float* __restrict__ x, const float* __restrict__ var,
                                       const float* __restrict__ gout, int64_t P, float* __restrict__ gx) {
  const int64_t pt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (pt >= P) return;
  const float* g = gout + pt * dim;
  float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float xc = x[pt * 3 + c];
    const float vc = mode == 0 ? var[pt * 3 + c] : 0.f;
    float a = 0.f;
    for (int k = 0; k < F; ++k) {
      const float sc = (float)(1u << k);
      const float arg = xc * sc;
      if (mode == 2) {
        // d sin(f x) = f cos(f x), d cos(f x) = -f sin(f x); columns 3 + 6k + c (sin) and 3 + 6k + 3 + c (cos)
        float sn, cs;
        sincosf(arg, &sn, &cs);
        a += sc * (g[3 + 6 * k + c] * cs - g[3 + 6 * k + 3 + c] * sn);
      } else {
        const float e = mode == 0 ? expf(-0.5f * (vc * (sc * sc))) : 1.0f;
        const float c0 = cosf(arg), c1 = cosf(arg + kHalfPi);          // the reference's "cos" block is sin(x + fp32(pi/2))
        a += sc * (e * (g[3 * k + c] * c0 + g[3 * F + 3 * k + c] * c1));
      }
    }
    if (mode == 1) a += g[6 * F + c];
    if (mode == 2) a += g[c];
    acc[c] = a;
  }
  gx[pt * 3] = acc[0]; gx[pt * 3 + 1] = acc[1]; gx[pt * 3 + 2] = acc[2];
}
void launch_encode_backward(int kind, const Dims& d, const float* x, const float* var, const float* gout, int64_t P, float* gx,
                            cudaStream_t s) {
  if (P == 0) return;
  const int F = kind == 0 ? d.fx : d.fd;
  int mode, dim;
  if (d.mip && kind == 0 && var != nullptr) { mode = 0; dim = 6 * F; }
  else if (d.mip) { mode = 1; dim = 6 * F + 3; }
  else { mode = 2; dim = 6 * F + 3; }
  encode_backward_kernel<<<(unsigned)((P + 127) / 128), 128, 0, s>>>(mode, F, dim, x, var, gout, P, gx);
}

// ---- volume_render_radiance_field backwards: one warp per ray -------------------------------------------------------------------
// Forward (render_utils.py:176-230): sigma = relu(raw_sigma), delta = dz * |d|, alpha = 1 - exp(-sigma delta),
//   f = 1 - alpha + 1e-10, T_s = prod_{k<s} f_k, w = alpha T, rgb = sum w c (+ 1 - acc), depth = sum w z_mid, acc = sum w.
// Given d rgb / d depth / d acc / d w:  g_w[s] = sum_c g_rgb_c (c_sc - [white]) + g_depth z_mid_s + g_acc + g_weights[s],
//   g_alpha[s] = g_w[s] T_s - (sum_{k>s} g_w[k] w_k) / f_s,   g_sigma = g_alpha exp(-sigma delta) delta [raw_sigma > 0],
//   g_delta = g_alpha exp(-sigma delta) sigma,  g_|d| = sum_s g_delta dz_s,  g_d = g_|d| d / |d|,  g_c[s] = w_s g_rgb.
constexpr int kBwdWarps = 4;
constexpr int kBwdSmem = 5 * (kMaxSamples + 8);   // floats per warp: alpha, f, T, g_w * w, z
__global__ void __launch_bounds__(kBwdWarps * 32) composite_backward_kernel(CompositeBwdParams p) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t ray = (int64_t)blockIdx.x * kBwdWarps + warp;
  if (ray >= p.n) return;
  float* sa = smem + warp * kBwdSmem;
  float* sf = sa + (kMaxSamples + 8);
  float* sT = sf + (kMaxSamples + 8);
  float* sg = sT + (kMaxSamples + 8);
  float* sz = sg + (kMaxSamples + 8);
  const int S = p.S, nz = p.mip ? S + 1 : S;
  const float d0 = p.rays_d[ray * 3], d1 = p.rays_d[ray * 3 + 1], d2 = p.rays_d[ray * 3 + 2];
  const float dnorm = sqrtf(d0 * d0 + d1 * d1 + d2 * d2);
  for (int j = lane; j < nz; j += 32) sz[j] = p.z[ray * nz + j];
  __syncwarp();
  for (int s = lane; s < S; s += 32) {
    const float sigma = fmaxf(p.raw[(ray * S + s) * 4 + 3], 0.0f);
    const float dist = ((p.mip || s < S - 1) ? (sz[s + 1] - sz[s]) : 1e10f) * dnorm;
    const float alpha = 1.0f - expf(-sigma * dist);
    sa[s] = alpha;
    sf[s] = 1.0f - alpha + 1e-10f;
  }
  __syncwarp();
  // exclusive transmittance, as the forward kernel: float64 running product, each prefix rounded to fp32
  const int seg = (S + 31) / 32;
  const int s0 = min(S, lane * seg), s1 = min(S, s0 + seg);
  {
    double prod = 1.0;
    for (int s = s0; s < s1; ++s) prod *= (double)sf[s];
    double incl = prod;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      double up = __shfl_up_sync(0xffffffffu, incl, off);
      if (lane >= off) incl *= up;
    }
    double run = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) run = 1.0;
    for (int s = s0; s < s1; ++s) { sT[s] = (float)run; run *= (double)sf[s]; }
  }
  __syncwarp();
  const float gr0 = p.grad_rgb ? p.grad_rgb[ray * 3] : 0.f, gr1 = p.grad_rgb ? p.grad_rgb[ray * 3 + 1] : 0.f,
              gr2 = p.grad_rgb ? p.grad_rgb[ray * 3 + 2] : 0.f;
  const float gdep = p.grad_depth ? p.grad_depth[ray] : 0.f, gacc = p.grad_acc ? p.grad_acc[ray] : 0.f;
  const float white = (p.flags & NMR_FLAG_WHITE_BG) ? 1.0f : 0.0f;
  for (int s = lane; s < S; s += 32) {
    const float* rw = p.raw + (ray * S + s) * 4;
    const float zmid = p.mip ? 0.5f * (sz[s] + sz[s + 1]) : sz[s];
    float gw = gr0 * (rw[0] - white) + gr1 * (rw[1] - white) + gr2 * (rw[2] - white) + gdep * zmid + gacc;
    if (p.grad_weights) gw += p.grad_weights[ray * S + s];
    const float w = sa[s] * sT[s];
    float* go = p.grad_raw + (ray * S + s) * 4;
    go[0] = w * gr0; go[1] = w * gr1; go[2] = w * gr2;
    sg[s] = gw;
  }
  __syncwarp();
  // suffix sums of g_w[k] w_k over k > s (float64 accumulation: the terms alternate in sign)
  double local = 0.0;
  for (int s = s0; s < s1; ++s) local += (double)sg[s] * (double)(sa[s] * sT[s]);
  double incl = local;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    double dn = __shfl_down_sync(0xffffffffu, incl, off);
    if (lane + off < 32) incl += dn;
  }
  double after = __shfl_down_sync(0xffffffffu, incl, 1);      // sum over the lanes behind this one
  if (lane == 31) after = 0.0;
  float gdn = 0.f;
  {
    double run = after;
    for (int s = s1 - 1; s >= s0; --s) {
      const float gw = sg[s], T = sT[s], f = sf[s];
      const float galpha = gw * T - (float)(run / (double)f);
      run += (double)gw * (double)(sa[s] * T);
      const float sraw = p.raw[(ray * S + s) * 4 + 3];
      const float sigma = fmaxf(sraw, 0.0f);
      const float dz = (p.mip || s < S - 1) ? (sz[s + 1] - sz[s]) : 1e10f;
      const float ex = 1.0f - sa[s];                          // exp(-sigma delta)
      p.grad_raw[(ray * S + s) * 4 + 3] = sraw > 0.0f ? galpha * ex * (dz * dnorm) : 0.0f;
      gdn += (galpha * ex * sigma) * dz;
    }
  }
  if (p.grad_rays_d) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) gdn += __shfl_xor_sync(0xffffffffu, gdn, o);
    if (lane == 0) {
      const float inv = dnorm > 0.f ? 1.0f / dnorm : 0.f;
      p.grad_rays_d[ray * 3] = gdn * d0 * inv; p.grad_rays_d[ray * 3 + 1] = gdn * d1 * inv; p.grad_rays_d[ray * 3 + 2] = gdn * d2 * inv;
    }
  }
}
void launch_composite_backward(const CompositeBwdParams& p, cudaStream_t s) {
  if (p.n == 0) return;
  const size_t smem = kBwdWarps * kBwdSmem * sizeof(float);
  composite_backward_kernel<<<(unsigned)((p.n + kBwdWarps - 1) / kBwdWarps), kBwdWarps * 32, smem, s>>>(p);
}

}  // namespace nmr
