// Device helpers shared by the sampling kernels and the fused tcgen05 MLP kernel: conical-frustum
// Gaussians (render_utils.py:326-402) and the three positional encodings (embedding.py:35-46, 66-84),
// column by column in the reference's output order.
#pragma once
#include "nmr_internal.cuh"

namespace nmr {

// ------------------------------------------------------------------------------------------
// conical frustum -> Gaussian (stable form), lifted to a diagonal covariance
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void frustum_tstats(float t0, float t1, float radius, float* t_mean, float* t_var,
                                               float* r_var) {
  float mu = (t0 + t1) / 2.0f, hw = (t1 - t0) / 2.0f;
  float mu2 = mu * mu, hw2 = hw * hw, hw4 = hw2 * hw2;
  float denom = fmaxf(kEps32, 3.0f * mu2 + hw2);
  *t_mean = mu + (2.0f * mu * hw2) / denom;
  *t_var = hw2 / 3.0f - 0.26666668f * ((hw4 * (12.0f * mu2 - hw2)) / (denom * denom));
  *r_var = (radius * radius) * (mu2 / 4.0f + 0.41666666f * hw2 - 0.26666668f * hw4 / denom);
}

__device__ __forceinline__ void frustum_gaussian(const float* r, float t0, float t1, float var_scale,
                                                 float mean[3], float var[3]) {
  float t_mean, t_var, r_var;
  frustum_tstats(t0, t1, r[11], &t_mean, &t_var, &r_var);
  float dmag = fmaxf(1e-10f, r[3] * r[3] + r[4] * r[4] + r[5] * r[5]);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float d = r[3 + c];
    mean[c] = d * t_mean + r[c];
    float douter = d * d;
    float v = t_var * douter + r_var * (1.0f - douter / dmag);
    var[c] = (var_scale > 0.0f) ? var_scale * v : v;
  }
}

// ------------------------------------------------------------------------------------------
// encodings.  Column j of each encoder, exactly in the reference's output order.
// ------------------------------------------------------------------------------------------
// mip xyz (IPE, 6F columns): [sin block | sin(x + pi/2) block], each frequency-major / xyz-minor,
// damped by exp(-0.5 * var * 4^k)   (embedding.py:66-79)
__device__ __forceinline__ float ipe_col(const float m[3], const float v[3], int F, int j) {
  int half = 3 * F;
  int jj = (j < half) ? j : j - half;
  int k = jj / 3, c = jj - 3 * k;
  float sc = (float)(1u << k);
  float x = m[c] * sc;
  if (j >= half) x = x + kHalfPi;
  float y = v[c] * (sc * sc);
  return expf(-0.5f * y) * sinf(x);
}
// mip dirs (PE, 6F+3 columns): [sin block | sin(x+pi/2) block | x]   (embedding.py:80-84)
__device__ __forceinline__ float pe_col(const float x3[3], int F, int j) {
  int half = 3 * F;
  if (j >= 2 * half) return x3[j - 2 * half];
  int jj = (j < half) ? j : j - half;
  int k = jj / 3, c = jj - 3 * k;
  float x = x3[c] * (float)(1u << k);
  if (j >= half) x = x + kHalfPi;
  return sinf(x);
}
// Fourier (6F+3 columns): [x, sin(f0 x), cos(f0 x), sin(f1 x), ...]   (embedding.py:35-46)
__device__ __forceinline__ float fourier_col(const float x3[3], int F, int j) {
  if (j < 3) return x3[j];
  int jj = j - 3;
  int k = jj / 6, r = jj - 6 * k;
  float f = (float)(1u << k);
  return (r < 3) ? sinf(f * x3[r]) : cosf(f * x3[r - 3]);
}

}  // namespace nmr
