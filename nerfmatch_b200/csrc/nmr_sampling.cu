// Ray generation, coarse stratified sampling, conical-frustum Gaussians and positional encodings.
// Compiled with -fmad=false so the op order matches the reference's un-fused fp32 arithmetic;
// the kernels are HBM / SFU bound, not FMA bound.
//
// Reference behaviour restated here (paths relative to the nv-dvl/nerfmatch tree):
//   nerfmatch/nerf/render_utils.py:23-41, 56-104   ray directions, radii, [ds/2::ds] subsampling
//   nerfmatch/nerf/scene_utils.py:101-120          far = forward intersection with the unit sphere
//   nerfmatch/nerf/render_utils.py:107-129, 405-450 coarse sampling (classic / mip)
//   nerfmatch/nerf/render_utils.py:326-402         conical frustum -> diagonal Gaussian
//   nerfmatch/nerf/embedding.py:35-46, 66-84       Fourier / IPE / PE encodings
#include "nmr_internal.cuh"
#include <algorithm>

#include "nmr_encode.cuh"

namespace nmr {

// ------------------------------------------------------------------------------------------
// ray generation
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void pixel_viewdir(const RayGenParams& p, int x, int y, float v[3]) {
  // dirs = [x, y, 1] @ Kinv^T ; d = dirs @ R^T ; viewdir = d / ||d||
  float fx = (float)x, fy = (float)y;
  float c[3], d[3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
    c[i] = fmaf(fx, p.kinv[3 * i + 0], fmaf(fy, p.kinv[3 * i + 1], p.kinv[3 * i + 2]));
#pragma unroll
  for (int i = 0; i < 3; ++i)
    d[i] = fmaf(c[0], p.rot[3 * i + 0], fmaf(c[1], p.rot[3 * i + 1], c[2] * p.rot[3 * i + 2]));
  float nrm = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
  v[0] = d[0] / nrm; v[1] = d[1] / nrm; v[2] = d[2] / nrm;
}

__global__ void raygen_kernel(RayGenParams p) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= p.h * p.w) return;
  int iy = idx / p.w, ix = idx - iy * p.w;
  int x = p.ds / 2 + ix * p.ds, y = p.ds / 2 + iy * p.ds;
  float v[3];
  pixel_viewdir(p, x, y, v);
  // rays_intersect_sphere(o, viewdirs, r=1)
  float odotd = p.org[0] * v[0] + p.org[1] * v[1] + p.org[2] * v[2];
  float dn = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
  float on = p.org[0] * p.org[0] + p.org[1] * p.org[1] + p.org[2] * p.org[2];
  float det = odotd * odotd + (1.0f - on) * dn;
  if (!(det >= 0.0f)) atomicExch(p.miss_flag, 1);
  float far = (sqrtf(det) - odotd) / dn;
  int ncols = p.mip ? 12 : 11;
  float* r = p.rays + (int64_t)idx * ncols;
  r[0] = p.org[0]; r[1] = p.org[1]; r[2] = p.org[2];
  r[3] = v[0]; r[4] = v[1]; r[5] = v[2];   // rays_d := viewdirs (render_utils.py:59)
  r[6] = 0.01f; r[7] = far;
  r[8] = v[0]; r[9] = v[1]; r[10] = v[2];
  if (p.mip) {
    // radii: ||d[y]-d[y+1]|| * 2/sqrt(12); the last image row re-uses the pair (H-3,H-2)
    int ya = (y < p.H - 1) ? y : p.H - 3;
    float a[3], b[3];
    if (ya == y) { a[0] = v[0]; a[1] = v[1]; a[2] = v[2]; } else pixel_viewdir(p, x, ya, a);
    pixel_viewdir(p, x, ya + 1, b);
    float e0 = a[0] - b[0], e1 = a[1] - b[1], e2 = a[2] - b[2];
    float dx = sqrtf(e0 * e0 + e1 * e1 + e2 * e2);
    r[11] = dx * 2.0f / 3.4641016151377544f;
  }
}

// The reference tests rays_intersect_sphere on ALL H x W pixels before it subsamples (render_utils.py:60-68): with the
// camera outside the unit sphere a pixel that the [ds/2::ds] grid skips can trigger the far = 1 fallback.  Launched only
// in that case (inside the sphere no ray can miss).
__global__ void raygen_miss_kernel(RayGenParams p) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= p.H * p.W) return;
  int y = idx / p.W, x = idx - y * p.W;
  float v[3];
  pixel_viewdir(p, x, y, v);
  float odotd = p.org[0] * v[0] + p.org[1] * v[1] + p.org[2] * v[2];
  float dn = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
  float on = p.org[0] * p.org[0] + p.org[1] * p.org[1] + p.org[2] * p.org[2];
  float det = odotd * odotd + (1.0f - on) * dn;
  if (!(det >= 0.0f)) atomicExch(p.miss_flag, 1);
}

__global__ void far_fallback_kernel(float* rays, int64_t n, int ncols, const int* miss_flag) {
  if (*miss_flag == 0) return;   // render_utils.py:62-68: any miss -> far = 1 for the whole view
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) rays[i * ncols + 7] = 1.0f;
}

void launch_raygen(const RayGenParams& p, cudaStream_t s) {
  int n = p.h * p.w;
  raygen_kernel<<<(n + 255) / 256, 256, 0, s>>>(p);
}
// ---- t_to_s (render_utils.py:618-636) over a whole call's z_fine: s = (g(t) - g(near)) / (g(far) - g(near)), g(x) = 1 / (x + 1e-6),
// near / far = min / max of ALL values (z_vals.min() / .max() over the batch).  g adds its epsilon IN PLACE (:634) and near goes
// through it twice, so the denominator's g(near) is 1 / (near + 1e-6 + 1e-6) - reproduced here -------------------------------------------
__device__ __forceinline__ uint32_t f2ord(float f) { uint32_t u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__device__ __forceinline__ float ord2f(uint32_t u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }
__global__ void minmax_kernel(const float* __restrict__ z, int64_t count, uint32_t* scratch) {
  uint32_t lo = 0xffffffffu, hi = 0u;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t v = f2ord(z[i]);
    lo = min(lo, v); hi = max(hi, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o)); hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o)); }
  if ((threadIdx.x & 31) == 0) { atomicMin(scratch, lo); atomicMax(scratch + 1, hi); }
}
__global__ void t_to_s_kernel(float* z, int64_t count, const uint32_t* scratch) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const float eps = 1e-6f;
  const float near1 = ord2f(scratch[0]) + eps, near2 = near1 + eps;
  z[i] = (1.0f / (z[i] + eps) - 1.0f / near1) / (1.0f / (ord2f(scratch[1]) + eps) - 1.0f / near2);
}
__global__ void minmax_init_kernel(uint32_t* scratch) { scratch[0] = 0xffffffffu; scratch[1] = 0u; }
void launch_t_to_s(float* z, int64_t count, uint32_t* scratch, cudaStream_t s) {
  if (count == 0) return;
  minmax_init_kernel<<<1, 1, 0, s>>>(scratch);
  minmax_kernel<<<(unsigned)std::min<int64_t>((count + 255) / 256, 1184), 256, 0, s>>>(z, count, scratch);
  t_to_s_kernel<<<(unsigned)((count + 255) / 256), 256, 0, s>>>(z, count, scratch);
}

void launch_raygen_miss(const RayGenParams& p, cudaStream_t s) {
  int n = p.H * p.W;
  raygen_miss_kernel<<<(n + 255) / 256, 256, 0, s>>>(p);
}
void launch_far_fallback(float* rays, int64_t n, int ncols, const int* miss_flag, cudaStream_t s) {
  far_fallback_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(rays, n, ncols, miss_flag);
}

// ---- the reference's ray helper functions as separate ops (render_utils.py:23-41, 81-104) -------------------
__global__ void ray_dirs_kernel(RayOpsParams p) {          // get_ray_dirs: [x, y, 1] @ inv(K)^T (* (1,-1,-1))
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= p.H * p.W) return;
  const float fx = (float)(idx % p.W), fy = (float)(idx / p.W);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    float c = fmaf(fx, p.kinv[3 * i + 0], fmaf(fy, p.kinv[3 * i + 1], p.kinv[3 * i + 2]));
    if (p.flip && i > 0) c = -c;
    p.out[(int64_t)idx * 3 + i] = c;
  }
}
__global__ void rays_c2w_kernel(RayOpsParams p) {          // get_rays_c2w: d = dirs @ R^T, o = t, viewdirs = d / |d|
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= p.n) return;
  const float* c = p.dirs + idx * 3;
  float d[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) d[i] = fmaf(c[0], p.rot[3 * i + 0], fmaf(c[1], p.rot[3 * i + 1], c[2] * p.rot[3 * i + 2]));
  const float nrm = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
#pragma unroll
  for (int i = 0; i < 3; ++i) { p.o[idx * 3 + i] = p.org[i]; p.d[idx * 3 + i] = d[i]; p.v[idx * 3 + i] = d[i] / nrm; }
}
__global__ void prepare_rays_kernel(RayOpsParams p) {      // prepare_rays_data: cat(o, d, near, far[, viewdirs][, radii])
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= p.n) return;
  float* r = p.out + idx * p.ncols;
#pragma unroll
  for (int i = 0; i < 3; ++i) { r[i] = p.o[idx * 3 + i]; r[3 + i] = p.d[idx * 3 + i]; }
  r[6] = p.near ? p.near[idx] : p.near_s;
  r[7] = p.far ? p.far[idx] : p.far_s;
  int col = 8;
  if (p.v) { r[8] = p.v[idx * 3]; r[9] = p.v[idx * 3 + 1]; r[10] = p.v[idx * 3 + 2]; col = 11; }
  if (p.comp_radii) {                                      // |d[y] - d[y+1]| * 2 / sqrt(12); last row: the pair (H-3, H-2)
    const int y = (int)(idx / p.W), x = (int)(idx % p.W);
    const int ya = (y < p.H - 1) ? y : p.H - 3;
    const float* a = p.d + ((int64_t)ya * p.W + x) * 3;
    const float* b = a + (int64_t)p.W * 3;
    const float e0 = a[0] - b[0], e1 = a[1] - b[1], e2 = a[2] - b[2];
    r[col] = sqrtf(e0 * e0 + e1 * e1 + e2 * e2) * 2.0f / 3.4641016151377544f;
  }
}
void launch_ray_op(int op, const RayOpsParams& p, cudaStream_t s) {
  const int64_t n = op == 0 ? (int64_t)p.H * p.W : p.n;
  if (n == 0) return;
  const unsigned grid = (unsigned)((n + 255) / 256);
  if (op == 0) ray_dirs_kernel<<<grid, 256, 0, s>>>(p);
  else if (op == 1) rays_c2w_kernel<<<grid, 256, 0, s>>>(p);
  else prepare_rays_kernel<<<grid, 256, 0, s>>>(p);
}

// ------------------------------------------------------------------------------------------
// coarse sampling
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float lerp_nf(float near, float far, float tau, int use_disp) {
  if (use_disp) return 1.0f / (1.0f / near * (1.0f - tau) + 1.0f / far * tau);
  return near * (1.0f - tau) + far * tau;
}

__global__ void sample_coarse_kernel(SampleParams p) {
  const int nz = p.mip ? p.S + 1 : p.S;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n * nz) return;
  int64_t ray = i / nz;
  int j = (int)(i - ray * nz);
  const float* r = p.rays + ray * p.ncols;
  float near = r[6], far = r[7];
  const int ud = p.mip ? 0 : p.use_disp;
  float t = lerp_nf(near, far, torch_linspace(1.0f, nz, j), ud);
  if ((p.mip || p.perturb) && (p.t_rand != nullptr || p.rng_mode == NMR_RNG_PHILOX)) {
    // stratified jitter inside half-bins: mip always (render_utils.py:440-445), classic in training mode (:118-126)
    float tp = (j > 0) ? lerp_nf(near, far, torch_linspace(1.0f, nz, j - 1), ud) : t;
    float tn = (j < nz - 1) ? lerp_nf(near, far, torch_linspace(1.0f, nz, j + 1), ud) : t;
    float upper = (j < nz - 1) ? 0.5f * (tn + t) : t;
    float lower = (j > 0) ? 0.5f * (t + tp) : t;
    float u = p.t_rand ? p.t_rand[i] : philox_uniform(p.seed, (uint64_t)(p.ray0 + ray) * 1024u + j, 0u);
    t = lower + (upper - lower) * u;
  }
  p.z[i] = t;
}

void launch_sample_coarse(const SampleParams& p, cudaStream_t s) {
  int64_t tot = p.n * (p.mip ? p.S + 1 : p.S);
  if (tot == 0) return;
  sample_coarse_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(p);
}

__global__ void points_kernel(PointsParams p) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n * p.S) return;
  int64_t ray = i / p.S;
  int s = (int)(i - ray * p.S);
  const float* r = p.rays + ray * p.ncols;
  float m[3], v[3];
  if (p.mip) {
    const float* z = p.z + ray * (p.S + 1);
    frustum_gaussian(r, z[s], z[s + 1], p.var_scale, m, v);
    if (p.var) { p.var[i * 3 + 0] = v[0]; p.var[i * 3 + 1] = v[1]; p.var[i * 3 + 2] = v[2]; }
  } else {
    float z = p.z[i];
#pragma unroll
    for (int c = 0; c < 3; ++c) m[c] = r[c] + r[3 + c] * z;
  }
  if (p.pts) { p.pts[i * 3 + 0] = m[0]; p.pts[i * 3 + 1] = m[1]; p.pts[i * 3 + 2] = m[2]; }
}

void launch_points(const PointsParams& p, cudaStream_t s) {
  int64_t tot = p.n * p.S;
  if (tot == 0) return;
  points_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(p);
}

__global__ void encode_kernel(EncodeParams p) {
  const Dims& d = p.d;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n * p.S * d.kx) return;
  int64_t pt = i / d.kx;
  int j = (int)(i - pt * d.kx);
  int64_t ray = pt / p.S;
  int s = (int)(pt - ray * p.S);
  const float* r = p.rays + ray * p.ncols;
  float out = 0.0f;
  if (j < d.xyz) {
    float m[3], v[3];
    if (d.mip) {
      const float* z = p.z + ray * (p.S + 1);
      frustum_gaussian(r, z[s], z[s + 1], p.var_scale, m, v);
      out = ipe_col(m, v, d.fx, j);
    } else {
      float z = p.z[pt];
#pragma unroll
      for (int c = 0; c < 3; ++c) m[c] = r[c] + r[3 + c] * z;
      out = fourier_col(m, d.fx, j);
    }
  } else if (j >= d.xyzp) {
    int jj = j - d.xyzp;
    if (jj < d.dirs) {
      const float* vd = (p.ncols >= 11) ? r + 8 : r + 3;   // renderer.py:188
      float v3[3] = {vd[0], vd[1], vd[2]};
      out = d.mip ? pe_col(v3, d.fd, jj) : fourier_col(v3, d.fd, jj);
    } else if (jj < d.dirs + d.app) {
      int64_t id = p.ray_id ? p.ray_id[ray] : 1;           // renderer.py:298-299
      id = id < 0 ? 0 : (id >= p.num_frames ? p.num_frames - 1 : id);
      out = p.app_table[id * d.app + (jj - d.dirs)];
    }
  }
  p.X[i] = out;
}

void launch_encode(const EncodeParams& p, cudaStream_t s) {
  int64_t tot = p.n * p.S * p.d.kx;
  if (tot == 0) return;
  encode_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(p);
}

__global__ void encode_stage_kernel(int mode, int F, int dim, const float* x, const float* var, int64_t P,
                                    float* out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P * dim) return;
  int64_t pt = i / dim;
  int j = (int)(i - pt * dim);
  float m[3] = {x[pt * 3], x[pt * 3 + 1], x[pt * 3 + 2]};
  float r;
  if (mode == 0) {
    float v[3] = {var[pt * 3], var[pt * 3 + 1], var[pt * 3 + 2]};
    r = ipe_col(m, v, F, j);
  } else if (mode == 1) {
    r = pe_col(m, F, j);
  } else {
    r = fourier_col(m, F, j);
  }
  out[i] = r;
}

void launch_encode_stage(int kind, const Dims& d, const float* x, const float* var, int64_t P, float* out,
                         cudaStream_t s) {
  int F = kind == 0 ? d.fx : d.fd;
  int mode, dim;
  if (d.mip && kind == 0 && var != nullptr) { mode = 0; dim = 6 * F; }
  else if (d.mip) { mode = 1; dim = 6 * F + 3; }
  else { mode = 2; dim = 6 * F + 3; }
  int64_t tot = P * dim;
  if (tot == 0) return;
  encode_stage_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(mode, F, dim, x, var, P, out);
}

__global__ void repack_input_kernel(Dims d, const float* x, int64_t P, float* X) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P * d.kx) return;
  int64_t pt = i / d.kx;
  int j = (int)(i - pt * d.kx);
  float v = 0.0f;
  if (j < d.xyz) v = x[pt * d.in_dim + j];
  else if (j >= d.xyzp && j - d.xyzp < d.dirs + d.app) v = x[pt * d.in_dim + d.xyz + (j - d.xyzp)];
  X[i] = v;
}

void launch_repack_input(const Dims& d, const float* x, int64_t P, float* X, cudaStream_t s) {
  int64_t tot = P * d.kx;
  if (tot == 0) return;
  repack_input_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(d, x, P, X);
}

}  // namespace nmr
