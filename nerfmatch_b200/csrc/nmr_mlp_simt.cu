// fp32 SIMT GEMM used by the reference-precision MLP path (NMR_PREC_FP32).
//
// The north star asks for composited features / depth within rel 1e-3 of the reference IN FP32;
// tcgen05 has no fp32 kind, so parity runs use this plain FFMA path (fp32 operands, fp32
// accumulate, ascending-k order) and the bf16 tcgen05 kernel (nmr_mlp_tc.cu) is the fast path.
// One launch evaluates  C = act(A1 * W1^T + A2 * W2^T + bias)  - the optional second segment
// implements the skip connection (models/nerf.py:105-106) and the [feature | dirs | app] input of
// views_linears (models/nerf.py:111-114) without materialising the concatenation.
#include "nmr_internal.cuh"

namespace nmr {

constexpr int BM = 128, BN = 128, BK = 16, TM = 8, TN = 8;
constexpr int kThreads = (BM / TM) * (BN / TN);  // 256

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == 1) return fmaxf(v, 0.0f);
  if (act == 2) return 1.0f / (1.0f + expf(-v));
  return v;
}

__global__ void __launch_bounds__(kThreads) sgemm_kernel(GemmParams p) {
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Ws[2][BK][BN + 4];
  const int tid = threadIdx.x;
  const int64_t row0 = (int64_t)blockIdx.x * BM;
  const int col0 = blockIdx.y * BN;
  const int tx = tid % (BN / TN), ty = tid / (BN / TN);

  const int tiles1 = p.K1 / BK, tiles2 = p.A2 ? p.K2 / BK : 0;
  const int ntiles = tiles1 + tiles2;

  // global -> register staging: each thread moves 2 float4 of A and 2 of W per k-tile
  const int lrow = tid / 4, lk = (tid % 4) * 4;
  float4 ra[2], rw[2];
  auto load_tile = [&](int t) {
    const float* A; const float* W; int lda, ldw, k0;
    if (t < tiles1) { A = p.A1; lda = p.lda1; W = p.W1; ldw = p.K1; k0 = t * BK; }
    else { A = p.A2; lda = p.lda2; W = p.W2; ldw = p.K2; k0 = (t - tiles1) * BK; }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      int64_t r = row0 + lrow + 64 * i;
      ra[i] = (r < p.M) ? *reinterpret_cast<const float4*>(A + r * lda + k0 + lk) : make_float4(0.f, 0.f, 0.f, 0.f);
      int c = col0 + lrow + 64 * i;
      rw[i] = (c < p.N) ? *reinterpret_cast<const float4*>(W + (int64_t)c * ldw + k0 + lk) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto store_tile = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      int r = lrow + 64 * i;
      As[buf][lk + 0][r] = ra[i].x; As[buf][lk + 1][r] = ra[i].y; As[buf][lk + 2][r] = ra[i].z; As[buf][lk + 3][r] = ra[i].w;
      Ws[buf][lk + 0][r] = rw[i].x; Ws[buf][lk + 1][r] = rw[i].y; Ws[buf][lk + 2][r] = rw[i].z; Ws[buf][lk + 3][r] = rw[i].w;
    }
  };

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  load_tile(0);
  store_tile(0);
  __syncthreads();
  for (int t = 0; t < ntiles; ++t) {
    const int buf = t & 1;
    if (t + 1 < ntiles) load_tile(t + 1);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TM], w[TN];
      float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      float4 w0 = *reinterpret_cast<const float4*>(&Ws[buf][k][tx * 4]);
      float4 w1 = *reinterpret_cast<const float4*>(&Ws[buf][k][64 + tx * 4]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      w[0] = w0.x; w[1] = w0.y; w[2] = w0.z; w[3] = w0.w; w[4] = w1.x; w[5] = w1.y; w[6] = w1.z; w[7] = w1.w;
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    if (t + 1 < ntiles) {
      store_tile(buf ^ 1);
      __syncthreads();
    }
  }

  // epilogue: thread owns rows {ty*4+i, 64+ty*4+i} x cols {tx*4+j, 64+tx*4+j}
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int64_t r = row0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (r >= p.M) continue;
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      int c = col0 + jh * 64 + tx * 4;
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int cc = c + j;
        float b = (p.bias && cc < p.N) ? p.bias[cc] : 0.f;
        float x = acc[i][jh * 4 + j] + b;
        if (cc < p.N) {
          if (p.r1a) x = fmaf(p.r1a[r * p.ldr1], p.r1b[cc], x);
          if (p.add) x += p.add[r * p.ldadd + cc];
        }
        x = apply_act(x, p.act);
        if (p.mask && cc < p.N && !(p.mask[r * p.ldm + cc] > 0.f)) x = 0.f;
        if (p.mask_bits && cc < p.N && !((p.mask_bits[mask_word_index(p.mask_row0 + r, p.mask_step, cc >> 5)] >> (cc & 31)) & 1u)) x = 0.f;
        v[j] = x;
      }
      float* out = p.C + r * p.ldc + c;
      if (c + 3 < p.N && (p.ldc % 4 == 0)) {
        *reinterpret_cast<float4*>(out) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (c + j < p.N) out[j] = v[j];
      }
    }
  }
}

// N <= 4 heads (alpha_linear, rgb_linear): one warp per row, lanes split K.
__global__ void thin_gemm_kernel(GemmParams p) {
  const int lane = threadIdx.x & 31;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= p.M) return;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  const float* a = p.A1 + row * p.lda1;
  for (int k = lane * 4; k < p.K1; k += 128) {
    float4 x = *reinterpret_cast<const float4*>(a + k);
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      if (n < p.N) {
        float4 w = *reinterpret_cast<const float4*>(p.W1 + n * p.K1 + k);
        acc[n] = fmaf(x.x, w.x, fmaf(x.y, w.y, fmaf(x.z, w.z, fmaf(x.w, w.w, acc[n]))));
      }
    }
  }
#pragma unroll
  for (int n = 0; n < 4; ++n)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[n] += __shfl_xor_sync(0xffffffffu, acc[n], o);
  if (lane == 0) {
    for (int n = 0; n < p.N; ++n) p.C[row * p.ldc + n] = apply_act(acc[n] + (p.bias ? p.bias[n] : 0.f), p.act);
  }
}

void launch_sgemm(const GemmParams& p, cudaStream_t s) {
  if (p.M == 0) return;
  if (p.N <= 4 && p.A2 == nullptr) {
    int64_t threads = p.M * 32;
    thin_gemm_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(p);
    return;
  }
  dim3 grid((unsigned)((p.M + BM - 1) / BM), (unsigned)((p.N + BN - 1) / BN));
  sgemm_kernel<<<grid, kThreads, 0, s>>>(p);
}

}  // namespace nmr
