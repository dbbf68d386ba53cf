// Micro-benchmark: tcgen05.ld (TMEM -> registers) throughput per SM.
// nwarps warps (4 or 8; warp w reads lane quarter w%4) each issue `iters` loads of shape 32x32b.xN.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_ld_bw tmem_ld_bw.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int N> __device__ __forceinline__ uint32_t ld(uint32_t taddr);
template <> __device__ __forceinline__ uint32_t ld<32>(uint32_t taddr) {
  uint32_t v[32];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
    : "=r"(v[0]),"=r"(v[1]),"=r"(v[2]),"=r"(v[3]),"=r"(v[4]),"=r"(v[5]),"=r"(v[6]),"=r"(v[7]),"=r"(v[8]),"=r"(v[9]),"=r"(v[10]),"=r"(v[11]),"=r"(v[12]),"=r"(v[13]),"=r"(v[14]),"=r"(v[15]),
      "=r"(v[16]),"=r"(v[17]),"=r"(v[18]),"=r"(v[19]),"=r"(v[20]),"=r"(v[21]),"=r"(v[22]),"=r"(v[23]),"=r"(v[24]),"=r"(v[25]),"=r"(v[26]),"=r"(v[27]),"=r"(v[28]),"=r"(v[29]),"=r"(v[30]),"=r"(v[31])
    : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  uint32_t a = 0;
#pragma unroll
  for (int i = 0; i < 32; ++i) a ^= v[i];
  return a;
}
template <> __device__ __forceinline__ uint32_t ld<8>(uint32_t taddr) {
  uint32_t v[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
    : "=r"(v[0]),"=r"(v[1]),"=r"(v[2]),"=r"(v[3]),"=r"(v[4]),"=r"(v[5]),"=r"(v[6]),"=r"(v[7]) : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  uint32_t a = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) a ^= v[i];
  return a;
}
// two x32 loads in flight before one wait
__device__ __forceinline__ uint32_t ld2x32(uint32_t t0, uint32_t t1) {
  uint32_t v[64];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
    : "=r"(v[0]),"=r"(v[1]),"=r"(v[2]),"=r"(v[3]),"=r"(v[4]),"=r"(v[5]),"=r"(v[6]),"=r"(v[7]),"=r"(v[8]),"=r"(v[9]),"=r"(v[10]),"=r"(v[11]),"=r"(v[12]),"=r"(v[13]),"=r"(v[14]),"=r"(v[15]),
      "=r"(v[16]),"=r"(v[17]),"=r"(v[18]),"=r"(v[19]),"=r"(v[20]),"=r"(v[21]),"=r"(v[22]),"=r"(v[23]),"=r"(v[24]),"=r"(v[25]),"=r"(v[26]),"=r"(v[27]),"=r"(v[28]),"=r"(v[29]),"=r"(v[30]),"=r"(v[31])
    : "r"(t0) : "memory");
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
    : "=r"(v[32]),"=r"(v[33]),"=r"(v[34]),"=r"(v[35]),"=r"(v[36]),"=r"(v[37]),"=r"(v[38]),"=r"(v[39]),"=r"(v[40]),"=r"(v[41]),"=r"(v[42]),"=r"(v[43]),"=r"(v[44]),"=r"(v[45]),"=r"(v[46]),"=r"(v[47]),
      "=r"(v[48]),"=r"(v[49]),"=r"(v[50]),"=r"(v[51]),"=r"(v[52]),"=r"(v[53]),"=r"(v[54]),"=r"(v[55]),"=r"(v[56]),"=r"(v[57]),"=r"(v[58]),"=r"(v[59]),"=r"(v[60]),"=r"(v[61]),"=r"(v[62]),"=r"(v[63])
    : "r"(t1) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  uint32_t a = 0;
#pragma unroll
  for (int i = 0; i < 64; ++i) a ^= v[i];
  return a;
}

__global__ void __launch_bounds__(256, 1) k(int mode, int iters, long long* cyc, uint32_t* sink) {
  __shared__ uint32_t tslot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tslot + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 256;
  uint32_t acc = 0;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    const uint32_t col = (i & 7) * 32;
    if (mode == 0) acc ^= ld<32>(tb + col);
    else if (mode == 1) { acc ^= ld<8>(tb + col); acc ^= ld<8>(tb + col + 8); acc ^= ld<8>(tb + col + 16); acc ^= ld<8>(tb + col + 24); }
    else acc ^= ld2x32(tb + (col & 0xff), tb + ((col + 32) & 0xff));
  }
  long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x % 32 == 0) cyc[blockIdx.x * 8 + warp] = t1 - t0;
  if (acc == 0x12345678) sink[0] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tslot), "r"(512));
}

int main() {
  long long* cyc; cudaMalloc(&cyc, 148 * 8 * sizeof(long long));
  uint32_t* sink; cudaMalloc(&sink, 4);
  const int iters = 2000;
  for (int nw : {4, 8}) for (int mode = 0; mode < 3; ++mode) {
    k<<<148, nw * 32>>>(mode, iters, cyc, sink);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[8]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double per = (double)h[0] / iters;
    double bytes_per_iter = (mode == 2 ? 2 : 1) * 32 * 32 * 4.0 * nw;   // all warps of the SM
    printf("warps=%d mode=%d (%s): %s  %.1f cycles per iteration per warp -> %.1f B/cycle/SM\n", nw, mode,
           mode == 0 ? "x32+wait" : mode == 1 ? "4 x (x8+wait)" : "2 x x32 then wait", cudaGetErrorString(e), per, bytes_per_iter / per);
  }
  return 0;
}
