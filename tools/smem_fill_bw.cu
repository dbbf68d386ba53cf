// Micro-benchmark: how fast can one SM fill shared memory from an L2-resident buffer?
//   mode 0: cp.async.bulk (1-D, UBLKCP), `chunk` bytes per copy, `stages` copies in flight
//   mode 1: cp.async (LDGSTS.128) by 128 producer threads, commit/wait groups
//   mode 2: ld.global.v4 + st.shared.v4 by 128 threads (8 loads in flight per thread)
//   mode 3: cp.async.bulk.tensor.2d (UTMALDG) with a 64B-swizzled [256 x 32 bf16] box
// One CTA per SM, every CTA streams the same 1.2 MB "weight" buffer `iters` times.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o smem_fill_bw smem_fill_bw.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
}

__global__ void __launch_bounds__(160, 1) fill_kernel(const uint8_t* __restrict__ src, int total_bytes, int chunk, int stages, int iters, int mode,
                                                       const __grid_constant__ CUtensorMap tmap, long long* cycles, int* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[32];
  const uint32_t sb = (smem_u32(smem) + 1023u) & ~1023u;
  uint8_t* sg = smem + (sb - smem_u32(smem));
  const int nchunks = total_bytes / chunk;
  if (threadIdx.x == 0) { for (int i = 0; i < 32; ++i) mbar_init(smem_u32(&bars[i]), 1); asm volatile("fence.mbarrier_init.release.cluster;"); }
  __syncthreads();
  long long t0 = clock64();
  int acc = 0;
  if (mode == 0 || mode == 3) {
    if (threadIdx.x == 0) {
      // producer + consumer in one thread: keep `stages` copies in flight
      int issued = 0, done = 0; const int total = nchunks * iters;
      uint32_t phase_bits = 0;
      while (done < total) {
        while (issued < total && issued - done < stages) {
          int st = issued % stages;
          uint32_t bar = smem_u32(&bars[st]);
          mbar_expect(bar, chunk);
          int c = issued % nchunks;
          if (mode == 0) {
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(sb + st * chunk), "l"(src + (size_t)c * chunk), "r"(chunk), "r"(bar) : "memory");
          } else {
            int rows = chunk / 64;
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         ::"r"(sb + st * chunk), "l"(&tmap), "r"(0), "r"(c * rows), "r"(bar) : "memory");
          }
          ++issued;
        }
        int st = done % stages;
        mbar_wait(smem_u32(&bars[st]), (phase_bits >> st) & 1);
        phase_bits ^= 1u << st;
        acc += sg[st * chunk + (done & 1023)];
        ++done;
      }
    }
  } else if (mode >= 4) {
    // mode 4 + P: P independent producer threads (one per warp), each keeping `stages` bulk copies in flight in its own
    // slice of the ring; chunk order rotated per CTA so the SMs do not all ask L2 for the same lines at the same time
    const int P = mode - 4 + 1;
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0 && w < P) {
      int issued = 0, done = 0; const int total = nchunks * iters / P;
      uint32_t phase_bits = 0;
      while (done < total) {
        while (issued < total && issued - done < stages) {
          int st = w * stages + issued % stages;
          uint32_t bar = smem_u32(&bars[st]);
          mbar_expect(bar, chunk);
          int c = (issued * P + w + blockIdx.x * 5) % nchunks;
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(sb + st * chunk), "l"(src + (size_t)c * chunk), "r"(chunk), "r"(bar) : "memory");
          ++issued;
        }
        int st = w * stages + done % stages;
        mbar_wait(smem_u32(&bars[st]), (phase_bits >> (done % stages)) & 1);
        phase_bits ^= 1u << (done % stages);
        ++done;
      }
    }
  } else if (mode == 1) {
    if (threadIdx.x < 128) {
      const int per_thread = chunk / 128 / 16;   // 16-byte copies per thread per chunk
      const int total = nchunks * iters;
      for (int i = 0; i < total + stages - 1; ++i) {
        if (i < total) {
          int st = i % stages, c = i % nchunks;
          for (int k = 0; k < per_thread; ++k) {
            int off = (k * 128 + threadIdx.x) * 16;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sb + st * chunk + off), "l"(src + (size_t)c * chunk + off) : "memory");
          }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        if (i >= stages - 1) {
          asm volatile("cp.async.wait_group %0;" ::"n"(5) : "memory");   // stages-1 = 5 groups may stay in flight (stages == 6)
          acc += sg[((i - (stages - 1)) % stages) * chunk + threadIdx.x * 16];
        }
      }
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
  } else {
    if (threadIdx.x < 128) {
      const int total = nchunks * iters;
      for (int i = 0; i < total; ++i) {
        int st = i % stages, c = i % nchunks;
        const uint4* g = reinterpret_cast<const uint4*>(src + (size_t)c * chunk);
        uint4* s = reinterpret_cast<uint4*>(sg + st * chunk);
        const int n16 = chunk / 16;
        for (int k0 = 0; k0 < n16; k0 += 128 * 8) {
          uint4 v[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) { int k = k0 + u * 128 + threadIdx.x; v[u] = (k < n16) ? __ldg(g + k) : make_uint4(0, 0, 0, 0); }
#pragma unroll
          for (int u = 0; u < 8; ++u) { int k = k0 + u * 128 + threadIdx.x; if (k < n16) s[k] = v[u]; }
        }
        acc += (int)s[threadIdx.x].x;
      }
    }
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (acc == 0x7fffffff) sink[0] = acc;
}

int main(int argc, char** argv) {
  int dev = 0; cudaSetDevice(dev);
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, dev);
  const int sms = prop.multiProcessorCount;
  const int total = 1184 * 1024;     // one net's chunk stream
  uint8_t* src; cudaMalloc(&src, total); cudaMemset(src, 1, total);
  long long* cyc; cudaMalloc(&cyc, sms * sizeof(long long));
  int* sink; cudaMalloc(&sink, 4);
  // tensor map over the buffer viewed as [total/64 rows][32 bf16], box 256 x 32, 64B swizzle
  CUtensorMap tmap;
  cuuint64_t gdim[2] = {32, (cuuint64_t)(total / 64)}; cuuint64_t gstr[1] = {64};
  cuuint32_t box[2] = {32, 256}; cuuint32_t estr[2] = {1, 1};
  CUresult cr = cuTensorMapEncodeTiled(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, src, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                       CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("tensor map encode: %d\n", (int)cr);
  const int smem = 6 * 16384 + 2048;
  cudaFuncSetAttribute(fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  struct Cfg { int mode, chunk, stages; const char* name; } cfgs[] = {
      {0, 16384, 3, "bulk1d 16K x3"}, {0, 16384, 6, "bulk1d 16K x6"}, {0, 4096, 24, "bulk1d 4K x24"}, {0, 2048, 48, "bulk1d 2K x48"},
      {0, 8192, 12, "bulk1d 8K x12"},
      {4, 16384, 6, "rot 1 thr 16K x6"}, {4, 8192, 12, "rot 1 thr 8K x12"}, {4, 32768, 3, "rot 1 thr 32K x3"}, {4, 32768, 2, "rot 1 thr 32K x2"},
      {5, 16384, 3, "rot 2 thr 16K x3"}, {5, 8192, 6, "rot 2 thr 8K x6"}, {7, 8192, 3, "rot 4 thr 8K x3"}, {7, 4096, 6, "rot 4 thr 4K x6"}, {7, 16384, 1, "rot 4 thr 16K x1"}, {1, 16384, 6, "LDGSTS 16K x6"}, {2, 16384, 6, "LDG+STS 16K"}, {3, 16384, 6, "tensor2d 16K x6"}, {3, 16384, 3, "tensor2d 16K x3"}};
  for (auto& c : cfgs) {
    if (c.mode == 3 && cr != CUDA_SUCCESS) continue;
    const int iters = 40;
    for (int rep = 0; rep < 2; ++rep) {
      cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
      cudaEventRecord(a);
      fill_kernel<<<sms, 160, smem>>>(src, total, c.chunk, c.stages > 16 && c.mode == 0 ? 16 : c.stages, iters, c.mode, tmap, cyc, sink);
      cudaEventRecord(b);
      cudaError_t e = cudaDeviceSynchronize();
      float ms = 0; cudaEventElapsedTime(&ms, a, b);
      std::vector<long long> h(sms); cudaMemcpy(h.data(), cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost);
      double avg = 0; for (auto v : h) avg += v; avg /= sms;
      double bytes = (double)total * iters;
      if (rep == 1) printf("%-18s: %s  %.3f ms  %.1f B/cycle/SM  aggregate %.2f TB/s\n", c.name, cudaGetErrorString(e), ms, bytes / avg, bytes * sms / (ms * 1e-3) / 1e12);
    }
  }
  return 0;
}
