// Micro-benchmark: issue rate of tcgen05.mma kind::f16 (bf16 x bf16 -> fp32), M = 128 per CTA, K = 16, on every SM.
// One thread per CTA issues `iters` MMAs back to back into one accumulator and commits once at the end:
// cycles / MMA is the tensor-core floor for that shape, without any data movement (operands are whatever is in
// shared memory / TMEM).
//   mode 0: A and B from shared memory (SS)     mode 1: A from TMEM (TS)
//   arg N : 64..256                              arg cg: 1 | 2 (CTA pairs, M = 256 per instruction)
//   arg commit_every: also commit to a second barrier every k MMAs (the ring-release pattern), 0 = never
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t mk_desc(uint32_t saddr) {   // K-major, SWIZZLE_64B, SBO 512 B
  uint32_t lo = ((saddr >> 4) & 0x3FFFu) | (1u << 16);
  uint32_t hi = (512u >> 4) | (1u << 14) | (4u << 29);
  uint64_t d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
  return d;
}
__device__ __forceinline__ bool try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}

template <int CG>
__global__ void __launch_bounds__(128, 1) k(int mode, int N, int iters, int commit_every, long long* cyc) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tslot;
  __shared__ __align__(8) uint64_t bar[2 + 16];
  const int warp = threadIdx.x >> 5;
  uint32_t rank = 0;
  if (CG == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[0])));
    for (int i = 1; i < 18; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    if (CG == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(512));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(512));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
  }
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + i;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tslot;
  if (threadIdx.x == 32 && rank == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)((128 * CG) >> 4) << 24);
    const uint32_t sa = smem_u32(smem), sb = smem_u32(smem) + 16384;
    long long t0 = clock64();
    if (commit_every >= 100) {
      // compile-time patterns of the issue thread's work per K=32 chunk (TS mode, cta_group::1 or 2):
      //   100: MMA MMA                 101: MMA MMA commit          102: MMA x4 commit
      //   103: wait fence MMA MMA commit   104: wait x3 fence MMA MMA commit   105: wait x3 fence MMA x4 commit
      const int pat = commit_every - 100;
      const uint64_t bd0 = mk_desc(sb);
      const uint32_t rb = smem_u32(&bar[14]);        // fresh barrier: waiting for parity 1 succeeds at once
      auto mma = [&](uint32_t ta, uint64_t bd, uint32_t acc) {
        if (CG == 2)
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                       ::"r"(tb), "r"(ta), "l"(bd), "r"(idesc), "r"(acc) : "memory");
        else
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                       ::"r"(tb), "r"(ta), "l"(bd), "r"(idesc), "r"(acc) : "memory");
      };
      auto commit = [&](uint32_t cb) {
        if (CG == 2)
          asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(cb), "h"((uint16_t)3) : "memory");
        else
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(cb) : "memory");
      };
      const int nm = (pat == 2 || pat == 5) ? 4 : 2;
      int stage = 0;
      for (int i = 0; i < iters; i += nm) {
        if (pat >= 3) {
          bool ok = try_wait(rb, 1);
          if (pat >= 4) { bool ok2 = try_wait(rb + 8, 1), ok3 = try_wait(rb + 16, 1); ok = ok && ok2 && ok3; }
          if (!ok) break;
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        const uint64_t bd = bd0 + (uint64_t)(stage * 256);
        mma(tb + 256 + stage * 16, bd, i > 0);
        mma(tb + 256 + stage * 16 + 8, bd + 2, 1u);
        if (nm == 4) { mma(tb + 256 + stage * 16 + 16, bd + 4, 1u); mma(tb + 256 + stage * 16 + 24, bd + 6, 1u); }
        if (pat >= 1) commit(smem_u32(&bar[1 + stage]));
        if (++stage == 6) stage = 0;
      }
    } else
    for (int i = 0; i < iters; ++i) {
      const uint32_t ko = (i & 1) * 2;                       // alternate the two K=16 halves of a 32-wide block
      const uint32_t st = ((i >> 1) % 3) * 8192;             // rotate through a few B stages
      const uint64_t bd = mk_desc(sb + st) + ko;
      const uint32_t acc = i > 0;
      if (mode == 0) {
        const uint64_t ad = mk_desc(sa + ((i >> 1) & 1) * 8192) + ko;
        if (CG == 2)
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                       ::"r"(tb), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
        else
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                       ::"r"(tb), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
      } else {
        const uint32_t ta = tb + 256 + (i & 15) * 8;
        if (CG == 2)
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                       ::"r"(tb), "r"(ta), "l"(bd), "r"(idesc), "r"(acc) : "memory");
        else
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                       ::"r"(tb), "r"(ta), "l"(bd), "r"(idesc), "r"(acc) : "memory");
      }
      const int ce = commit_every < 0 ? -commit_every : commit_every;
      if (ce && (i + 1) % ce == 0) {
        const uint32_t cb = smem_u32(&bar[1 + (commit_every < 0 ? ((i + 1) / ce) % 12 : 0)]);   // < 0: rotate over 12 barriers
        if (CG == 2)
          asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(cb), "h"((uint16_t)3) : "memory");
        else
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(cb) : "memory");
      }
    }
    if (CG == 2)
      asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&bar[0])), "h"((uint16_t)1) : "memory");
    else
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[0])) : "memory");
    long long t1 = clock64();
    while (!try_wait(smem_u32(&bar[0]), 0)) {
      if (clock64() - t1 > 2000000000LL) break;
    }
    long long t2 = clock64();
    cyc[blockIdx.x * 2] = t2 - t0;
    cyc[blockIdx.x * 2 + 1] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (CG == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512));
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512));
  }
}

template <int CG>
static void run(int mode, int N, int iters, int commit_every, int grid) {
  long long* d;
  cudaMalloc(&d, grid * 2 * sizeof(long long));
  cudaMemset(d, 0, grid * 2 * sizeof(long long));
  const int smem = 64 * 1024;
  cudaFuncSetAttribute(k<CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int rep = 0; rep < 2; ++rep) {
    if (CG == 2) {
      cudaLaunchConfig_t cfg{};
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.gridDim = dim3(grid & ~1); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem; cfg.attrs = at; cfg.numAttrs = 1;
      cudaLaunchKernelEx(&cfg, k<CG>, mode, N, iters, commit_every, d);
    } else {
      k<CG><<<grid, 128, smem>>>(mode, N, iters, commit_every, d);
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); exit(1); }
  }
  long long h[2 * 160];
  cudaMemcpy(h, d, grid * 2 * sizeof(long long), cudaMemcpyDeviceToHost);
  double tot = 0, iss = 0; int cnt = 0;
  for (int i = 0; i < grid; ++i) if (h[2 * i] > 0) { tot += h[2 * i]; iss += h[2 * i + 1]; ++cnt; }
  printf("cg=%d mode=%s N=%3d iters=%d commit_every=%d: %.1f cycles/MMA (issue loop alone %.1f), %d issuing CTAs -> %.0f flop/cycle/SM\n", CG,
         mode ? "TS" : "SS", N, iters, commit_every, tot / cnt / iters, iss / cnt / iters, cnt, 2.0 * 128 * N * 16 * iters / (tot / cnt));
  cudaFree(d);
}

int main(int argc, char** argv) {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int iters = 2048;
  const bool full = argc > 1;
  for (int cg = 1; cg <= 2; ++cg)
    for (int pat = 100; pat <= 105; ++pat) { if (cg == 1) run<1>(1, 256, iters, pat, sms); else run<2>(1, 256, iters, pat, 2 * 64); }
  for (int cg = 1; cg <= 2; ++cg)
    for (int mode = 0; mode < 2; ++mode)
      for (int N : {256, 128, 160, 64}) {
        if (!full && N != 256) continue;
        for (int ce : {0, 1, 2, 4, 8, 16, -2, -4}) {
          if (cg == 1) run<1>(mode, N, iters, ce, sms); else run<2>(mode, N, iters, ce, 2 * 64);
        }
      }
  return 0;
}
