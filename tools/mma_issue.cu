// Micro-benchmark: cost of ONE ITERATION of the MLP kernel's MMA-issue loop, in the form the kernel uses it: the whole
// warp runs the loop converged, one elected lane (elect.sync) issues.  No data movement: every barrier the loop probes is
// already complete, operands are whatever is in shared memory / TMEM.  The question it answers: is a [wait, fence,
// n x tcgen05.mma, commit] iteration bounded by the tensor core (n x 128 N/256 cycles) or by the issuing warp?
//   pattern  0: n MMAs                         1: n MMAs + commit (per-stage barrier)
//            2: try_wait x2 + fence + n MMAs + commit (the kernel's loop)
//            3: as 2, but stage / descriptors fully unrolled (compile-time ring position)
//   N = 256 | 128 | 64 | 32, n = MMAs per iteration (2 | 4), TS form (A from TMEM)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_issue mma_issue.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr) { return ((saddr >> 4) & 0x3FFFu) | (1u << 16); }
__device__ __forceinline__ uint64_t mk_desc(uint32_t lo) {
  constexpr uint32_t hi = (512u >> 4) | (1u << 14) | (4u << 29);
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
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

constexpr int kStages = 6;

template <int PAT, int NM>
__global__ void __launch_bounds__(128, 1) k(int N, int iters, long long* cyc) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tslot;
  __shared__ __align__(8) uint64_t bar[32];
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 32; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  for (int i = threadIdx.x; i < 120 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + i;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tslot;
  if (warp == 1) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t w_desc0 = desc_lo(smem_u32(smem));
    const uint32_t bars = smem_u32(&bar[0]);
    const uint32_t ready = bars + 8 * 24;          // never-armed barriers: waiting for parity 1 succeeds at once
    const uint32_t a_tmem = tb + 256;
    long long t0 = clock64();
    int stage = 0;
    uint32_t acc = 0;
    if (PAT == 3) {
      for (int i = 0; i < iters; i += kStages) {
#pragma unroll
        for (int st = 0; st < kStages; ++st) {
          const bool ok_h = try_wait(ready + 8 * (st & 3), 1), ok_w = try_wait(ready + 32 + 8 * (st & 3), 1);
          if (!ok_h || !ok_w) __trap();
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          if (elect_one()) {
            const uint32_t b0 = w_desc0 + st * (16384 >> 4);
#pragma unroll
            for (int m = 0; m < NM; ++m) mma_ts(tb, a_tmem + (st & 7) * 32 + (m & 3) * 8, mk_desc(b0 + 2 * (m & 1) + (m >> 1) * (8192 >> 4)), idesc, m ? 1u : acc);
            commit(bars + 8 * st);
          }
          __syncwarp();
          acc = 1;
        }
      }
    } else {
      for (int i = 0; i < iters; ++i) {
        if (PAT >= 2) {
          const bool ok_h = try_wait(ready + 8 * (i & 7), 1), ok_w = try_wait(bars + 8 * (12 + stage), 1);
          if (!ok_h || !ok_w) __trap();
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        const uint32_t b0 = w_desc0 + stage * (16384 >> 4);
        if (elect_one()) {
#pragma unroll
          for (int m = 0; m < NM; ++m) mma_ts(tb, a_tmem + (i & 7) * 32 + (m & 3) * 8, mk_desc(b0 + 2 * (m & 1) + (m >> 1) * (8192 >> 4)), idesc, m ? 1u : acc);
          if (PAT >= 1) commit(bars + 8 * stage);
        }
        __syncwarp();
        acc = 1;
        if (++stage == kStages) stage = 0;
      }
    }
    if (elect_one()) commit(bars + 8 * 20);
    __syncwarp();
    long long t1 = clock64();
    while (!try_wait(bars + 8 * 20, 0)) {
      if (clock64() - t1 > 2000000000LL) break;
    }
    long long t2 = clock64();
    if (lane == 0) { cyc[blockIdx.x * 2] = t2 - t0; cyc[blockIdx.x * 2 + 1] = t1 - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512));
  }
}


// ---- the same loop FED BY A WEIGHT STREAM: producer thread(s) bulk-copy [N x 32] bf16 chunks (N * 64 bytes) from an
// L2-resident 1.2 MB buffer into a ring of shared-memory stages, the MMA warp waits for each stage, issues 2 MMAs on it and
// frees it with tcgen05.commit - the MLP kernel's main loop without any epilogue.  Is an iteration still 2 x N / 2 cycles?
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  long long t0 = clock64();
  while (!try_wait(bar, parity)) { if (clock64() - t0 > 2000000000LL) __trap(); }
}
template <int NPROD, int CPS>
__global__ void __launch_bounds__(192, 1) ks(int N, int stages, int iters, const uint8_t* __restrict__ src, int src_chunks, long long* cyc) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tslot;
  __shared__ __align__(8) uint64_t bar[40];
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  const uint32_t bars = smem_u32(&bar[0]);          // full[s] = bars + 8 s, empty[s] = bars + 8 (16 + s)
  if (threadIdx.x == 0) {
    for (int i = 0; i < 40; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bars + 8 * i));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tslot;
  const uint32_t bytes = (uint32_t)N * 64u * CPS;   // one stage = CPS chunks of [N x 32] bf16, ONE bulk copy, one barrier pair
  const uint32_t sbase = smem_u32(smem);
  if (warp >= 2 && warp < 2 + NPROD && lane == 0) {
    const int pidx = warp - 2;
    int stage = 0; uint32_t phase = 0;
    for (int i = 0; i < iters; ++i) {
      if (i % NPROD == pidx) {
        mbar_wait(bars + 8 * (16 + stage), phase ^ 1);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bars + 8 * stage), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(sbase + stage * bytes), "l"(src + (size_t)((i + blockIdx.x * 5) % src_chunks) * bytes), "r"(bytes), "r"(bars + 8 * stage) : "memory");
      }
      if (++stage == stages) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t w_desc0 = desc_lo(sbase);
    const uint32_t a_tmem = tb + 256;
    long long t0 = clock64();
    int stage = 0; uint32_t phase = 0, acc = 0;
    for (int i = 0; i < iters; ++i) {
      mbar_wait(bars + 8 * stage, phase);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t b0 = w_desc0 + stage * (bytes >> 4);
      if (elect_one()) {
#pragma unroll
        for (int cc = 0; cc < CPS; ++cc) {
          mma_ts(tb, a_tmem + ((i * CPS + cc) & 7) * 32, mk_desc(b0 + cc * ((uint32_t)N * 64u >> 4)), idesc, cc ? 1u : acc);
          mma_ts(tb, a_tmem + ((i * CPS + cc) & 7) * 32 + 8, mk_desc(b0 + cc * ((uint32_t)N * 64u >> 4) + 2), idesc, 1u);
        }
        commit(bars + 8 * (16 + stage));
      }
      __syncwarp();
      acc = 1;
      if (++stage == stages) { stage = 0; phase ^= 1; }
    }
    if (elect_one()) commit(bars + 8 * 39);
    __syncwarp();
    mbar_wait(bars + 8 * 39, 0);
    long long t2 = clock64();
    if (lane == 0) { cyc[blockIdx.x * 2] = t2 - t0; cyc[blockIdx.x * 2 + 1] = t2 - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512));
  }
}

template <int NPROD, int CPS>
static void run_stream(int N, int stages, int grid, const uint8_t* src) {
  const int iters = 4096 / CPS;
  long long* d;
  cudaMalloc(&d, grid * 2 * sizeof(long long));
  cudaMemset(d, 0, grid * 2 * sizeof(long long));
  const int smem = 96 * 1024 + 1024;
  cudaFuncSetAttribute(ks<NPROD, CPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int rep = 0; rep < 2; ++rep) {
    ks<NPROD, CPS><<<grid, 192, smem>>>(N, stages, iters, src, (1184 * 1024) / (N * 64 * CPS), d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); exit(1); }
  }
  long long h[2 * 160];
  cudaMemcpy(h, d, grid * 2 * sizeof(long long), cudaMemcpyDeviceToHost);
  double tot = 0; int cnt = 0;
  for (int i = 0; i < grid; ++i) if (h[2 * i] > 0) { tot += h[2 * i]; ++cnt; }
  printf("streamed: N=%3d, %d chunk(s) = %2d KB per stage, %2d stages, %d producer thread(s): %.1f cycles/stage (tensor floor %d) = %.1f B/cycle/SM of weights\n", N,
         CPS, N * 64 * CPS / 1024, stages, NPROD, tot / cnt / iters, N * CPS, (double)N * 64 * CPS / (tot / cnt / iters));
  cudaFree(d);
}

// ---- TMEM port contention: the bare MMA loop (TS form, 2 MMAs + commit per iteration, operands resident) while `nld` other
// warps read accumulator-sized slices with tcgen05.ld.32x32b.x32 as fast as they can (the epilogue's access pattern), each
// optionally writing 16 columns back with tcgen05.st (the packed activations).
__global__ void __launch_bounds__(416, 1) kt(int N, int iters, int nld, int with_st, int pace, long long* cyc, unsigned long long* nloads) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tslot;
  __shared__ __align__(8) uint64_t bar[32];
  __shared__ volatile int done;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    done = 0;
    for (int i = 0; i < 32; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  for (int i = threadIdx.x; i < 120 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + i;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tslot;
  if (warp == 1) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t w_desc0 = desc_lo(smem_u32(smem));
    const uint32_t bars = smem_u32(&bar[0]);
    const uint32_t a_tmem = tb + 256;
    long long t0 = clock64();
    int stage = 0; uint32_t acc = 0;
    for (int i = 0; i < iters; ++i) {
      const uint32_t b0 = w_desc0 + stage * (16384 >> 4);
      if (elect_one()) {
        mma_ts(tb, a_tmem + (i & 7) * 32, mk_desc(b0), idesc, acc);
        mma_ts(tb, a_tmem + (i & 7) * 32 + 8, mk_desc(b0 + 2), idesc, 1u);
        commit(bars + 8 * stage);
      }
      __syncwarp();
      acc = 1;
      if (++stage == kStages) stage = 0;
    }
    if (elect_one()) commit(bars + 8 * 20);
    __syncwarp();
    while (!try_wait(bars + 8 * 20, 0)) { }
    long long t2 = clock64();
    done = 1;
    if (lane == 0) { cyc[blockIdx.x * 2] = t2 - t0; cyc[blockIdx.x * 2 + 1] = t2 - t0; }
  } else if (warp >= 4 && warp < 4 + nld) {
    const uint32_t t_lane = tb + ((uint32_t)((warp & 3) * 32) << 16) + 256;     // read the OTHER buffer (no data hazard with D)
    unsigned long long n = 0; uint32_t x = 0;
    while (!done) {
      uint32_t v[32];
      const uint32_t ta = t_lane + ((uint32_t)n & 7) * 32;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
            "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
            "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
            "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
          : "r"(ta) : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int i = 0; i < 32; ++i) x ^= v[i];
      if (with_st) {
        asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                     ::"r"(tb + ((uint32_t)((warp & 3) * 32) << 16) + 480), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
                       "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      }
      ++n;
      if (pace) { const long long t = clock64(); while (clock64() - t < pace) { } }
    }
    if (lane == 0) { atomicAdd(nloads, n); if (x == 0x12345678u) cyc[0] = 0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512));
  }
}
static void run_tmem(int N, int nld, int with_st, int pace, int grid) {
  const int iters = 3072;
  long long* d; unsigned long long* nl;
  cudaMalloc(&d, grid * 2 * sizeof(long long)); cudaMemset(d, 0, grid * 2 * sizeof(long long));
  cudaMalloc(&nl, 8); cudaMemset(nl, 0, 8);
  const int smem = 120 * 1024;
  cudaFuncSetAttribute(kt, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  kt<<<grid, 416, smem>>>(N, iters, nld, with_st, pace, d, nl);
  cudaMemset(nl, 0, 8);
  kt<<<grid, 416, smem>>>(N, iters, nld, with_st, pace, d, nl);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); exit(1); }
  long long h[2 * 160]; unsigned long long hn = 0;
  cudaMemcpy(h, d, grid * 2 * sizeof(long long), cudaMemcpyDeviceToHost);
  cudaMemcpy(&hn, nl, 8, cudaMemcpyDeviceToHost);
  double tot = 0; int cnt = 0;
  for (int i = 0; i < grid; ++i) if (h[2 * i] > 0) { tot += h[2 * i]; ++cnt; }
  const double cyc_per_cta = tot / cnt;
  printf("TMEM contention: N=%3d MMAs + %d warps of tcgen05.ld.x32%s (pace %4d): %.1f cycles/iteration (floor %d); loads: %.1f B/cycle/SM\n", N, nld,
         with_st ? " + st.x16" : "", pace, cyc_per_cta / iters, N, (double)hn / cnt * 4096.0 / cyc_per_cta);
  cudaFree(d); cudaFree(nl);
}

template <int PAT, int NM>
static void run(int N, int grid) {
  const int iters = 1536;   // multiple of kStages
  long long* d;
  cudaMalloc(&d, grid * 2 * sizeof(long long));
  cudaMemset(d, 0, grid * 2 * sizeof(long long));
  const int smem = 120 * 1024;
  cudaFuncSetAttribute(k<PAT, NM>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int rep = 0; rep < 2; ++rep) {
    k<PAT, NM><<<grid, 128, smem>>>(N, iters, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); exit(1); }
  }
  long long h[2 * 160];
  cudaMemcpy(h, d, grid * 2 * sizeof(long long), cudaMemcpyDeviceToHost);
  double tot = 0, iss = 0; int cnt = 0;
  for (int i = 0; i < grid; ++i) if (h[2 * i] > 0) { tot += h[2 * i]; iss += h[2 * i + 1]; ++cnt; }
  printf("pattern %d, %d MMAs/iteration, N=%3d: %.1f cycles/iteration (issue loop alone %.1f); tensor floor %d\n", PAT, NM, N,
         tot / cnt / iters, iss / cnt / iters, NM * N / 2);
  cudaFree(d);
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  uint8_t* src; cudaMalloc(&src, 1184 * 1024); cudaMemset(src, 0x3c, 1184 * 1024);
  run_stream<1, 1>(256, 6, sms, src); run_stream<2, 1>(256, 6, sms, src);
  run_stream<1, 2>(256, 3, sms, src); run_stream<2, 2>(256, 3, sms, src);
  run_stream<1, 2>(160, 3, sms, src); run_stream<1, 2>(128, 6, sms, src); run_stream<1, 4>(128, 3, sms, src); run_stream<2, 4>(128, 3, sms, src);
  run_stream<1, 3>(256, 2, sms, src);
  for (int N : {256, 128}) {
    run_tmem(N, 0, 0, 0, sms); run_tmem(N, 4, 0, 0, sms); run_tmem(N, 8, 0, 0, sms); run_tmem(N, 8, 1, 0, sms);
    run_tmem(N, 8, 1, 200, sms); run_tmem(N, 8, 1, 400, sms); run_tmem(N, 8, 1, 800, sms);
  }
  if (getenv("MMA_ISSUE_STREAM_ONLY")) return 0;
  for (int N : {256, 160, 128, 64, 32}) {
    run<0, 2>(N, sms); run<1, 2>(N, sms); run<2, 2>(N, sms); run<3, 2>(N, sms);
    run<0, 4>(N, sms); run<1, 4>(N, sms); run<2, 4>(N, sms); run<3, 4>(N, sms);
  }
  return 0;
}
