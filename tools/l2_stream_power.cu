// Micro-benchmark: what does the MLP kernel's L2 -> SM weight stream cost in POWER?
// Every SM streams the same 1.2 MB buffer (L2 resident) into a 6-stage shared-memory ring with cp.async.bulk, exactly like the
// MLP kernel's weight producer, at a chosen rate (bytes per cycle per SM; the MLP kernel moves ~50 B/cycle/SM = 12.4 TB/s), for
// a few seconds, while `nvidia-smi --query-gpu=power.draw,clocks.sm -lms 200` runs beside it.  rate 0 = the same CTAs spinning
// on the clock without any copy (baseline), rate < 0 = unthrottled.  With -mc 1 CTA pairs (cluster of 2) fetch half of every
// chunk each and multicast it to both CTAs' rings (cp.async.bulk ... .multicast::cluster): the same bytes land in shared memory
// with half the L2 -> SM traffic.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2_stream_power l2_stream_power.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  long long t0 = clock64();
  while (!mbar_try(bar, parity)) if (clock64() - t0 > 4000000000LL) __trap();
}
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }

constexpr int kChunk = 16384, kStages = 6;

template <int MC>
__global__ void __launch_bounds__(64, 1) stream_kernel(const uint8_t* __restrict__ src, int total_bytes, long long total_chunks, float bytes_per_cycle,
                                                      long long spin_cycles, unsigned long long* moved) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t full[kStages];
  const uint32_t sb = (smem_u32(smem) + 1023u) & ~1023u;
  const uint32_t rank = MC ? cluster_rank() : 0u;
  if (threadIdx.x == 0) { for (int i = 0; i < kStages; ++i) mbar_init(smem_u32(&full[i]), 1); asm volatile("fence.mbarrier_init.release.cluster;"); }
  __syncthreads();
  if (MC) cluster_sync_all();
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    if (bytes_per_cycle == 0.f) {
      while (clock64() - t0 < spin_cycles) {}
    } else {
      const int nchunks = total_bytes / kChunk;
      long long issued = 0, done = 0;
      uint32_t phase_bits = 0;
      while (done < total_chunks) {
        while (issued < total_chunks && issued - done < kStages) {
          if (bytes_per_cycle > 0.f) {                      // pace: chunk i may start at t0 + i * chunk / rate
            const long long due = t0 + (long long)((double)issued * kChunk / bytes_per_cycle);
            while (clock64() < due) {}
          }
          const int st = (int)(issued % kStages);
          const uint32_t bar = smem_u32(&full[st]);
          const uint8_t* g = src + (size_t)(issued % nchunks) * kChunk;
          mbar_expect(bar, kChunk);
          if (MC) {       // this CTA fetches its half and multicasts it to both CTAs of the pair; the peer does the other half
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                         ::"r"(sb + st * kChunk + rank * (kChunk / 2)), "l"(g + rank * (kChunk / 2)), "r"(kChunk / 2), "r"(bar), "h"((uint16_t)3) : "memory");
          } else {
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(sb + st * kChunk), "l"(g), "r"(kChunk), "r"(bar) : "memory");
          }
          ++issued;
        }
        const int st = (int)(done % kStages);
        mbar_wait(smem_u32(&full[st]), (phase_bits >> st) & 1);
        phase_bits ^= 1u << st;
        ++done;
        // (pair mode: both CTAs run the same schedule; a stage is re-armed only after this CTA saw it complete, and the peer cannot
        //  be more than kStages ahead because its own waits need our halves)
      }
    }
    if (blockIdx.x == 0) *moved = (unsigned long long)(clock64() - t0);
  }
  __syncthreads();
  if (MC) cluster_sync_all();
}

int main(int argc, char** argv) {
  float rate = 50.f; float seconds = 4.f; int mc = 0;
  for (int i = 1; i + 1 < argc; i += 2) {
    if (!strcmp(argv[i], "-rate")) rate = atof(argv[i + 1]);
    if (!strcmp(argv[i], "-sec")) seconds = atof(argv[i + 1]);
    if (!strcmp(argv[i], "-mc")) mc = atoi(argv[i + 1]);
  }
  cudaSetDevice(0);
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount & ~1;
  const int total = 1184 * 1024;
  uint8_t* src; cudaMalloc(&src, total); cudaMemset(src, 1, total);
  unsigned long long* moved; cudaMalloc(&moved, 8);
  const int smem = kStages * kChunk + 2048;
  cudaFuncSetAttribute(stream_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(stream_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const double clk = 1.7e9;                                    // nominal pacing clock; the achieved rate is reported from events
  const double eff_rate = rate > 0 ? rate : 120.0;
  long long chunks = (long long)(seconds * clk * eff_rate / kChunk);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  if (mc) {
    cudaLaunchConfig_t cfg{}; cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.gridDim = dim3(sms); cfg.blockDim = dim3(64); cfg.dynamicSmemBytes = smem; cfg.attrs = at; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, stream_kernel<1>, (const uint8_t*)src, total, chunks, rate, (long long)(seconds * clk), moved);
  } else {
    stream_kernel<0><<<sms, 64, smem>>>(src, total, chunks, rate, (long long)(seconds * clk), moved);
  }
  cudaEventRecord(b);
  cudaError_t e = cudaDeviceSynchronize();
  float ms = 0; cudaEventElapsedTime(&ms, a, b);
  unsigned long long cyc = 0; cudaMemcpy(&cyc, moved, 8, cudaMemcpyDeviceToHost);
  const double bytes_into_smem = rate == 0 ? 0 : (double)chunks * kChunk * sms;
  printf("{\"mode\": \"%s\", \"multicast\": %d, \"status\": \"%s\", \"ms\": %.1f, \"smem_fill_TBps\": %.2f, \"l2_to_sm_TBps\": %.2f, \"sm_clock_ghz\": %.3f}\n",
         rate == 0 ? "spin" : (rate < 0 ? "unthrottled" : "paced"), mc, cudaGetErrorString(e), ms, bytes_into_smem / (ms * 1e-3) / 1e12,
         bytes_into_smem / (mc ? 2 : 1) / (ms * 1e-3) / 1e12, cyc / (ms * 1e-3) / 1e9);
  return 0;
}
