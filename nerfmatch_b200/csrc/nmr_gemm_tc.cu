// fp32-in / fp32-out GEMM on the 5th-gen tensor cores: C = epi(A1 * W1^T + A2 * W2^T), the same contract as the SIMT GEMM
// (nmr_mlp_simt.cu, GemmParams) at ~10x its rate, for the two places that need fp32-grade results but not the SIMT
// kernel's exact summation order:
//   * the dgrad chain (and the activation re-forward) of the iNeRF refinement's backward pass (nmr_stage_mlp_backward;
//     nerfmatch_evaluator.py:288-500 differentiates nerf_fine w.r.t. its inputs),
//   * the similarity matrix of the NeRFMatch-Mini coarse matcher (nerfmatch_coarse_trainer.py:223-234: einsum over 256-d
//     features of ~6 K x 6 K pairs).
// tcgen05 has no fp32 kind, so every fp32 operand is split on the fly into bf16 hi + bf16 lo (x = hi + lo + O(2^-17 |x|)) and
// each K = 16 step issues three kind::f16 MMAs into one fp32 TMEM accumulator:  hi*hi + hi*lo + lo*hi  (the dropped lo*lo term
// is O(2^-16) relative, the same size as the split residual) - "3 x bf16": ~2^-16 relative error per product instead of
// bf16's 2^-8, i.e. the accuracy class of TF32x3 / fp32 for well-conditioned sums.
//
// One CTA computes a 128 x 256 tile: all 256 threads stream fp32 K-chunks (32 wide) of A and W from global memory, split
// them and store the four bf16 blocks in the UMMA K-major SWIZZLE_64B layout (double buffered, the next chunk prefetched; weight
// panels that were pre-split on the host arrive by one 32 KB bulk copy per chunk instead); one thread issues the MMAs
// (tcgen05.commit frees a stage when its MMAs have retired); warps 0-3 then read the accumulator (tcgen05.ld) and apply
// the epilogue (bias, rank-1 term, added matrix, activation, ReLU mask).  Two CTAs fit an SM (96 KB + 256 TMEM columns
// each), so one CTA's loads overlap the other's MMAs.
#include <cstring>

#include "nmr_internal.cuh"

namespace nmr {
namespace {

constexpr int BM = 128, BN = 256, BK = 32;
constexpr int kThreadsG = 256;
constexpr int kABlk = BM * BK * 2;          // 8 KB: [128 x 32] bf16
constexpr int kBBlk = BN * BK * 2;          // 16 KB: [256 x 32] bf16
constexpr int kStage = 2 * kABlk + 2 * kBBlk;   // hi + lo of both operands: 48 KB
constexpr int kStagesG = 2;
constexpr int kSmemG = kStagesG * kStage + 1024 /* alignment slack */ + 64;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// bounded: a protocol bug must not hang the box
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 2000000000LL) __trap();
  }
}
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr) { return ((saddr >> 4) & 0x3FFFu) | (1u << 16); }
__device__ __forceinline__ uint64_t mk_desc(uint32_t lo) {
  constexpr uint32_t hi = (512u >> 4) | (1u << 14) | (4u << 29);   // SBO = 512 B | version 1 | SWIZZLE_64B
  uint64_t d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
  return d;
}
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t sw64_off(int r, int c) {   // 16-byte chunk c of row r in a [rows x 32 bf16] 64B-swizzled block
  return (uint32_t)(((r >> 3) << 9) + ((r & 7) << 6) + ((c ^ ((r & 7) >> 1)) << 4));
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
__device__ __forceinline__ float bf16_round(float x) { return __uint_as_float(pack_bf16(x, 0.f) << 16); }

// 8 consecutive fp32 -> 8 bf16 hi and 8 bf16 lo (x - hi, exact in fp32, rounded to bf16), one 16-byte store each
__device__ __forceinline__ void split_store(const float4& u, const float4& v, uint32_t hi_addr, uint32_t lo_addr) {
  const float x[8] = {u.x, u.y, u.z, u.w, v.x, v.y, v.z, v.w};
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float h0 = bf16_round(x[2 * i]), h1 = bf16_round(x[2 * i + 1]);
    h[i] = pack_bf16(h0, h1);
    l[i] = pack_bf16(x[2 * i] - h0, x[2 * i + 1] - h1);
  }
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(hi_addr), "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]) : "memory");
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(lo_addr), "r"(l[0]), "r"(l[1]), "r"(l[2]), "r"(l[3]) : "memory");
}

__device__ __forceinline__ float act_of(float v, int act) {
  if (act == 1) return fmaxf(v, 0.0f);
  if (act == 2) return 1.0f / (1.0f + expf(-v));
  return v;
}

__global__ void __launch_bounds__(kThreadsG, 2) gemm3_tc_kernel(const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;                 // the swizzled UMMA tiles need a 1024-byte aligned window
  uint8_t* gbase = smem_raw + (base - raw);
  const uint32_t bars = base + kStagesG * kStage;               // empty[0], empty[1], accfull, tmem slot, bfull[0], bfull[1]
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gbase + kStagesG * kStage + 32);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t row0 = (int64_t)blockIdx.x * BM;
  const int col0 = blockIdx.y * BN;

  if (tid == 0) {
    mbar_init(bars + 0, 1); mbar_init(bars + 8, 1); mbar_init(bars + 16, 1); mbar_init(bars + 40, 1); mbar_init(bars + 48, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(bars + 32), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;

  const int c1 = (p.K1 + BK - 1) / BK, c2 = p.A2 ? (p.K2 + BK - 1) / BK : 0;
  const int nchunks = c1 + c2;
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

  // global -> registers for chunk t: issued one iteration AHEAD (right after chunk t-1's MMAs are queued), so the load latency
  // overlaps the tensor core's work on the previous chunk and the wait for the stage
  float4 ua[2], va[2], ub[4], vb[4];
  auto load_chunk = [&](int t) {
    const float* A; const float* W; int lda, ldw, k0, K;
    if (t < c1) { A = p.A1; lda = p.lda1; W = p.W1; ldw = p.K1; k0 = t * BK; K = p.K1; }
    else { A = p.A2; lda = p.lda2; W = p.W2; ldw = p.K2; k0 = (t - c1) * BK; K = p.K2; }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int idx = i * kThreadsG + tid, r = idx >> 2, c = idx & 3;
      const int64_t gr = row0 + r;
      const bool ok = gr < p.M && k0 + 8 * c < K;
      const float4* src = reinterpret_cast<const float4*>(A + gr * lda + k0 + 8 * c);
      ua[i] = ok ? __ldg(src) : make_float4(0.f, 0.f, 0.f, 0.f);
      va[i] = ok ? __ldg(src + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if ((t < c1 ? p.W1p : p.W2p) != nullptr) return;               // pre-split panel: fetched by a bulk copy (below)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = i * kThreadsG + tid, r = idx >> 2, c = idx & 3;
      const int gn = col0 + r;
      const bool ok = gn < p.N && k0 + 8 * c < K;
      const float4* src = reinterpret_cast<const float4*>(W + (int64_t)gn * ldw + k0 + 8 * c);
      ub[i] = ok ? __ldg(src) : make_float4(0.f, 0.f, 0.f, 0.f);
      vb[i] = ok ? __ldg(src + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  // pre-split W chunk t -> both B blocks of its stage with one 32 KB bulk copy (thread 0; completes on bfull[stage])
  auto fetch_packed = [&](int t) {
    const uint8_t* Wp = t < c1 ? p.W1p : p.W2p;
    if (Wp == nullptr) return;
    const int st = t & 1, kc = t < c1 ? t : t - c1;
    const uint32_t bar = bars + 40 + 8 * st;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(2 * kBBlk) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(base + st * kStage + 2 * kABlk), "l"(Wp + (size_t)kc * (2 * kBBlk)), "r"(2 * kBBlk), "r"(bar) : "memory");
  };
  uint32_t bfull_phase = 0;                                          // bit st: parity of the next completion of bfull[st] (thread 0)
  if (tid == 0) fetch_packed(0);
  load_chunk(0);
  for (int t = 0; t < nchunks; ++t) {
    const int st = t & 1;
    // ---- the MMAs that read this stage two chunks ago have retired ----
    if (t >= kStagesG) mbar_wait(bars + 8 * st, (uint32_t)((t / kStagesG) - 1) & 1u);
    const uint32_t sA_hi = base + st * kStage, sA_lo = sA_hi + kABlk, sB_hi = sA_lo + kABlk, sB_lo = sB_hi + kBBlk;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int idx = i * kThreadsG + tid, r = idx >> 2, c = idx & 3;
      split_store(ua[i], va[i], sA_hi + sw64_off(r, c), sA_lo + sw64_off(r, c));
    }
    const bool packed = (t < c1 ? p.W1p : p.W2p) != nullptr;
    if (!packed) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int idx = i * kThreadsG + tid, r = idx >> 2, c = idx & 3;
        split_store(ub[i], vb[i], sB_hi + sw64_off(r, c), sB_lo + sw64_off(r, c));
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy stores -> visible to the tensor core's async proxy
    __syncthreads();
    if (tid == 0) {
      if (packed) { mbar_wait(bars + 40 + 8 * st, (bfull_phase >> st) & 1u); bfull_phase ^= 1u << st; }
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t ah = desc_lo(sA_hi), al = desc_lo(sA_lo), bh = desc_lo(sB_hi), bl = desc_lo(sB_lo);
#pragma unroll
      for (int k = 0; k < 2; ++k) {                                    // two K = 16 steps per chunk: +32 bytes = +2 descriptor units
        umma(tmem, mk_desc(ah + 2 * k), mk_desc(bh + 2 * k), idesc, (t > 0 || k > 0) ? 1u : 0u);
        umma(tmem, mk_desc(ah + 2 * k), mk_desc(bl + 2 * k), idesc, 1u);
        umma(tmem, mk_desc(al + 2 * k), mk_desc(bh + 2 * k), idesc, 1u);
      }
      umma_commit(bars + 8 * st);                                      // stage free once these MMAs retire
      if (t == nchunks - 1) umma_commit(bars + 16);                    // accumulator complete
      if (t + 1 < nchunks && ((t + 1 < c1 ? p.W1p : p.W2p) != nullptr)) {
        // the next chunk's stage was last read by the MMAs of chunk t - 1: once they have retired its B blocks may be refilled
        if (t + 1 >= kStagesG) mbar_wait(bars + 8 * ((t + 1) & 1), (uint32_t)(((t + 1) / kStagesG) - 1) & 1u);
        fetch_packed(t + 1);
      }
    }
    if (t + 1 < nchunks) load_chunk(t + 1);
  }

  // ---- epilogue: warps 0..3 own TMEM lane quarters 0..3.  tcgen05.ld hands every thread ONE ROW (32 columns of it per
  // chunk); touching global memory in that shape would cost 32 wavefronts per access (rows are >= 1 KB apart), so each warp
  // transposes the chunk through shared memory (the operand stages are free by now; 33-float rows: conflict-free both ways)
  // and then works on four rows x 128 contiguous bytes per instruction: mask / add loads and the stores are fully coalesced.
  if (warp < 4) {
    mbar_wait(bars + 16, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    float* tile = reinterpret_cast<float*>(gbase) + warp * (32 * 33);
    const int64_t rbase = row0 + warp * 32;
    const int rsub = lane >> 3, c4 = (lane & 7) * 4;                  // this lane's row within a group of four, and its 4 columns
    const bool vec_ok = (p.ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0);
    const bool vec_in = (!p.mask || ((p.ldm % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.mask) & 15) == 0))) &&
                        (!p.add || ((p.ldadd % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.add) & 15) == 0)));
#pragma unroll 1
    for (int j = 0; j < BN / 32; ++j) {
      const int c0 = col0 + j * 32;
      if (c0 >= p.N) break;                                            // warp-uniform
      const int cc0 = c0 + c4;
      const bool full = cc0 + 4 <= p.N;
      // all global reads of this column chunk are issued FIRST (16 independent 16-byte loads per thread): their latency
      // overlaps the TMEM load and the transposition instead of being paid once per row group
      float4 mk[8], ad[8];
      float r1[8];
      uint32_t mb[8];                                                  // bit-mask form: this lane's 4 bits of the row's word
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int64_t r = rbase + 4 * i + rsub;
        mk[i] = make_float4(1.f, 1.f, 1.f, 1.f); ad[i] = make_float4(0.f, 0.f, 0.f, 0.f); r1[i] = 0.f; mb[i] = 0xfu;
        if (r < p.M) {
          if (p.r1a) r1[i] = __ldg(p.r1a + r * p.ldr1);
          if (p.mask_bits) mb[i] = (__ldg(p.mask_bits + mask_word_index(p.mask_row0 + r, p.mask_step, c0 >> 5)) >> c4) & 0xfu;
          if (full && vec_in) {
            if (p.mask) mk[i] = __ldg(reinterpret_cast<const float4*>(p.mask + r * p.ldm + cc0));
            if (p.add) ad[i] = __ldg(reinterpret_cast<const float4*>(p.add + r * p.ldadd + cc0));
          } else {
            float m4[4] = {1.f, 1.f, 1.f, 1.f}, a4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              if (cc0 + e < p.N) {
                if (p.mask) m4[e] = p.mask[r * p.ldm + cc0 + e];
                if (p.add) a4[e] = p.add[r * p.ldadd + cc0 + e];
              }
            }
            mk[i] = make_float4(m4[0], m4[1], m4[2], m4[3]); ad[i] = make_float4(a4[0], a4[1], a4[2], a4[3]);
          }
        }
      }
      float bias4[4] = {0.f, 0.f, 0.f, 0.f}, r1b4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (cc0 + e < p.N) {
          if (p.bias) bias4[e] = p.bias[cc0 + e];
          if (p.r1a) r1b4[e] = p.r1b[cc0 + e];
        }
      }
      uint32_t v[32];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
            "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
            "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
            "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
          : "r"(tmem + ((uint32_t)(warp * 32) << 16) + j * 32) : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      __syncwarp();                                                    // the previous chunk's reads of the tile are done
#pragma unroll
      for (int c = 0; c < 32; ++c) tile[lane * 33 + c] = __uint_as_float(v[c]);
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int rr = 4 * i + rsub;
        const int64_t r = rbase + rr;
        if (r >= p.M) continue;
        const float m4[4] = {mk[i].x, mk[i].y, mk[i].z, mk[i].w}, a4[4] = {ad[i].x, ad[i].y, ad[i].z, ad[i].w};
        float x[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float y = tile[rr * 33 + c4 + e] + bias4[e];
          y = fmaf(r1[i], r1b4[e], y);
          y += a4[e];
          y = act_of(y, p.act);
          if (!(m4[e] > 0.f) || !((mb[i] >> e) & 1u)) y = 0.f;
          x[e] = y;
        }
        float* out = p.C + r * p.ldc + cc0;
        if (vec_ok && full) {
          *reinterpret_cast<float4*>(out) = make_float4(x[0], x[1], x[2], x[3]);
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (cc0 + e < p.N) out[e] = x[e];
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
  }
}

}  // namespace

static uint16_t f2bf_host(float f) {
  uint32_t u; memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
size_t packed_panel_bytes(int K) { return (size_t)((K + BK - 1) / BK) * 2 * kBBlk; }
void pack_split_panel(const float* W, int N, int K, uint8_t* out) {
  const int nch = (K + BK - 1) / BK;
  memset(out, 0, packed_panel_bytes(K));
  for (int ch = 0; ch < nch; ++ch)
    for (int n = 0; n < N && n < BN; ++n)
      for (int k = 0; k < BK; ++k) {
        const int kk = ch * BK + k;
        if (kk >= K) continue;
        const float x = W[(size_t)n * K + kk];
        const uint16_t hi = f2bf_host(x);
        uint32_t hu = (uint32_t)hi << 16; float hf; memcpy(&hf, &hu, 4);
        const uint16_t lo = f2bf_host(x - hf);
        const size_t off = (size_t)(n >> 3) * 512 + (n & 7) * 64 + (((k >> 3) ^ ((n & 7) >> 1)) << 4) + (k & 7) * 2;
        memcpy(out + (size_t)ch * 2 * kBBlk + off, &hi, 2);
        memcpy(out + (size_t)ch * 2 * kBBlk + kBBlk + off, &lo, 2);
      }
}

// Requirements beyond launch_sgemm's: rows of A / W 16-byte aligned (lda, K multiples of 4 floats), K multiples of 8.
bool gemm_tc_supported(const GemmParams& p) {
  auto ok = [](const float* a, int ld, int K) { return a && (ld % 4 == 0) && (K % 8 == 0) && ((reinterpret_cast<uintptr_t>(a) & 15) == 0); };
  if (p.N <= 4 && !p.A2) return false;                                // thin heads stay on the warp-per-row kernel
  if (!ok(p.A1, p.lda1, p.K1) || !ok(p.W1, p.K1, p.K1)) return false;
  if (p.A2 && (!ok(p.A2, p.lda2, p.K2) || !ok(p.W2, p.K2, p.K2))) return false;
  return true;
}

int launch_gemm_tc(const GemmParams& p, cudaStream_t s) {
  if (p.M == 0) return 0;
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    if (cudaFuncSetAttribute(gemm3_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemG) != cudaSuccess) return -1;
    attr_set[dev] = true;
  }
  dim3 grid((unsigned)((p.M + BM - 1) / BM), (unsigned)((p.N + BN - 1) / BN));
  gemm3_tc_kernel<<<grid, kThreadsG, kSmemG, s>>>(p);
  return 0;
}

}  // namespace nmr
