// Kernel (b): the NeRF MLP (models/nerf.py:94-144) as one persistent, warp-specialised sm_100a kernel.
//
//   * every Linear layer is a 128 x N x K tile GEMM on the 5th-gen tensor cores: tcgen05.mma
//     (kind::f16, bf16 operands, fp32 accumulate), accumulators in TMEM;
//   * positional encoding (kernel (a)'s encode half) is fused in: four encoder warps turn ray /
//     sample data into the bf16 A-operand tile of layer 0 / the skip layer directly in shared memory -
//     the encoded inputs never touch HBM;
//   * activations never leave the tensor core's own memory: the epilogue warps read the accumulator with
//     tcgen05.ld, apply ReLU, round to bf16 and write the packed result back INTO TMEM (tcgen05.st, over the
//     first half of the slice's own accumulator columns), where the next layer's MMAs take it as their A
//     operand (A-from-TMEM form of tcgen05.mma) - no shared-memory round trip, no swizzle, no proxy fence.
//     The skip connection is a second K segment into the same accumulator; feature_linear (no activation) is
//     folded into views_linears on the host (W_view_f * W_feat: one layer and 11 % of the weights disappear);
//     the sigma head (alpha_linear) rides in that GEMM as two extra output rows (bf16 hi / lo split of its
//     weights, fp32 accumulation); rgb (rgb_linear + sigmoid) is an fp32 dot product in the views epilogue;
//     the [dirs | appearance] part of views_linears is a per-ray fp32 bias computed once per ray; trunk biases
//     ride in the GEMM as well (ones columns of the encoded tile x bf16 hi / lo bias rows);
//   * the accumulator is double buffered in TMEM (2 x 256 columns, alternating by layer) and the epilogue
//     hands the next layer's operand over in 32-column chunks (one mbarrier per chunk), so the tensor core
//     starts layer l+1 on chunk 0 while the epilogue is still draining layer l.  Eight epilogue warps
//     (two per TMEM lane quarter, even / odd chunks) keep the hand-over faster than the MMA consumes it;
//   * the MMA-issuing warp runs its loop CONVERGED and one elected lane issues (elect.sync): descriptors then
//     live in uniform registers.  Inside an `if (lane == 0)` region every UTCHMMA / UTCBAR is wrapped in R2UR
//     moves and an ELECT ... BRA.U.ANY loop, ~110 instructions per K=32 chunk on one thread, which starved
//     the tensor core (36.6 K -> 28.0 K cycles per tile from this change alone; tools/mma_rate.cu);
//   * weights stream from L2 through a 3-stage ring of 32 KB shared-memory stages (two K=32 chunks per stage: a stage
//     round trip costs ~320 cycles whatever its size, more than one chunk of MMAs) filled by the bulk-copy
//     (TMA) engine (cp.async.bulk + mbarrier complete_tx); they are pre-swizzled on the host, so no tensor
//     map is needed;
//   * the tapped feature tile (stop_layer) is read back from TMEM after the hand-over (off the critical path) and
//     staged in shared memory, where it waits for the tile's densities (five layers later).  The encoder warps then
//     compute tile-local compositing weights (alpha, segmented exclusive product of 1 - alpha over the rows of the
//     same ray) and reduce the tile to ONE 256-vector per ray segment; the compositor only scales each segment by the
//     transmittance in front of it.  The per-sample features (512 B each, 100 GB per 480x640 view) never reach HBM.
//     Without that fusion (fp32 consumers of the stage API, feat_comb = max, < 43 samples per ray, NMR_TC_FUSE=0)
//     two otherwise idle warps stream the staged tile home with fully coalesced stores, in the tiled layout the
//     compositor reads;
//   * optional CTA-pair mode (NMR_TC_CG=2, tcgen05 cta_group::2): one leader thread issues M=256 MMAs for two
//     CTAs, each streaming only half of every weight chunk.  Correct and tested, but the cross-CTA hand-over
//     latency costs more cycles (26.7 K vs 24.1 K per tile) than the halved L2 traffic returns.
//
// Warp roles (512 threads): 0 weight producer, 1 MMA issuer, 2 TMEM allocator + tap drain (unfused), 3 tap drain (pair
// mode: weight-arrival relay of the peer CTA), 4-7 trunk epilogue (even chunks), 8-11 trunk epilogue (odd chunks) - both
// also run the views-layer epilogue -, 12-15 encoder (one tile ahead) + final rgb / sigma combine, raw store and the
// tile-local feature compositing.
#include <cstdlib>
#include <cstring>

#include "nmr_encode.cuh"
#include "nmr_tc.cuh"

namespace nmr {
namespace {

constexpr int kTileM = 128;
constexpr int kBlk = 8192;                 // one [128 rows x 32 bf16] block, 64B swizzle
constexpr int kEBytes = 3 * kBlk;          // 96-wide encoded-xyz tile
constexpr int kStageBytes = 16384;         // one [256 x 32] weight chunk
constexpr int kRingBytes = 6 * kStageBytes;
constexpr int kMaxStages = 12;             // CTA pair (cta_group::2): 12 stages of half chunks; single CTA: 6 full chunks
constexpr int kSteps = 9;                  // 8 trunk layers + (feature_linear folded into) views_linears.0
constexpr int kViewStep = 8;
constexpr int kThreads = 512;
constexpr int kTlSlots = 8 * kSteps * 4 + 64;   // clock64 trace: [it][step][4] of the MMA / epilogue warps + encoder-side stamps of iterations 2..5
constexpr int kFeatTileBytes = 8 * kBlk;   // tiled bf16 features (unfused tap): 64 KB per 128 points

constexpr int OFF_E = 0;
constexpr int OFF_W = OFF_E + 2 * kEBytes;
constexpr int OFF_TAP = OFF_W + kRingBytes;                   // feature-tap staging: 8 epilogue warps x 4 buffers x [32 rows x 64 B]
constexpr int kTapBytes = 2048;                               // (a whole tile: it waits there for the tile's densities)
constexpr int kTapBufs = 4;
constexpr int OFF_TAB = OFF_TAP + 8 * kTapBufs * kTapBytes;        // fp32 table: wrgb [3][128]
constexpr int kTabFloats = 3 * kHalf;
constexpr int OFF_PARTV = OFF_TAB + kTabFloats * 4;           // 2 tiles x 2 groups x float4 per row: (rgb partial sums, sigma) of the views epilogue
constexpr int kVbRays = 4;                                    // rays a 128-point tile can touch when S >= 43
constexpr int OFF_VBS = OFF_PARTV + 4 * kTileM * 16;          // 2 x [kVbRays][128] fp32: per-ray view bias of the tile (double buffered)
constexpr int OFF_WL = OFF_VBS + 2 * kVbRays * kHalf * 4;     // 2 x [128] fp32 local compositing weights + 2 x 4 x (value, flag) scan carries
constexpr int OFF_BAR = OFF_WL + 2 * kTileM * 4 + 64;
constexpr int kSeg = kVbRays;                                 // ray segments a 128-point tile can hold when S >= 43
constexpr int kSmemBytes = OFF_BAR + 1280;                     // dynamic smem is 1024-aligned (checked at kernel entry)
static_assert(kSmemBytes <= 232448, "shared memory budget");

// barrier slots (8 bytes each) inside the OFF_BAR area
enum { BAR_WFULL = 0, BAR_WEMPTY = BAR_WFULL + kMaxStages, BAR_PFULL = BAR_WEMPTY + kMaxStages, BAR_ACCFULL = BAR_PFULL + kMaxStages,
       BAR_HREADY = BAR_ACCFULL + 2, BAR_EFULL = BAR_HREADY + 8, BAR_EEMPTY = BAR_EFULL + 2, BAR_VEMPTY = BAR_EEMPTY + 2, BAR_VFULL = BAR_VEMPTY + 2,
       BAR_VPART = BAR_VFULL + 1, BAR_TAPFULL = BAR_VPART + 1, BAR_TAPFREE = BAR_TAPFULL + 8 * kTapBufs, BAR_TAPREAD = BAR_TAPFREE + 8 * kTapBufs, BAR_S0FULL = BAR_TAPREAD + 1, BAR_COUNT = BAR_S0FULL + 1 };
static_assert(BAR_COUNT * 8 + 8 <= 1280, "barrier area");

struct TcParams {
  const uint8_t* chunks;
  const float* wrgb;       // [3][128] rgb_linear weights (staged in shared memory: indexed constant-bank or L1-less global loads
                           //          cost the epilogue 2-3x, measured); trunk biases and the sigma head ride in the GEMMs
  float balpha; float brgb0, brgb1, brgb2;
  const float* vbias; int vb_div;          // view bias row = point / vb_div
  // input mode 0: rays + z (fused sampling-to-encoding); mode 1: padded fp32 rows X[P,kx]
  int mode; const float* rays; int ncols; const float* z; int S; int mip; int xyz; int fx; float var_scale;
  const float* X; int kx;
  int64_t P; int n_iters;
  int stop_step; int want_feat;
  float* raw; uint8_t* feat_tiles;
  float* partials;         // fused local compositing (mode 0, S >= 43): [tile][kSeg][256] fp32 instead of feat_tiles
  int debug_step; float* debug_out;
  uint32_t* masks;         // optional ReLU masks for the backward pass: [tile][9 steps][8 chunks][128 rows] words, bit c = column 32 j + c active
  int* error_flag;
  long long* timeline;   // optional clock64 trace of CTA 0 (NMR_TC_TIMELINE=1): [it][step][4]
};

__host__ __device__ constexpr int step_chunks_e(int s) { return (s == 0 || s == 5) ? 3 : 0; }
__host__ __device__ constexpr int step_chunks_h(int s) { return s == 0 ? 0 : 8; }
// trunk biases ride in the GEMM: columns 94 / 95 of the encoded tile are 1.0 and the matching weight rows hold the
// bf16 (lo, hi) split of the bias.  Layers fed by the encoded tile (0, 5) get it for free; the H-fed trunk layers
// append one K=16 MMA on a "bias chunk" (A = encoded columns 80..95, all other weight rows zero).
__host__ __device__ constexpr int step_bias_chunk(int s) { return (s >= 1 && s <= 7 && s != 5) ? 1 : 0; }
constexpr int kViewN = 160;   // 128 views outputs + sigma head (alpha_linear weights as bf16 hi/lo rows) + padding to a multiple of 32
__host__ __device__ constexpr int step_chunk_bytes(int s) { return s == kViewStep ? kViewN * 64 : 16384; }
// position of step s's first chunk in the tile's chunk sequence, chunks per tile, bytes of the ci-th chunk of a tile
__host__ __device__ constexpr int step_first_chunk(int s) {
  int n = 0;
  for (int i = 0; i < s; ++i) n += step_chunks_e(i) + step_chunks_h(i) + step_bias_chunk(i);
  return n;
}
constexpr int kTileChunks = step_first_chunk(kSteps);
static_assert(kTileChunks % 2 == 0 && step_first_chunk(kViewStep) % 2 == 0, "chunk pairs must not mix 16 KB and views chunks");
__host__ __device__ constexpr uint32_t tile_chunk_bytes(int ci) { return ci < step_first_chunk(kViewStep) ? 16384u : (uint32_t)(kViewN * 64); }
__host__ __device__ inline int64_t step_offset(int s) {   // byte offset of step s in the weight stream
  int64_t off = 0;
  for (int i = 0; i < s; ++i) off += (int64_t)(step_chunks_e(i) + step_chunks_h(i) + step_bias_chunk(i)) * step_chunk_bytes(i);
  return off;
}

// ---- PTX wrappers --------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must never hang the GPU box - after ~0.2 s record where and trap.
__device__ __noinline__ void mbar_wait_slow(uint32_t bar, uint32_t parity, int* error_flag, int code) {
  long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 400000000LL) {
      atomicExch(error_flag, code);
      __threadfence_system();
      __trap();
    }
  }
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int* error_flag, int code) {
  if (mbar_try_wait(bar, parity)) return;
  mbar_wait_slow(bar, parity, error_flag, code);
}
// ---- CTA-pair (cta_group::2) plumbing: the leader CTA (cluster rank 0) issues every MMA for both CTAs, so the
// barriers its MMA thread waits on collect arrivals from both CTAs (remote arrive through the cluster window).
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t mapa_rank0(uint32_t addr) {
  uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(r) : "r"(addr)); return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// (waits on barriers fed by both CTAs use the ordinary try_wait: a cluster-scope acquire on every probe costs several
// hundred cycles - measured 4x slower MMA issue - and the data the MMA reads was made visible by the arriving side's
// tcgen05.wait::st / fence.proxy.async before its arrive)
// one elected lane per warp signals the leader CTA's barrier (call after the warp's own fences, converged)
template <int CG>
__device__ __forceinline__ void warp_arrive_leader(uint32_t bar, uint32_t rank, int lane) {
  __syncwarp();
  if (lane == 0) {
    if (CG == 2 && rank != 0) mbar_arrive_cluster(mapa_rank0(bar)); else mbar_arrive(bar);
  }
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ bool elect_one() {      // true in exactly one lane of a converged warp
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// UMMA shared-memory descriptor, K-major, SWIZZLE_64B: rows of 64 B, 8-row groups 512 B apart (SBO), version 1.
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr) { return ((saddr >> 4) & 0x3FFFu) | (1u << 16); }
__device__ __forceinline__ uint64_t mk_desc(uint32_t lo) {
  constexpr uint32_t hi = (512u >> 4) | (1u << 14) | (4u << 29);   // SBO | version | SWIZZLE_64B  (bits 32..63)
  uint64_t d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
  return d;
}
// instruction descriptor: kind::f16, A = B = bf16 (format 1) or fp16 (format 0), K-major, D = fp32, M = 128.  Both formats run
// at the same tensor-core rate; fp16 carries 11 significand bits instead of 8 (8x smaller operand rounding, range +-65504).
template <int F16>
__device__ __forceinline__ uint32_t umma_idesc(int N, int M) {
  return (1u << 4) | (F16 ? 0u : ((1u << 7) | (1u << 10))) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// A operand from TMEM (the previous layer's activations, packed bf16 in place of its accumulator), B from shared memory
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// CTA-pair forms: one instruction drives both SMs' tensor cores (M = 256: 128 rows per CTA; each CTA supplies half of
// B's N rows from its own shared memory, at the same offsets); commit arrives on the same barrier in both CTAs.
__device__ __forceinline__ void umma2_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma2_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma2_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((uint16_t)3) : "memory");
}
template <int CG> __device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (CG == 2) umma2_bf16(d, a, b, idesc, acc); else umma_bf16(d, a, b, idesc, acc);
}
template <int CG> __device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (CG == 2) umma2_bf16_ts(d, a, b, idesc, acc); else umma_bf16_ts(d, a, b, idesc, acc);
}
template <int CG> __device__ __forceinline__ void mma_commit(uint32_t bar) {
  if (CG == 2) umma2_commit(bar); else umma_commit(bar);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* w) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]),
        "r"(w[8]), "r"(w[9]), "r"(w[10]), "r"(w[11]), "r"(w[12]), "r"(w[13]), "r"(w[14]), "r"(w[15]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// two fp32 -> packed 16-bit operands (a in the low half), round to nearest even; fp16 saturates at +-65504
template <int F16>
__device__ __forceinline__ uint32_t pack16(float a, float b) {
  uint32_t r;
  if (F16) asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  else asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
// low / high half of a packed pair of NON-NEGATIVE 16-bit values as fp32 bit patterns that a following multiply turns into
// the value: bf16 is the upper half of an fp32 (scale 1); an fp16 bit pattern shifted into the fp32 exponent / mantissa
// fields is the value times 2^-112, exactly, subnormals included (scale 2^112 - folded into the weight by the caller)
template <int F16> __device__ __forceinline__ float unpack_lo(uint32_t v) { return __uint_as_float(F16 ? ((v & 0xffffu) << 13) : (v << 16)); }
template <int F16> __device__ __forceinline__ float unpack_hi(uint32_t v) { return __uint_as_float(F16 ? ((v >> 16) << 13) : (v & 0xffff0000u)); }
constexpr float kF16Scale = 5.192296858534828e33f;   // 2^112
// byte offset of 16-byte chunk c (0..3) of row r inside a [128 x 32 bf16] 64B-swizzled block
__device__ __forceinline__ uint32_t sw64_off(int r, int c) {
  return (uint32_t)(((r >> 3) << 9) + ((r & 7) << 6) + ((c ^ ((r & 7) >> 1)) << 4));
}

// ---- fused row encoders (bf16 path) -----------------------------------------------------------
// The reference evaluates sin / cos of 2^k * x for k = 0..14 (embedding.py:35-46, 66-79).  2^k * x is exact
// in fp32, so a precise sincosf is taken at every 5th octave and the octaves in between follow from the
// double-angle identities (error growth 2^4 ulp << bf16 resolution); likewise exp(-0.5 var 4^k) is precise
// at every 4th octave and ^4 in between.  In mip mode the reference's "cos" is sin(fl(x + pi/2)); here it is
// cos(x) - they differ by at most half an ulp of the ARGUMENT (2^-11 at the top octave), below bf16
// resolution.  (The fp32 parity path, nmr_sampling.cu, evaluates the reference expressions literally.)
// One thread encodes one row: columns are produced in output order, 8 at a time -> one 16-byte chunk.
template <int F16>
struct RowWriter {
  uint32_t e_base; int r;
  __device__ __forceinline__ void store(int chunk, const float* f) const {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(e_base + (chunk >> 2) * kBlk + sw64_off(r, chunk & 3)),
                 "r"(pack16<F16>(f[0], f[1])), "r"(pack16<F16>(f[2], f[3])), "r"(pack16<F16>(f[4], f[5])), "r"(pack16<F16>(f[6], f[7])) : "memory");
  }
};

template <int F, class Writer>
__device__ __forceinline__ void encode_row_fourier(const float m[3], const Writer& w) {
  float f[8];
  float sn[3], cs[3];
#pragma unroll
  for (int col = 0; col < 96; ++col) {
    float v = 0.f;
    if (col < 3) {
      v = m[col];
    } else if (col < 3 + 6 * F) {
      const int jj = col - 3, k = jj / 6, rr = jj % 6;
      if (rr == 0) {
        if (k % 5 == 0) {
#pragma unroll
          for (int c = 0; c < 3; ++c) sincosf(m[c] * (float)(1 << k), &sn[c], &cs[c]);
        } else {
#pragma unroll
          for (int c = 0; c < 3; ++c) { float s2 = 2.f * sn[c] * cs[c]; cs[c] = fmaf(-2.f * sn[c], sn[c], 1.f); sn[c] = s2; }
        }
      }
      v = rr < 3 ? sn[rr] : cs[rr - 3];
    } else if (col >= 94) {
      v = 1.0f;                                    // bias carrier columns
    }
    f[col & 7] = v;
    if ((col & 7) == 7) w.store(col >> 3, f);
  }
}

template <int F, class Writer>
__device__ __forceinline__ void encode_row_ipe(const float m[3], const float var[3], const Writer& w) {
  // The output order is [all sin columns | all cos columns]: the octave chain is walked once, the cos half is kept in
  // registers (3F values, compile-time indexed) and written when its columns come up.
  float f[8];
  float sn[3], cs[3], dm[3];
  float cosv[3 * F];
#pragma unroll
  for (int col = 0; col < 96; ++col) {
    float v = 0.f;
    if (col < 3 * F) {
      const int k = col / 3, c = col % 3;
      if (c == 0) {
        if (k % 5 == 0) {
#pragma unroll
          for (int i = 0; i < 3; ++i) sincosf(m[i] * (float)(1 << k), &sn[i], &cs[i]);
        } else {
#pragma unroll
          for (int i = 0; i < 3; ++i) { float s2 = 2.f * sn[i] * cs[i]; cs[i] = fmaf(-2.f * sn[i], sn[i], 1.f); sn[i] = s2; }
        }
        if (k % 4 == 0) {
#pragma unroll
          for (int i = 0; i < 3; ++i) dm[i] = expf(-0.5f * (var[i] * (float)(1 << k) * (float)(1 << k)));
        } else {
#pragma unroll
          for (int i = 0; i < 3; ++i) { float d2 = dm[i] * dm[i]; dm[i] = d2 * d2; }
        }
      }
      v = dm[c] * sn[c];
      cosv[col] = dm[c] * cs[c];
    } else if (col < 6 * F) {
      v = cosv[col - 3 * F];
    } else if (col >= 94) {
      v = 1.0f;                                    // bias carrier columns
    }
    f[col & 7] = v;
    if ((col & 7) == 7) w.store(col >> 3, f);
  }
}

// One 32-column slice of a trunk-layer epilogue.  The bias is already in the accumulator, so a slice is 16 ReLU-fused
// bf16x2 conversions (F2FP.RELU) and ONE tcgen05.st: the packed activations go back into TMEM, over the first 16 of the
// slice's own 32 accumulator columns, where the next layer's MMAs read them as their A operand (no shared-memory
// round trip).  Only the tapped layer (stop_layer) also stages its slice in shared memory for the bulk store home.
static_assert(kTcMaskWordsPerTile == kSteps * 8 * kTileM, "mask layout");
struct EpiCtx {
  int64_t tile;                        // tile index of the row (ReLU-mask output)
  uint32_t tap_st; int r; int s;       // tap_st: this row's 64-byte segment in the warp's staging buffer 0
  uint32_t tap_bars; uint32_t tap_n;   // this warp's TAPFULL barrier pair; running count of staged chunks
  int* error_flag;
};
// Tapped layer (stop_layer): after the hand-over, the packed slice is read back from TMEM (it stays valid until the layer
// after next starts) and staged - the row's 64-byte segment, same swizzle as the UMMA tiles - for the tap warps, which
// stream it home with coalesced stores.  One staging buffer per chunk of the tile: nothing here waits in steady state.
__device__ __forceinline__ void tap_chunk(EpiCtx& e, uint32_t t_chunk) {
  uint32_t w[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]),
        "=r"(w[8]), "=r"(w[9]), "=r"(w[10]), "=r"(w[11]), "=r"(w[12]), "=r"(w[13]), "=r"(w[14]), "=r"(w[15])
      : "r"(t_chunk) : "memory");
  const uint32_t b = e.tap_n % kTapBufs;
  mbar_wait(e.tap_bars + 8 * (8 * kTapBufs + b), ((e.tap_n / kTapBufs) & 1u) ^ 1u, e.error_flag, 500);   // TAPFREE
  tmem_ld_wait();
#pragma unroll
  for (int c = 0; c < 4; ++c)
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(e.tap_st + b * kTapBytes + ((c ^ ((e.r & 7) >> 1)) << 4)),
                 "r"(w[4 * c]), "r"(w[4 * c + 1]), "r"(w[4 * c + 2]), "r"(w[4 * c + 3]) : "memory");
  __syncwarp();
  if ((threadIdx.x & 31) == 0) mbar_arrive(e.tap_bars + 8 * b);                            // TAPFULL
  ++e.tap_n;
}

template <int F16>
__device__ __forceinline__ void epi_trunk(EpiCtx& e, const uint32_t* v, uint32_t t_chunk, int j, const TcParams& p, bool valid,
                                          int64_t pt) {
  uint32_t w[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    if (F16) asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(w[i]) : "f"(__uint_as_float(v[2 * i + 1])), "f"(__uint_as_float(v[2 * i])));
    else asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(w[i]) : "f"(__uint_as_float(v[2 * i + 1])), "f"(__uint_as_float(v[2 * i])));
  }
  tmem_st16(t_chunk, w);
  if (p.masks) {                       // which of the 32 activations are non-zero: the backward pass's ReLU mask
    uint32_t m = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) m |= ((w[i] & 0xffffu) ? 1u : 0u) << (2 * i) | ((w[i] >> 16) ? 2u : 0u) << (2 * i);
    p.masks[(e.tile * kSteps + e.s) * (8 * kTileM) + j * kTileM + e.r] = m;
  }
  if (p.debug_out && e.s == p.debug_step && valid) {
#pragma unroll
    for (int i = 0; i < 32; ++i) p.debug_out[pt * kHid + j * 32 + i] = fmaxf(__uint_as_float(v[i]), 0.f);
  }
}

__device__ __forceinline__ uint64_t pk2(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ uint64_t pk2u(uint32_t a, uint32_t b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ void upk2(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) { uint64_t r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

// views_linears.0 epilogue (run by the encoder warps so the trunk epilogue warps can move on to the next tile):
// four 32-column slices: per-ray view bias + ReLU, rgb_linear + sigmoid, raw [P,4] store.
// Slices [j0, j0 + NS) of the 128 views columns; returns this warp's partial rgb dot products.
// Math of one 32-column slice jj of the views layer, accumulator values already in registers.
__device__ __forceinline__ void epi_view_slice(const float* tab, const float* vbrow, const uint32_t* v, int jj, uint64_t& q0, uint64_t& q1,
                                               uint64_t& q2, const TcParams& p, bool valid, int64_t pt, int64_t tile, int r) {
  float x[32];
  const float4* w0 = reinterpret_cast<const float4*>(tab + jj * 32);
  const float4* w1 = w0 + kHalf / 4;
  const float4* w2 = w1 + kHalf / 4;
#pragma unroll
  for (int i = 0; i < 8; ++i) {                // packed fp32x2: bias add (FADD2), rgb dot products (FFMA2)
    float4 t = vbrow ? reinterpret_cast<const float4*>(vbrow + jj * 32)[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    upk2(add2(pk2u(v[4 * i], v[4 * i + 1]), pk2(t.x, t.y)), x[4 * i], x[4 * i + 1]);
    upk2(add2(pk2u(v[4 * i + 2], v[4 * i + 3]), pk2(t.z, t.w)), x[4 * i + 2], x[4 * i + 3]);
    x[4 * i] = fmaxf(x[4 * i], 0.f); x[4 * i + 1] = fmaxf(x[4 * i + 1], 0.f);
    x[4 * i + 2] = fmaxf(x[4 * i + 2], 0.f); x[4 * i + 3] = fmaxf(x[4 * i + 3], 0.f);
    const uint64_t xa = pk2(x[4 * i], x[4 * i + 1]), xb = pk2(x[4 * i + 2], x[4 * i + 3]);
    float4 a = w0[i], b = w1[i], c = w2[i];
    q0 = fma2(xb, pk2(a.z, a.w), fma2(xa, pk2(a.x, a.y), q0));
    q1 = fma2(xb, pk2(b.z, b.w), fma2(xa, pk2(b.x, b.y), q1));
    q2 = fma2(xb, pk2(c.z, c.w), fma2(xa, pk2(c.x, c.y), q2));
  }
  if (p.masks) {
    uint32_t m = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) m |= (x[i] > 0.f ? 1u : 0u) << i;
    p.masks[(tile * kSteps + kViewStep) * (8 * kTileM) + jj * kTileM + r] = m;
  }
  if (p.debug_out && p.debug_step == kViewStep && valid) {
#pragma unroll
    for (int i = 0; i < 32; ++i) p.debug_out[pt * kHid + jj * 32 + i] = x[i];
  }
}

// ---- the kernel ----------------------------------------------------------------------------
// CG = 1: every CTA is on its own.  CG = 2: CTA pairs (cluster of 2) share every MMA (cta_group::2): each CTA keeps its own
// 128-point tile, encoder, epilogues and TMEM, but streams only HALF of every weight chunk (its half of the output
// columns) - the per-SM weight delivery from L2, which bounds the single-CTA kernel, is halved.
template <int CG, int F16>
__global__ void __launch_bounds__(kThreads, 1) nerf_mlp_tc_kernel(const __grid_constant__ TcParams p) {
  // Weight ring.  A stage costs a fixed ~310-400 cycles of round trip (tcgen05.commit -> empty barrier -> producer ->
  // cp.async.bulk -> full barrier) whatever its size: with one 16 KB chunk (256 cycles of MMAs) per stage the ring, not
  // the tensor core, set the pace - 76 stages x ~320 cycles = the whole 24 K-cycle tile (tools/mma_issue.cu: the bare loop
  // reaches 63 % of the MMA rate with 16 KB stages and 95 % with 32 KB stages).  Single CTAs therefore move TWO consecutive
  // chunks per stage: one bulk copy, one barrier pair, 512 cycles of MMAs; three 32 KB stages.  (A tile has 76 chunks = 38
  // pairs; the 8 views chunks are 10 KB each and pair up among themselves.)  CTA pairs keep 12 stages of 8 KB half chunks.
  constexpr int kCps = CG == 2 ? 1 : 2;                 // chunks per stage
  constexpr int kStages = CG == 2 ? kMaxStages : 3;
  constexpr int kStageB = kRingBytes / kStages;         // 8 KB half chunks | 32 KB chunk pairs
  const uint32_t rank = CG == 2 ? cluster_rank() : 0u;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  uint8_t* gbase = smem_raw;
  const uint32_t bars = base + OFF_BAR;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gbase + OFF_BAR + BAR_COUNT * 8);
  float* tab = reinterpret_cast<float*>(gbase + OFF_TAB);
  float4* partv = reinterpret_cast<float4*>(gbase + OFF_PARTV);
  float* vbs = reinterpret_cast<float*>(gbase + OFF_VBS);
  const bool use_vbs = p.vb_div >= 43;             // a tile then spans at most kVbRays rays
  // Tile boundary: the trunk epilogue warps are busy with the views layer of tile i exactly when layer 0 of tile i+1
  // completes, so the (otherwise idle) encoder warps run that one epilogue - unless layer 0 is the tapped layer.
  const bool enc_step0 = !(p.want_feat && p.stop_step == 0);
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // warp index: provably warp-uniform
  const int G = gridDim.x;
  if (base & 1023u) {                              // the swizzled UMMA tiles need a 1024-byte aligned window
    if (threadIdx.x == 0) { atomicExch(p.error_flag, 900); __threadfence_system(); }
    __trap();
  }
  for (int i = threadIdx.x; i < 3 * kHalf; i += kThreads) tab[i] = p.wrgb[i];

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(bars + 8 * (BAR_WFULL + i), 1); mbar_init(bars + 8 * (BAR_WEMPTY + i), 1); mbar_init(bars + 8 * (BAR_PFULL + i), 1);
    }
    for (int i = 0; i < 2; ++i) {                    // arrivals are one elected lane per warp, from both CTAs of a pair
      mbar_init(bars + 8 * (BAR_ACCFULL + i), 1);
      mbar_init(bars + 8 * (BAR_EFULL + i), 4 * CG);
      mbar_init(bars + 8 * (BAR_EEMPTY + i), 1);
      mbar_init(bars + 8 * (BAR_VEMPTY + i), 8 * CG);
    }
    mbar_init(bars + 8 * BAR_VFULL, 1);
    mbar_init(bars + 8 * BAR_VPART, 8);
    for (int c = 0; c < 8; ++c) mbar_init(bars + 8 * (BAR_HREADY + c), 4 * CG);
    for (int c = 0; c < 16 * kTapBufs; ++c) mbar_init(bars + 8 * (BAR_TAPFULL + c), 1);
    mbar_init(bars + 8 * BAR_TAPREAD, 8 * CG);
    mbar_init(bars + 8 * BAR_S0FULL, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    if (CG == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(bars + BAR_COUNT * 8), "r"(512));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(bars + BAR_COUNT * 8), "r"(512));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();                   // the peer's barriers are initialised before anything arrives on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== weight producer: bulk-copy chunks global -> smem ring =====
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int it = 0; it < p.n_iters; ++it) {
        const uint8_t* src = p.chunks;
        if (CG == 2) {
          for (int s = 0; s < kSteps; ++s) {
            const int nch = step_chunks_e(s) + step_chunks_h(s) + step_bias_chunk(s);
            const uint32_t bytes = step_chunk_bytes(s) / 2;       // this CTA's half of the chunk's output rows
            for (int c = 0; c < nch; ++c) {
              mbar_wait(bars + 8 * (BAR_WEMPTY + stage), phase ^ 1, p.error_flag, 100 + stage);
              mbar_arrive_expect_tx(bars + 8 * (BAR_WFULL + stage), bytes);
              bulk_g2s(base + OFF_W + stage * kStageB, src + rank * bytes, bytes, bars + 8 * (BAR_WFULL + stage));
              src += bytes * 2;
              if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
          }
        } else {
          for (int ci = 0; ci < kTileChunks; ci += 2) {           // chunks ci, ci + 1: contiguous in the stream, one copy
            const uint32_t bytes = tile_chunk_bytes(ci) + tile_chunk_bytes(ci + 1);
            mbar_wait(bars + 8 * (BAR_WEMPTY + stage), phase ^ 1, p.error_flag, 100 + stage);
            mbar_arrive_expect_tx(bars + 8 * (BAR_WFULL + stage), bytes);
            bulk_g2s(base + OFF_W + stage * kStageB, src, bytes, bars + 8 * (BAR_WFULL + stage));
            src += bytes;
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (CG == 2 && warp == 3 && rank == 1) {
    // ===== peer CTA: tell the leader's MMA thread when this CTA's half of a weight stage has landed =====
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int it = 0; it < p.n_iters; ++it)
        for (int s = 0; s < kSteps; ++s) {
          const int nch = step_chunks_e(s) + step_chunks_h(s) + step_bias_chunk(s);
          for (int c = 0; c < nch; ++c) {
            mbar_wait(bars + 8 * (BAR_WFULL + stage), phase, p.error_flag, 120 + stage);
            mbar_arrive_cluster(mapa_rank0(bars + 8 * (BAR_PFULL + stage)));
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
        }
    }
  } else if (warp == 1 && rank == 0) {
    // ===== MMA issuer (leader CTA of a pair issues for both) =====
    // The WHOLE warp runs this loop converged and one elected lane issues: every operand is then warp-uniform and
    // stays in uniform registers.  (Inside an `if (lane == 0)` region the compiler has to move each descriptor with
    // R2UR and wrap every UTCHMMA / UTCBAR in an ELECT ... BRA.U.ANY loop - ~110 instructions per K=32 chunk on one
    // thread, which made this thread, not the tensor core, the bottleneck: tools/mma_rate.cu, scripts/tc_timeline.py.)
    int stage = 0; uint32_t wphase = 0, hphase = 0;
    uint32_t ph_efull = 0;                                       // bit eb: phase of E buffer eb (bitmask: no local-memory array)
    int gstep = 0;                                               // running layer count: 9 steps per tile is odd, so the
    const uint32_t w_desc0 = desc_lo(base + OFF_W), e_desc0 = desc_lo(base + OFF_E);
    const bool tl_on = p.timeline != nullptr && blockIdx.x == 0 && lane == 0;
    const int tap_s2 = (p.stop_step + 2) % kSteps, tap_dt = (p.stop_step + 2) / kSteps;   // step (and tile offset) that reuses the tapped buffer
    for (int it = 0; it < p.n_iters; ++it) {                     // accumulator ping-pong must carry across tiles
      const int eb = it & 1;
      const uint32_t e_desc = e_desc0 + eb * (kEBytes >> 4);
#pragma unroll
      for (int s = 0; s < kSteps; ++s, ++gstep) {
        const int ne = step_chunks_e(s), nh = step_chunks_h(s);
        const uint32_t idesc = umma_idesc<F16>((s == kViewStep) ? kViewN : 256, kTileM * CG);
        const int ab = gstep & 1;
        const uint32_t d_tmem = tmem_base + ab * 256;            // accumulator ping-pong by layer
        const uint32_t a_tmem = tmem_base + (ab ^ 1) * 256;      // previous layer's activations, packed in place of its accumulator
        if (s == 0) {
          mbar_wait(bars + 8 * (BAR_EFULL + eb), (ph_efull >> eb) & 1u, p.error_flag, 210 + eb);
          ph_efull ^= 1u << eb;
        }
        if (p.want_feat && s == tap_s2 && it >= tap_dt)          // this step overwrites the tapped layer's buffer: its read-back is done
          mbar_wait(bars + 8 * BAR_TAPREAD, (uint32_t)(it - tap_dt) & 1u, p.error_flag, 260);
        if (tl_on && it < 8) p.timeline[(it * kSteps + s) * 4 + 0] = clock64();
        uint32_t accumulate = 0;
        const int ci0 = step_first_chunk(s);                     // position of this step's chunks in the tile's sequence: chunk ci lives
        if (step_bias_chunk(s)) {                                // in stage slot ci % kCps (compile-time after unrolling)
          // + bias: ones columns of the encoded tile x (lo, hi) bias rows.
          // Issued FIRST so the tensor core has work while the epilogue hands over chunk 0.  It overwrites the other
          // accumulator buffer, last read by the epilogue two layers back: for s >= 2 the previous layer's hand-over
          // waits already ordered that; for s == 1 it is the previous TILE's view epilogue, which signals VEMPTY.
          const bool first = ci0 % kCps == 0, last = ci0 % kCps == kCps - 1;
          const uint32_t boff = first ? 0u : (tile_chunk_bytes(ci0 - 1) >> 4);
          if (s == 1 && it > 0) mbar_wait(bars + 8 * (BAR_VEMPTY + ((it - 1) & 1)), ((it - 1) >> 1) & 1, p.error_flag, 240);
          if (first) {
            mbar_wait(bars + 8 * (BAR_WFULL + stage), wphase, p.error_flag, 220 + stage);
            if (CG == 2) mbar_wait(bars + 8 * (BAR_PFULL + stage), wphase, p.error_flag, 250 + stage);
          }
          tc_fence_after();
          if (elect_one()) {
            mma_ss<CG>(d_tmem, mk_desc(e_desc + 2 * (kBlk >> 4) + 2), mk_desc(w_desc0 + stage * (kStageB >> 4) + boff + 2), idesc, 0u);
            if (last) mma_commit<CG>(bars + 8 * (BAR_WEMPTY + stage));
          }
          __syncwarp();
          if (last) { if (++stage == kStages) { stage = 0; wphase ^= 1; } }
          accumulate = 1;
        }
#pragma unroll
        for (int c = 0; c < ne + nh; ++c) {
          const int ci = ci0 + step_bias_chunk(s) + c;
          const bool first = ci % kCps == 0, last = ci % kCps == kCps - 1;
          const uint32_t boff = first ? 0u : (tile_chunk_bytes(ci - 1) >> 4);
          const uint32_t wbar = bars + 8 * (BAR_WFULL + stage), pbar = bars + 8 * (BAR_PFULL + stage);
          if (c < ne) {
            if (first) {
              mbar_wait(wbar, wphase, p.error_flag, 220 + stage);
              if (CG == 2) mbar_wait(pbar, wphase, p.error_flag, 250 + stage);
            }
          } else {
            // chunk (c-ne) of the previous layer's output: packed by the epilogue into that layer's accumulator buffer
            const uint32_t hbar = bars + 8 * (BAR_HREADY + (c - ne));
            const bool ok_h = mbar_try_wait(hbar, hphase), ok_w = first ? mbar_try_wait(wbar, wphase) : true;   // all probes in flight
            const bool ok_p = (CG == 2 && first) ? mbar_try_wait(pbar, wphase) : true;
            if (!ok_h) mbar_wait_slow(hbar, hphase, p.error_flag, 230 + (c - ne));
            if (!ok_w) mbar_wait_slow(wbar, wphase, p.error_flag, 220 + stage);
            if (!ok_p) mbar_wait_slow(pbar, wphase, p.error_flag, 250 + stage);
          }
          tc_fence_after();
          const uint32_t b0 = w_desc0 + stage * (kStageB >> 4) + boff;
          if (elect_one()) {
            if (c < ne) {
              const uint32_t a0 = e_desc + c * (kBlk >> 4);
              mma_ss<CG>(d_tmem, mk_desc(a0), mk_desc(b0), idesc, accumulate);
              mma_ss<CG>(d_tmem, mk_desc(a0 + 2), mk_desc(b0 + 2), idesc, 1u);
            } else {                                               // K = 16 bf16 = 8 TMEM columns per MMA
              mma_ts<CG>(d_tmem, a_tmem + (c - ne) * 32, mk_desc(b0), idesc, accumulate);
              mma_ts<CG>(d_tmem, a_tmem + (c - ne) * 32 + 8, mk_desc(b0 + 2), idesc, 1u);
            }
            if (last) mma_commit<CG>(bars + 8 * (BAR_WEMPTY + stage));   // stage free (in both CTAs) once these MMAs retire
          }
          __syncwarp();
          accumulate = 1;
          if (last) { if (++stage == kStages) { stage = 0; wphase ^= 1; } }
        }
        if (tl_on && it < 8) p.timeline[(it * kSteps + s) * 4 + 1] = clock64();
        if (elect_one()) {
          if (s == 0 && enc_step0) {
            mma_commit<CG>(bars + 8 * BAR_S0FULL);                 // layer-0 accumulator complete -> encoder warps (one phase per tile)
          } else if (s < kViewStep) {
            mma_commit<CG>(bars + 8 * (BAR_ACCFULL + ab));         // accumulator complete -> trunk epilogue
          } else {
            mma_commit<CG>(bars + 8 * BAR_VFULL);                  // views accumulator -> view epilogue (encoder warps)
            mma_commit<CG>(bars + 8 * (BAR_EEMPTY + eb));          // encoded tile (incl. its ones columns) no longer needed
          }
        }
        __syncwarp();
        if (nh) hphase ^= 1;
      }
    }
  } else if (warp >= 4 && warp < 12) {
    // ===== epilogue warps: TMEM -> bias/activation -> bf16 A operand of the next layer =====
    const int g = (warp - 4) >> 2;                  // 0: even 32-column chunks, 1: odd chunks
    const int q = warp & 3;                         // TMEM lane quarter this warp may access
    const int r = q * 32 + lane;                    // row of the tile
    const int et = (warp - 4) * 32 + lane;          // 0..255
    uint32_t ph_acc = 0;                            // bit ab: phase of accumulator buffer ab
    uint32_t ph_vfull = 0;
    EpiCtx e;
    e.r = r;
    e.tap_st = base + OFF_TAP + (warp - 4) * kTapBufs * kTapBytes + ((lane >> 3) << 9) + ((lane & 7) << 6);   // this warp's staging buffers
    e.tap_bars = bars + 8 * (BAR_TAPFULL + (warp - 4) * kTapBufs); e.tap_n = 0; e.error_flag = p.error_flag;
    int gstep = 0;
    for (int it = 0; it < p.n_iters; ++it) {
      const int64_t tile = (int64_t)it * G + blockIdx.x;
      const int64_t pt = tile * kTileM + r;
      const bool valid = pt < p.P;
      e.tile = tile;
      for (int s = 0; s < kViewStep; ++s, ++gstep) {
        e.s = s;
        const int ab = gstep & 1;
        const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16) + ab * 256;
        if (s == 0 && enc_step0) continue;          // layer 0's epilogue belongs to the encoder warps (below; its own barrier)
        mbar_wait(bars + 8 * (BAR_ACCFULL + ab), (ph_acc >> ab) & 1u, p.error_flag, 300 + ab);
        ph_acc ^= 1u << ab;
        tc_fence_after();
        if (p.timeline && blockIdx.x == 0 && it < 8 && et == 0) p.timeline[(it * kSteps + s) * 4 + 2] = clock64();
        const bool tap = (s == p.stop_step) && p.want_feat && tile * kTileM < p.P;   // tapped layer: also goes home as tiled bf16 features
        auto hand_over = [&](int j) {
          tmem_st_wait();                            // packed activations are in TMEM ...
          tc_fence_before();                         // ... and ordered before the MMA thread's reads
          warp_arrive_leader<CG>(bars + 8 * (BAR_HREADY + j), rank, lane);
        };
        uint32_t va[32], vb[32];
        tmem_ld32(t_lane + g * 32, va);
#pragma unroll 1
        for (int jj = 0; jj < 4; jj += 2) {         // this warp's chunks j = g, g+2, g+4, g+6; software pipelined
          const int j = g + 2 * jj;
          tmem_ld_wait();
          tmem_ld32(t_lane + (j + 2) * 32, vb);
          epi_trunk<F16>(e, va, t_lane + j * 32, j, p, valid, pt);
          hand_over(j);
          tmem_ld_wait();
          if (jj + 2 < 4) tmem_ld32(t_lane + (j + 4) * 32, va);
          epi_trunk<F16>(e, vb, t_lane + (j + 2) * 32, j + 2, p, valid, pt);
          hand_over(j + 2);
        }
        tc_fence_before();
        if (p.timeline && blockIdx.x == 0 && it < 8 && et == 0) p.timeline[(it * kSteps + s) * 4 + 3] = clock64();
        if (s == p.stop_step && p.want_feat) {
          if (tap) {
#pragma unroll 1
            for (int jj = 0; jj < 4; ++jj) tap_chunk(e, t_lane + (g + 2 * jj) * 32);
          }
          tc_fence_before();                         // the layer after next may now overwrite this accumulator buffer
          warp_arrive_leader<CG>(bars + 8 * BAR_TAPREAD, rank, lane);
        }
      }
      {  // views layer: this warp takes slices g and g + 2 of the 128 columns; group 0 also fetches the sigma head's two
         // columns.  The encoder warps only combine the partial sums (they never touch TMEM), so the accumulator is
         // free again - and the next tile's layer 1 can start - as soon as these eight warps are through.
        const int ab = gstep & 1;
        const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16) + ab * 256;
        const float* vbrow = nullptr;
        if (valid) {
          const uint32_t ray = (uint32_t)pt / (uint32_t)p.vb_div;
          vbrow = use_vbs ? vbs + ((it & 1) * kVbRays + (int)(ray - (uint32_t)(tile * kTileM) / (uint32_t)p.vb_div)) * kHalf
                          : p.vbias + (size_t)ray * kHalf;
        }
        mbar_wait(bars + 8 * BAR_VFULL, ph_vfull, p.error_flag, 430);
        ph_vfull ^= 1;
        tc_fence_after();
        if (p.timeline && blockIdx.x == 0 && it < 8 && et == 0) p.timeline[(it * kSteps + kViewStep) * 4 + 2] = clock64();
        // Pull this warp's two slices (and the sigma head) into registers, release the accumulator at once - the next tile's
        // layer 1 overwrites it - and only then do the arithmetic.
        uint32_t va[32], vb[32], a0 = 0, a1 = 0;
        tmem_ld32(t_lane + g * 32, va);
        tmem_ld32(t_lane + (g + 2) * 32, vb);
        if (g == 0) asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(a0), "=r"(a1) : "r"(t_lane + kHalf) : "memory");
        tmem_ld_wait();
        tc_fence_before();
        warp_arrive_leader<CG>(bars + 8 * (BAR_VEMPTY + (it & 1)), rank, lane);
        uint64_t q0 = 0, q1 = 0, q2 = 0;                 // (even, odd) partial sums of the three rgb dot products
        epi_view_slice(tab, vbrow, va, g, q0, q1, q2, p, valid, pt, tile, r);
        epi_view_slice(tab, vbrow, vb, g + 2, q0, q1, q2, p, valid, pt, tile, r);
        float c0, c1, c2, d0, d1, d2;
        upk2(q0, c0, d0); upk2(q1, c1, d1); upk2(q2, c2, d2);
        // sigma = h7 . (w_hi + w_lo), accumulated in fp32 by the tensor core
        partv[((it & 1) * 2 + g) * kTileM + r] = make_float4(c0 + d0, c1 + d1, c2 + d2, __uint_as_float(a0) + __uint_as_float(a1));
        __syncwarp();
        if (lane == 0) mbar_arrive(bars + 8 * BAR_VPART);
        if (p.timeline && blockIdx.x == 0 && it < 8 && et == 0) p.timeline[(it * kSteps + kViewStep) * 4 + 3] = clock64();
      }
      ++gstep;
    }
  } else if (warp == 2 || (warp == 3 && CG == 1)) {
    // ===== tap warps: staged feature chunks (stop_layer output) -> tiled bf16 features in HBM, 512 contiguous bytes per store.
    // Single CTAs use warps 2 and 3 (one chunk group each); in a CTA pair warp 3 is the weight relay, so warp 2 drains both.
    const int ew0 = (CG == 1) ? (warp - 2) * 4 : 0, ew1 = (CG == 1) ? ew0 + 4 : 8;
    if (p.want_feat && !p.partials) {
      uint32_t n = 0;                                // chunks taken from each epilogue warp so far
      for (int it = 0; it < p.n_iters; ++it) {
        const int64_t tile = (int64_t)it * G + blockIdx.x;
        if (tile * kTileM >= p.P) break;
        uint8_t* tdst = p.feat_tiles + (size_t)tile * kFeatTileBytes + lane * 16;
        for (int jj = 0; jj < 4; ++jj, ++n) {
          const uint32_t b = n % kTapBufs, ph = (n / kTapBufs) & 1u;
#pragma unroll 1
          for (int ew = ew0; ew < ew1; ++ew) {       // epilogue warp ew = g * 4 + q staged chunk j = g + 2 jj of rows 32 q .. 32 q + 31
            const int j = (ew >> 2) + 2 * jj, q = ew & 3;
            mbar_wait(bars + 8 * (BAR_TAPFULL + ew * kTapBufs + b), ph, p.error_flag, 510 + ew);
            const uint32_t src = base + OFF_TAP + (ew * kTapBufs + b) * kTapBytes + lane * 16;
            uint4 v[4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v[i].x), "=r"(v[i].y), "=r"(v[i].z), "=r"(v[i].w) : "r"(src + i * 512));
#pragma unroll
            for (int i = 0; i < 4; ++i) *reinterpret_cast<uint4*>(tdst + j * kBlk + q * kTapBytes + i * 512) = v[i];
            __syncwarp();
            if (lane == 0) mbar_arrive(bars + 8 * (BAR_TAPFREE + ew * kTapBufs + b));
          }
        }
      }
    }
  } else if (warp >= 12) {
    // ===== encoder warps: (1) sampling data -> positional encoding -> bf16 A tile of layer 0 / skip, one tile ahead;
    //                      (2) the views-layer epilogue of the current tile (rgb + sigma -> raw) =====
    const int q = warp & 3;
    const int r = q * 32 + lane;
    uint32_t ph_eempty = 0, ph_vpart = 0, tapn = 0;
    // Returns the row's sample interval times |d| (negative at the first sample of a ray), which the tile-local
    // compositing of the same tile needs one iteration later (kept in a register: same thread, same row).
    auto encode_tile = [&](int it) -> float {
      const int eb = it & 1;
      const int64_t tile = (int64_t)it * G + blockIdx.x;
      const int64_t pt = tile * kTileM + r;
      const bool valid = pt < p.P;
      float m[3] = {0.f, 0.f, 0.f}, vv[3] = {0.f, 0.f, 0.f};
      const float* xrow = nullptr;
      float dist = 0.f;
      if (valid) {
        if (p.mode == 0) {
          const uint32_t ray = (uint32_t)pt / (uint32_t)p.S;   // points per launch < 2^31 (checked on the host): 32-bit division
          const int sidx = (int)((uint32_t)pt - ray * (uint32_t)p.S);
          const float* rr = p.rays + (size_t)ray * p.ncols;
          float d0, d1, d2;
          if (p.mip) {
            const float* z = p.z + (size_t)ray * (p.S + 1);
            float rl[12];
#pragma unroll
            for (int i = 0; i < 12; ++i) rl[i] = __ldg(rr + i);
            const float t0 = __ldg(z + sidx), t1 = __ldg(z + sidx + 1);
            frustum_gaussian(rl, t0, t1, p.var_scale, m, vv);
            dist = t1 - t0; d0 = rl[3]; d1 = rl[4]; d2 = rl[5];
          } else {
            const float zz = __ldg(p.z + pt);
            d0 = __ldg(rr + 3); d1 = __ldg(rr + 4); d2 = __ldg(rr + 5);
            m[0] = __ldg(rr) + d0 * zz; m[1] = __ldg(rr + 1) + d1 * zz; m[2] = __ldg(rr + 2) + d2 * zz;
            dist = (sidx < p.S - 1) ? (__ldg(p.z + pt + 1) - zz) : 1e10f;
          }
          dist = dist * sqrtf(d0 * d0 + d1 * d1 + d2 * d2);          // volume_render_radiance_field: delta * |rays_d|
          if (sidx == 0) dist = -dist;                               // marks the head of a ray segment (intervals are >= 0)
        } else {
          xrow = p.X + pt * p.kx;
        }
      }
      mbar_wait(bars + 8 * (BAR_EEMPTY + eb), ((ph_eempty >> eb) & 1u) ^ 1u, p.error_flag, 400 + eb);
      ph_eempty ^= 1u << eb;
      if (use_vbs) {                               // per-ray view bias of this tile's (<= 4) rays -> shared memory
        const int64_t ray = (int64_t)((uint32_t)(tile * kTileM) / (uint32_t)p.vb_div) + (r >> 5);   // (this buffer's last reader was our own view epilogue)
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ray * p.vb_div < p.P) t = __ldg(reinterpret_cast<const float4*>(p.vbias + ray * kHalf) + (r & 31));
        reinterpret_cast<float4*>(vbs + (eb * kVbRays + (r >> 5)) * kHalf)[r & 31] = t;
      }
      const RowWriter<F16> rw{base + OFF_E + eb * kEBytes, r};
      if (p.mode == 1 || !valid) {
        for (int ch = 0; ch < 12; ++ch) {
          float f[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) { const int col = ch * 8 + i; f[i] = (valid && col < p.xyz) ? __ldg(xrow + col) : (col >= 94 ? 1.f : 0.f); }
          rw.store(ch, f);
        }
      } else if (p.mip) {
        encode_row_ipe<15>(m, vv, rw);
      } else {
        encode_row_fourier<15>(m, rw);
      }
      fence_proxy_async();
      warp_arrive_leader<CG>(bars + 8 * (BAR_EFULL + eb), rank, lane);
      return dist;
    };
    // Layer-0 epilogue of tile t (see enc_step0): this warp's quarter of the rows, all eight 32-column chunks.
    auto step0_epilogue = [&](int t) {
      const int g0 = t * kSteps, ab = g0 & 1;                          // layer 0 of tile t is global step 9 t
      const int64_t pt = ((int64_t)t * G + blockIdx.x) * kTileM + r;
      EpiCtx e0;
      e0.tile = (int64_t)t * G + blockIdx.x;
      e0.r = r; e0.s = 0; e0.tap_st = 0; e0.tap_bars = 0; e0.tap_n = 0; e0.error_flag = p.error_flag;
      const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16) + ab * 256;
      mbar_wait(bars + 8 * BAR_S0FULL, (uint32_t)t & 1u, p.error_flag, 310 + ab);
      tc_fence_after();
      if (p.timeline && blockIdx.x == 0 && t < 8 && r == 0) p.timeline[(t * kSteps) * 4 + 2] = clock64();
      uint32_t va[32], vb[32];
      tmem_ld32(t_lane, va);
#pragma unroll 1
      for (int j = 0; j < 8; j += 2) {
        tmem_ld_wait();
        tmem_ld32(t_lane + (j + 1) * 32, vb);
        epi_trunk<F16>(e0, va, t_lane + j * 32, j, p, pt < p.P, pt);
        tmem_st_wait(); tc_fence_before();
        warp_arrive_leader<CG>(bars + 8 * (BAR_HREADY + j), rank, lane);
        tmem_ld_wait();
        if (j + 2 < 8) tmem_ld32(t_lane + (j + 2) * 32, va);
        epi_trunk<F16>(e0, vb, t_lane + (j + 1) * 32, j + 1, p, pt < p.P, pt);
        tmem_st_wait(); tc_fence_before();
        warp_arrive_leader<CG>(bars + 8 * (BAR_HREADY + j + 1), rank, lane);
      }
      if (p.timeline && blockIdx.x == 0 && t < 8 && r == 0) p.timeline[(t * kSteps) * 4 + 3] = clock64();
    };
    float dist_cur = encode_tile(0);
    if (enc_step0) step0_epilogue(0);
    const bool etl = p.timeline != nullptr && blockIdx.x == 0 && r == 0;      // encoder-side stamps of iterations 2..5
    for (int it = 0; it < p.n_iters; ++it) {
      if (etl && it >= 2 && it < 6) p.timeline[8 * kSteps * 4 + (it - 2) * 4 + 0] = clock64();
      float dist_next = 0.f;
      if (it + 1 < p.n_iters) dist_next = encode_tile(it + 1);
      if (etl && it >= 2 && it < 6) p.timeline[8 * kSteps * 4 + (it - 2) * 4 + 1] = clock64();
      if (enc_step0 && it + 1 < p.n_iters) step0_epilogue(it + 1);
      // ---- finish the views layer of tile `it`: add the eight trunk warps' partial sums, sigmoid, raw [P,4] store ----
      const int64_t tile = (int64_t)it * G + blockIdx.x;
      const int64_t pt = tile * kTileM + r;
      mbar_wait(bars + 8 * BAR_VPART, ph_vpart, p.error_flag, 440);      // the trunk warps' partial sums are in shared memory
      ph_vpart ^= 1;
      if (etl && it >= 2 && it < 6) p.timeline[8 * kSteps * 4 + (it - 2) * 4 + 2] = clock64();
      float sigma = 0.f;
      if (pt < p.P) {
        const float4 p0 = partv[((it & 1) * 2) * kTileM + r], p1 = partv[((it & 1) * 2 + 1) * kTileM + r];
        float4 o;
        o.x = 1.f / (1.f + __expf(-(p0.x + p1.x + p.brgb0)));
        o.y = 1.f / (1.f + __expf(-(p0.y + p1.y + p.brgb1)));
        o.z = 1.f / (1.f + __expf(-(p0.z + p1.z + p.brgb2)));
        o.w = p0.w + p.balpha;
        sigma = o.w;
        *reinterpret_cast<float4*>(p.raw + pt * 4) = o;
      }
      if (p.partials && p.want_feat) {
        // ---- local compositing weights of the tile (volume_render_radiance_field, render_utils.py:176-230, restricted to the
        //      tile): alpha_r and the exclusive product of (1 - alpha + 1e-10) over the earlier rows of the SAME ray ----
        float alpha = 0.f, f = 1.f;
        bool head = (lane == 0 && q == 0);
        if (pt < p.P) {
          head = head || dist_cur < 0.f || __float_as_uint(dist_cur) == 0x80000000u;   // (-0.0: zero-length first interval)
          alpha = 1.0f - expf(-fmaxf(sigma, 0.0f) * fabsf(dist_cur));
          f = 1.0f - alpha + 1e-10f;
        }
        float v = f; int flag = head ? 1 : 0;                   // segmented inclusive product scan inside the warp
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
          const float uv = __shfl_up_sync(0xffffffffu, v, off);
          const int uf = __shfl_up_sync(0xffffffffu, flag, off);
          if (lane >= off) { if (!flag) v *= uv; flag |= uf; }
        }
        float* scan = reinterpret_cast<float*>(gbase + OFF_WL + 2 * kTileM * 4) + (it & 1) * 8;
        if (lane == 31) { scan[2 * q] = v; scan[2 * q + 1] = __int_as_float(flag); }
        asm volatile("bar.sync 1, 128;" ::: "memory");              // the four encoder warps
        float carry = 1.f;                                         // product reaching this warp from the rows before it
        for (int w = 0; w < q; ++w) carry = __float_as_int(scan[2 * w + 1]) ? scan[2 * w] : carry * scan[2 * w];
        const float incl = flag ? v : carry * v;
        const float up = __shfl_up_sync(0xffffffffu, incl, 1);
        const float excl = head ? 1.f : (lane == 0 ? carry : up);
        float* wl = reinterpret_cast<float*>(gbase + OFF_WL) + (it & 1) * kTileM;
        wl[r] = alpha * excl;
        asm volatile("bar.sync 1, 128;" ::: "memory");
        // ---- F[seg][c] = sum_r wl_r f[r][c] per ray segment of the tile.  The tile's tapped features have been sitting in the
        //      staging area since layer stop_layer; warp q reduces chunks 2q and 2q + 1 (staging buffer q of all eight
        //      epilogue warps).  The compositor then only scales each segment by the transmittance in front of it - the
        //      per-sample features never go to HBM. ----
        const int64_t pt0 = tile * kTileM;
        if (pt0 < p.P) {
          const int rows = (int)((p.P - pt0 < kTileM) ? (p.P - pt0) : kTileM);
          const int64_t ray0 = (uint32_t)pt0 / (uint32_t)p.S;
          const int nseg = (int)((uint32_t)(pt0 + rows - 1) / (uint32_t)p.S - (uint32_t)ray0) + 1;
          const int pc = lane & 3, rsub = lane >> 2;               // 16-byte piece (8 channels) of the row's 64 B; row = 8 i + rsub
          for (int ew = 0; ew < 8; ++ew) mbar_wait(bars + 8 * (BAR_TAPFULL + ew * kTapBufs + q), tapn & 1u, p.error_flag, 510 + ew);
#pragma unroll 1
          for (int k = 0; k < nseg; ++k) {
            const int64_t a64 = (ray0 + k) * p.S - pt0, b64 = a64 + p.S;
            const int ra = a64 < 0 ? 0 : (int)a64, rb = b64 > rows ? rows : (int)b64;
            float acc[2][8];
#pragma unroll
            for (int gg = 0; gg < 2; ++gg)
#pragma unroll
              for (int c = 0; c < 8; ++c) acc[gg][c] = 0.f;
            for (int qq = ra >> 5; qq <= (rb - 1) >> 5; ++qq) {
              const uint32_t unit = base + OFF_TAP + (qq * kTapBufs + q) * kTapBytes;     // epilogue warp (gg, qq), buffer q
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int rr = 8 * i + rsub, row = qq * 32 + rr;
                const float w = (row >= ra && row < rb) ? (F16 ? wl[row] * kF16Scale : wl[row]) : 0.f;
                const uint32_t src = unit + rr * 64 + ((pc ^ ((rr & 7) >> 1)) << 4);
#pragma unroll
                for (int gg = 0; gg < 2; ++gg) {
                  uint32_t v0, v1, v2, v3;
                  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3)
                               : "r"(src + gg * 4 * kTapBufs * kTapBytes));
                  acc[gg][0] = fmaf(w, unpack_lo<F16>(v0), acc[gg][0]); acc[gg][1] = fmaf(w, unpack_hi<F16>(v0), acc[gg][1]);
                  acc[gg][2] = fmaf(w, unpack_lo<F16>(v1), acc[gg][2]); acc[gg][3] = fmaf(w, unpack_hi<F16>(v1), acc[gg][3]);
                  acc[gg][4] = fmaf(w, unpack_lo<F16>(v2), acc[gg][4]); acc[gg][5] = fmaf(w, unpack_hi<F16>(v2), acc[gg][5]);
                  acc[gg][6] = fmaf(w, unpack_lo<F16>(v3), acc[gg][6]); acc[gg][7] = fmaf(w, unpack_hi<F16>(v3), acc[gg][7]);
                }
              }
            }
#pragma unroll
            for (int gg = 0; gg < 2; ++gg) {
#pragma unroll
              for (int c = 0; c < 8; ++c) {
                float v = acc[gg][c];
                v += __shfl_xor_sync(0xffffffffu, v, 4); v += __shfl_xor_sync(0xffffffffu, v, 8); v += __shfl_xor_sync(0xffffffffu, v, 16);
                acc[gg][c] = v;
              }
              if (lane < 4) {                                      // lanes 0..3 hold the four 8-channel pieces of chunk gg + 2 q
                float4* dst = reinterpret_cast<float4*>(p.partials + ((size_t)tile * kSeg + k) * kHid + (gg + 2 * q) * 32 + pc * 8);
                dst[0] = make_float4(acc[gg][0], acc[gg][1], acc[gg][2], acc[gg][3]);
                dst[1] = make_float4(acc[gg][4], acc[gg][5], acc[gg][6], acc[gg][7]);
              }
            }
          }
          __syncwarp();
          if (lane == 0)
            for (int ew = 0; ew < 8; ++ew) mbar_arrive(bars + 8 * (BAR_TAPFREE + ew * kTapBufs + q));
          ++tapn;
        }
      }
      if (etl && it >= 2 && it < 6) p.timeline[8 * kSteps * 4 + (it - 2) * 4 + 3] = clock64();
      dist_cur = dist_next;
    }
  }

  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();                   // neither CTA leaves (or frees TMEM) while its peer can still reach it
  if (warp == 2) {
    tc_fence_after();
    if (CG == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// per-ray (or per-row) view bias: b_view + W_view[:, 256:] . [dirs encoding | appearance]   (fp32)
struct VbParams {
  int mode; const float* rays; int ncols; const int64_t* ray_id; const float* app_table; int num_frames;
  const float* X; int kx; int xyzp;
  int mip; int fd; int dirs; int app; int dap;
  const float* wview_d; const float* bview; int64_t n; float* out;
};
constexpr int kVbRows = 64;   // rays (rows) per block: the transposed weight panel is staged once per block
constexpr int kVbBatch = 8;   // rows per inner iteration: one weight load feeds eight independent FMA chains
constexpr int kVbMaxK = 80;   // largest dirs + appearance width nmr_create accepts (dirs_num_freqs 8 -> 51, + 16), padded to 16
__global__ void __launch_bounds__(128) view_bias_kernel(VbParams p) {
  extern __shared__ __align__(16) float vb_smem[];   // sized by the launch: dap x (8 + 129) floats
  float (*de)[kVbBatch] = reinterpret_cast<float (*)[kVbBatch]>(vb_smem);            // [k][row of the batch]: encoded dirs | appearance
  float (*wt)[129] = reinterpret_cast<float (*)[129]>(vb_smem + p.dap * kVbBatch);   // [k][output], transposed weight panel
  const int t = threadIdx.x;
  const int nk = p.dirs + p.app;
  for (int i = t; i < kHalf * p.dap; i += 128) { int o = i / p.dap, k = i - o * p.dap; wt[k][o] = p.wview_d[i]; }
  const float bv = p.bview[t];
  const int64_t row0 = (int64_t)blockIdx.x * kVbRows;
  for (int rb = 0; rb < kVbRows; rb += kVbBatch) {
    if (row0 + rb >= p.n) break;
    __syncthreads();                                   // previous batch consumed (first pass: weight panel staged)
    for (int i = t; i < p.dap * kVbBatch; i += 128) {  // element (k, rr) of the batch
      const int k = i / kVbBatch, rr = i - k * kVbBatch;
      const int64_t row = row0 + rb + rr;
      float v = 0.f;
      if (row < p.n) {
        if (p.mode == 1) {
          v = p.X[row * p.kx + p.xyzp + k];
        } else if (k < p.dirs) {
          const float* ry = p.rays + row * p.ncols;
          const float* vd = (p.ncols >= 11) ? ry + 8 : ry + 3;
          float v3[3] = {vd[0], vd[1], vd[2]};
          v = p.mip ? pe_col(v3, p.fd, k) : fourier_col(v3, p.fd, k);
        } else if (k < nk) {
          int64_t id = p.ray_id ? p.ray_id[row] : 1;
          id = id < 0 ? 0 : (id >= p.num_frames ? p.num_frames - 1 : id);
          v = p.app_table[id * p.app + (k - p.dirs)];
        }
      }
      de[k][rr] = v;
    }
    __syncthreads();
    float acc[kVbBatch];
#pragma unroll
    for (int rr = 0; rr < kVbBatch; ++rr) acc[rr] = 0.f;
    for (int k = 0; k < nk; ++k) {                     // same k order as before: bit-identical sums
      const float w = wt[k][t];
      const float4 d0 = *reinterpret_cast<const float4*>(&de[k][0]), d1 = *reinterpret_cast<const float4*>(&de[k][4]);
      acc[0] = fmaf(d0.x, w, acc[0]); acc[1] = fmaf(d0.y, w, acc[1]); acc[2] = fmaf(d0.z, w, acc[2]); acc[3] = fmaf(d0.w, w, acc[3]);
      acc[4] = fmaf(d1.x, w, acc[4]); acc[5] = fmaf(d1.y, w, acc[5]); acc[6] = fmaf(d1.z, w, acc[6]); acc[7] = fmaf(d1.w, w, acc[7]);
    }
#pragma unroll
    for (int rr = 0; rr < kVbBatch; ++rr) {
      const int64_t row = row0 + rb + rr;
      if (row < p.n) p.out[row * kHalf + t] = acc[rr] + bv;
    }
  }
}

inline size_t vb_smem_bytes(int dap) { return (size_t)dap * (kVbBatch + 129) * sizeof(float); }
static_assert(kVbMaxK * (kVbBatch + 129) * sizeof(float) <= 48 * 1024, "view_bias_kernel: dynamic shared memory without opt-in");

// tiled bf16 features -> fp32 [P,256] rows (stage API only)
__global__ void detile_kernel(const uint8_t* tiles, int64_t P, float* out, int f16) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // one 16-byte chunk (8 channels) per thread
  if (i >= P * 32) return;
  int64_t pt = i >> 5;
  int l = (int)(i & 31);
  int j = l >> 2, c = l & 3;
  int r = (int)(pt & 127);
  const uint8_t* src = tiles + (pt >> 7) * kFeatTileBytes + j * kBlk + sw64_off(r, c);
  uint4 v = *reinterpret_cast<const uint4*>(src);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
  const __half2* hh = reinterpret_cast<const __half2*>(&v);
  float* o = out + pt * kHid + j * 32 + c * 8;
#pragma unroll
  for (int k = 0; k < 4; ++k) { float2 f = f16 ? __half22float2(hh[k]) : __bfloat1622float2(h[k]); o[2 * k] = f.x; o[2 * k + 1] = f.y; }
}

// ---- host side -----------------------------------------------------------------------------
uint16_t f2bf(float f) {
  uint32_t u; memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
// fp32 -> fp16 bits, round to nearest even, subnormals kept, finite overflow saturates at 65504 (like cvt.rn.satfinite)
uint16_t f2h(float f) {
  uint32_t u; memcpy(&u, &f, 4);
  const uint16_t sign = (uint16_t)((u >> 16) & 0x8000u);
  const uint32_t a = u & 0x7fffffffu;
  if (a > 0x7f800000u) return (uint16_t)(sign | 0x7e00u);            // NaN
  if (a >= 0x477ff000u) return (uint16_t)(sign | 0x7bffu);           // >= 65520 rounds beyond the largest finite value: saturate
  if (a < 0x33000001u) return sign;                                  // <= 2^-25: rounds to zero
  const int e = (int)(a >> 23) - 127;
  uint32_t m = (a & 0x7fffffu) | 0x800000u;                          // 24-bit significand
  int shift = e < -14 ? (13 + (-14 - e)) : 13;                       // bits dropped (subnormal results drop more)
  uint32_t q = m >> shift, rem = m & ((1u << shift) - 1u), half = 1u << (shift - 1);
  if (rem > half || (rem == half && (q & 1u))) ++q;
  const uint32_t base = e < -14 ? 0u : ((uint32_t)(e + 15) << 10) - 0x400u;   // q carries the implicit bit for normals
  return (uint16_t)(sign | (uint16_t)(base + q));
}
float h2f(uint16_t h) {
  const uint32_t sign = (uint32_t)(h & 0x8000u) << 16, e = (h >> 10) & 0x1fu, m = h & 0x3ffu;
  float v;
  if (e == 0) v = ldexpf((float)m, -24);
  else if (e == 31) v = m ? NAN : INFINITY;
  else v = ldexpf((float)(m | 0x400u), (int)e - 25);
  return sign ? -v : v;
}
// the 16-bit operand format of one packed net: 0 = bf16, 1 = fp16
inline uint16_t to16(float f, int fmt) { return fmt ? f2h(f) : f2bf(f); }
inline float from16(uint16_t h, int fmt) {
  if (fmt) return h2f(h);
  uint32_t u = (uint32_t)h << 16; float f; memcpy(&f, &u, 4); return f;
}

// pack W[rows, ldw] columns [k0, k0+32) (zero beyond kmax) as one swizzled chunk
void pack_chunk(std::vector<uint8_t>& out, const float* W, int rows, int ldw, int k0, int kmax, int fmt) {
  size_t off = out.size();
  out.resize(off + (size_t)rows * 64, 0);
  for (int n = 0; n < rows; ++n)
    for (int k = 0; k < 32; ++k) {
      int kk = k0 + k;
      float v = (kk < kmax) ? W[(size_t)n * ldw + kk] : 0.f;
      uint16_t h = to16(v, fmt);
      size_t b = off + (size_t)(n >> 3) * 512 + (n & 7) * 64 + (((k >> 3) ^ ((n & 7) >> 1)) << 4) + (k & 7) * 2;
      memcpy(&out[b], &h, 2);
    }
}

struct TcWs {   // carved from one device allocation
  float* raw; uint8_t* feat; float* vbias; int* error_flag; float* xfeat; float* partials;
};
// with_feat: reserve the tiled per-sample feature buffer (512 B per point; only the unfused tap needs it)
size_t ws_layout(int64_t points, int64_t vb_rows, TcWs* w, uint8_t* basep, bool with_feat = true) {
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 1023) / 1024 * 1024; return o; };
  int64_t tiles = (points + kTileM - 1) / kTileM + 2;
  size_t o_raw = take((size_t)tiles * kTileM * 4 * sizeof(float));
  size_t o_feat = take(with_feat ? (size_t)tiles * kFeatTileBytes : 1024);
  size_t o_vb = take((size_t)std::max<int64_t>(vb_rows, 1) * kHalf * sizeof(float));
  size_t o_err = take(1024);
  size_t o_part = take((size_t)tiles * kSeg * kHid * sizeof(float));
  if (w && basep) {
    w->raw = (float*)(basep + o_raw); w->feat = basep + o_feat; w->vbias = (float*)(basep + o_vb);
    w->error_flag = (int*)(basep + o_err); w->partials = (float*)(basep + o_part);
  }
  return off;
}

int* g_trap_host = nullptr;   // mapped pinned int: survives a trapped launch (portable: one flag for every device)
int* g_trap_dev = nullptr;
long long* g_timeline = nullptr;

int g_cta_group = 0;           // 1 | 2 (CTA pairs); 0: not decided yet (NMR_TC_CG, default 1)
// One-time setup is per DEVICE: the opt-in shared-memory attribute of a kernel, the SM count and the number of
// co-schedulable CTA pairs all belong to the device that is current at launch (one process may drive several).
constexpr int kMaxDevices = 64;
struct DevInit { bool done = false; int num_sms = 0; int pair_ctas = 0; };
DevInit g_dev[kMaxDevices];

int launch_tc(const TcNet& net, TcParams& p, int* launches, cudaStream_t s, std::string* err) {
  if (p.P >= ((int64_t)1 << 31)) { *err = "more than 2^31 sample points in one launch: lower max_chunk_points"; return -1; }
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) { *err = "cudaGetDevice failed (or device ordinal >= 64)"; return -1; }
  if (!g_trap_host) {
    if (cudaHostAlloc((void**)&g_trap_host, sizeof(int), cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess ||
        cudaHostGetDevicePointer((void**)&g_trap_dev, g_trap_host, 0) != cudaSuccess) {
      *err = "cudaHostAlloc(trap flag) failed"; return -1;
    }
    *g_trap_host = 0;
    if (g_cta_group == 0) {
      const char* ev = getenv("NMR_TC_CG");
      g_cta_group = (ev && atoi(ev) == 2) ? 2 : 1;
    }
  }
  DevInit& di = g_dev[dev];
  if (!di.done) {
    cudaDeviceGetAttribute(&di.num_sms, cudaDevAttrMultiProcessorCount, dev);
    cudaError_t e = cudaFuncSetAttribute(nerf_mlp_tc_kernel<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(nerf_mlp_tc_kernel<2, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(nerf_mlp_tc_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(nerf_mlp_tc_kernel<2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e != cudaSuccess) { *err = std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e); return -1; }
    cudaLaunchConfig_t qc{};
    cudaLaunchAttribute qa[1];
    qa[0].id = cudaLaunchAttributeClusterDimension; qa[0].val.clusterDim.x = 2; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
    qc.gridDim = dim3(di.num_sms & ~1); qc.blockDim = dim3(kThreads); qc.dynamicSmemBytes = kSmemBytes; qc.attrs = qa; qc.numAttrs = 1;
    int ncl = 0;
    e = cudaOccupancyMaxActiveClusters(&ncl, nerf_mlp_tc_kernel<2, 0>, &qc);
    if (e != cudaSuccess || ncl < 1) { cudaGetLastError(); ncl = 0; }
    di.pair_ctas = 2 * ncl;
    di.done = true;
  }
  const int num_sms = di.num_sms;
  const int cta_group = (g_cta_group == 2 && di.pair_ctas >= 2) ? 2 : 1;   // a device that cannot co-schedule pairs runs single CTAs
  p.error_flag = g_trap_dev;
  if (!g_timeline && getenv("NMR_TC_TIMELINE")) { cudaMalloc((void**)&g_timeline, (kTlSlots) * sizeof(long long)); cudaMemset(g_timeline, 0, (kTlSlots) * sizeof(long long)); }
  p.timeline = g_timeline;
  int64_t tiles = (p.P + kTileM - 1) / kTileM;
  p.chunks = net.chunks; p.wrgb = net.wrgb;
  p.balpha = net.balpha; p.brgb0 = net.brgb[0]; p.brgb1 = net.brgb[1]; p.brgb2 = net.brgb[2];
  cudaError_t e;
  if (cta_group == 2) {
    int grid = (int)std::min<int64_t>(di.pair_ctas, (std::max<int64_t>(tiles, 1) + 1) / 2 * 2);
    p.n_iters = (int)((tiles + grid - 1) / grid);
    cudaLaunchConfig_t cfg{};
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = kSmemBytes; cfg.stream = s; cfg.attrs = at; cfg.numAttrs = 1;
    e = net.fmt ? cudaLaunchKernelEx(&cfg, nerf_mlp_tc_kernel<2, 1>, p) : cudaLaunchKernelEx(&cfg, nerf_mlp_tc_kernel<2, 0>, p);
  } else {
    int grid = (int)std::min<int64_t>(num_sms, std::max<int64_t>(tiles, 1));
    p.n_iters = (int)((tiles + grid - 1) / grid);
    if (net.fmt) nerf_mlp_tc_kernel<1, 1><<<grid, kThreads, kSmemBytes, s>>>(p);
    else nerf_mlp_tc_kernel<1, 0><<<grid, kThreads, kSmemBytes, s>>>(p);
    e = cudaGetLastError();
  }
  if (e != cudaSuccess) { *err = std::string("nerf_mlp_tc_kernel launch: ") + cudaGetErrorString(e); return -1; }
  *launches += 1;
  return 0;
}

}  // namespace

int tc_pack_net(const Dims& d, int skip_layer, const HostNet& hn, TcNet* out, std::string* err, int fmt) {
  const int skipl = skip_layer + 1;
  out->set = false;
  out->fmt = fmt;
  if (fmt) {   // fp16 operands: every weight must be representable (|w| <= 65504); return 1 -> this net stays on bf16 / fp32
    auto too_big = [](const std::vector<float>& v) { for (float x : v) if (!(fabsf(x) <= 65504.f)) return true; return false; };
    bool bad = too_big(hn.feature_w) || too_big(hn.views_w) || too_big(hn.alpha_w) || too_big(hn.rgb_w);
    for (int i = 0; i < kLayers; ++i) bad = bad || too_big(hn.pts_w[i]) || too_big(hn.pts_b[i]);
    if (bad) { *err = "a weight exceeds the fp16 range (65504): use bf16 or fp32 for this network"; return 1; }
  }
  if (skipl != 5 || d.xyzp != 96 || d.fx != 15 || d.xyz > 94 || d.dap > kVbMaxK) {   // return 1: configuration valid for the fp32 path only
    *err = "the tcgen05 path is built for skips=[4] and a 96-wide padded xyz encoding (15 frequencies)";
    return 1;
  }
  const int vin = kHid + d.dirs + d.app;
  // feature_linear has no activation, so views_linears.0(feature_linear(h)) = (Wv_f Wf) h + Wv_f b_f + ...:
  // fold the two on the host in double precision (models/nerf.py:110-114).
  std::vector<float> wfold((size_t)kHalf * kHid);
  std::vector<float> bfold(kHalf);
  for (int o = 0; o < kHalf; ++o) {
    for (int i = 0; i < kHid; ++i) {
      double acc = 0.0;
      for (int m = 0; m < kHid; ++m) acc += (double)hn.views_w[(size_t)o * vin + m] * (double)hn.feature_w[(size_t)m * kHid + i];
      wfold[(size_t)o * kHid + i] = (float)acc;
    }
    double accb = hn.views_b[o];
    for (int m = 0; m < kHid; ++m) accb += (double)hn.views_w[(size_t)o * vin + m] * (double)hn.feature_b[m];
    bfold[o] = (float)accb;
  }
  std::vector<uint8_t> chunks;
  for (int s = 0; s < kSteps; ++s) {
    if (s < kLayers) {
      const int kin = (s == 0) ? d.xyz : (s == skipl ? d.xyz + kHid : kHid);
      const float* W = hn.pts_w[s].data();
      const float* bs = hn.pts_b[s].data();
      auto put_bias = [&](size_t chunk_off) {      // k = 30 (column 94): lo part, k = 31 (column 95): hi part
        for (int n = 0; n < kHid; ++n) {
          uint16_t hi = to16(bs[n], fmt);
          const float hf = from16(hi, fmt);
          uint16_t lo = to16(bs[n] - hf, fmt);
          for (int k = 30; k < 32; ++k) {
            size_t b = chunk_off + (size_t)(n >> 3) * 512 + (n & 7) * 64 + (((k >> 3) ^ ((n & 7) >> 1)) << 4) + (k & 7) * 2;
            uint16_t v = (k == 30) ? lo : hi;
            memcpy(&chunks[b], &v, 2);
          }
        }
      };
      auto both = [&](const float* Wp, int k0, int kmax) { pack_chunk(chunks, Wp, kHid, kin, k0, kmax, fmt); };
      if (s == 0) {
        for (int c = 0; c < 3; ++c) both(W, 32 * c, d.xyz);
        put_bias(chunks.size() - kStageBytes);
      } else if (s == skipl) {
        for (int c = 0; c < 3; ++c) both(W, 32 * c, d.xyz);
        put_bias(chunks.size() - kStageBytes);
        for (int c = 0; c < 8; ++c) both(W + d.xyz, 32 * c, kHid);
      } else {
        chunks.resize(chunks.size() + kStageBytes, 0);     // bias chunk (first): zero but for the two bias rows
        put_bias(chunks.size() - kStageBytes);
        for (int c = 0; c < 8; ++c) both(W, 32 * c, kHid);
      }
    } else {
      // views_linears.0 (folded) rows 0..127, then the sigma head: alpha_linear weights split into bf16 hi / lo rows
      std::vector<float> wv((size_t)kViewN * kHid, 0.f);
      memcpy(wv.data(), wfold.data(), wfold.size() * sizeof(float));
      for (int i = 0; i < kHid; ++i) {
        const float hf = from16(to16(hn.alpha_w[i], fmt), fmt);
        wv[(size_t)kHalf * kHid + i] = hf;
        wv[(size_t)(kHalf + 1) * kHid + i] = hn.alpha_w[i] - hf;
      }
      for (int c = 0; c < 8; ++c) pack_chunk(chunks, wv.data(), kViewN, kHid, 32 * c, kHid, fmt);
    }
  }
  if ((int64_t)chunks.size() != step_offset(kSteps)) { *err = "internal: chunk stream size mismatch"; return -1; }
  std::vector<float> fl;
  auto put = [&](const float* src, size_t n) { size_t o = fl.size(); fl.insert(fl.end(), src, src + n); fl.resize((fl.size() + 63) / 64 * 64, 0.f); return o; };
  std::vector<float> wvd((size_t)kHalf * d.dap, 0.f);
  for (int o = 0; o < kHalf; ++o) memcpy(&wvd[(size_t)o * d.dap], &hn.views_w[(size_t)o * vin + kHid], (d.dirs + d.app) * sizeof(float));
  size_t o_wvd = put(wvd.data(), wvd.size());
  size_t o_bv = put(bfold.data(), kHalf);
  std::vector<float> bias(kLayers * kHid);
  for (int i = 0; i < kLayers; ++i) memcpy(&bias[i * kHid], hn.pts_b[i].data(), kHid * sizeof(float));
  size_t o_bias = put(bias.data(), bias.size());
  size_t o_wa = put(hn.alpha_w.data(), kHid);
  size_t o_wr = put(hn.rgb_w.data(), 3 * kHalf);

  size_t cbytes = (chunks.size() + 1023) / 1024 * 1024;
  size_t total = cbytes + fl.size() * sizeof(float);
  if (out->pool == nullptr || out->bytes < total) {
    if (out->pool) cudaFree(out->pool);
    out->pool = nullptr;
    cudaError_t e = cudaMalloc(&out->pool, total);
    if (e != cudaSuccess) { *err = cudaGetErrorString(e); return -1; }
    out->bytes = total;
  }
  uint8_t* base = (uint8_t*)out->pool;
  cudaError_t e = cudaMemcpy(base, chunks.data(), chunks.size(), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(base + cbytes, fl.data(), fl.size() * sizeof(float), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { *err = cudaGetErrorString(e); return -1; }
  const float* fb = (const float*)(base + cbytes);
  out->chunks = base;
  out->wview_d = fb + o_wvd; out->bview = fb + o_bv;
  out->bias = fb + o_bias; out->walpha = fb + o_wa; out->wrgb = fb + o_wr;
  out->balpha = hn.alpha_b[0];
  for (int i = 0; i < 3; ++i) out->brgb[i] = hn.rgb_b[i];
  out->set = true;
  return 0;
}

int tc_set_cluster(int cs) {   // 1: independent CTAs (default), 2: CTA pairs sharing every MMA (cta_group::2)
  if (cs != 1 && cs != 2) return -1;
  int dev = 0;
  if (cs == 2 && cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < kMaxDevices && g_dev[dev].done && g_dev[dev].pair_ctas < 2)
    return -1;                                                 // pairs cannot be co-scheduled on the current device
  g_cta_group = cs;
  return 0;
}

int tc_read_timeline(long long* host, int n) {
  if (!g_timeline) return 0;
  int m = std::min(n, kTlSlots);
  cudaMemcpy(host, g_timeline, m * sizeof(long long), cudaMemcpyDeviceToHost);
  return m;
}

int tc_trap_code() { return g_trap_host ? *(volatile int*)g_trap_host : 0; }

void tc_free_net(TcNet* n) {
  if (n->pool) cudaFree(n->pool);
  *n = TcNet{};
}

static int reserve(void** ws, size_t* cap, size_t need, std::string* err) {
  if (*cap >= need && *ws) return 0;
  if (*ws) cudaFree(*ws);
  *ws = nullptr; *cap = 0;
  cudaError_t e = cudaMalloc(ws, need);
  if (e != cudaSuccess) { *err = std::string("cudaMalloc(") + std::to_string(need) + "): " + cudaGetErrorString(e); return -1; }
  *cap = need;
  return cudaMemset(*ws, 0, 1024) == cudaSuccess ? 0 : -1;
}

int tc_ensure_workspace(const Dims& d, int64_t chunk_rays, void** ws, size_t* cap, std::string* err, bool with_feat) {
  size_t need = ws_layout(chunk_rays * std::max(d.sc, d.sf), chunk_rays, nullptr, nullptr, with_feat);
  return reserve(ws, cap, need, err);
}

static_assert(kSeg == kTcSegs, "partial layout");

int tc_run(const Dims& d, const TcNet& net, const TcRun& r, float** raw_out, const void** feat_out, const float** partials_out,
           int* launches, cudaStream_t s, std::string* err) {
  if (!net.set) { *err = "weights not packed"; return -1; }
  if (r.S < 1) { *err = "S < 1"; return -1; }
  TcWs w{};
  ws_layout(r.ws_rays * std::max(d.sc, d.sf), r.ws_rays, &w, (uint8_t*)r.ws, r.ws_with_feat);
  VbParams vp{};
  vp.mode = 0; vp.rays = r.rays; vp.ncols = r.ncols; vp.ray_id = r.ray_id; vp.app_table = r.app_table; vp.num_frames = r.num_frames;
  vp.mip = d.mip; vp.fd = d.fd; vp.dirs = d.dirs; vp.app = d.app; vp.dap = d.dap;
  vp.wview_d = net.wview_d; vp.bview = net.bview; vp.n = r.n; vp.out = w.vbias;
  view_bias_kernel<<<(unsigned)((r.n + kVbRows - 1) / kVbRows), 128, vb_smem_bytes(d.dap), s>>>(vp);
  *launches += 1;
  TcParams p{};
  p.vbias = w.vbias; p.vb_div = r.S;
  p.mode = 0; p.rays = r.rays; p.ncols = r.ncols; p.z = r.z; p.S = r.S; p.mip = d.mip; p.xyz = d.xyz; p.fx = d.fx; p.var_scale = r.var_scale;
  p.P = r.n * r.S;
  p.stop_step = r.stop_layer < 0 ? kLayers - 1 : r.stop_layer; p.want_feat = r.want_feat ? 1 : 0;
  p.raw = w.raw; p.feat_tiles = w.feat; p.debug_step = -1; p.debug_out = nullptr; p.error_flag = w.error_flag;
  const bool fuse = r.fuse_composite && r.want_feat && r.S >= kTcMinFuseS;
  if (r.want_feat && !fuse && !r.ws_with_feat) { *err = "internal: workspace was sized without the per-sample feature tiles"; return -1; }
  p.partials = fuse ? w.partials : nullptr;
  if (launch_tc(net, p, launches, s, err) != 0) return -1;
  *raw_out = w.raw; *feat_out = w.feat; *partials_out = p.partials;
  return 0;
}

int tc_run_rows(const Dims& d, const TcNet& net, const float* X, int64_t P, int stop_layer, float* raw, float* feat,
                void** ws, size_t* cap, int* launches, cudaStream_t s, std::string* err, int debug_steps, float* debug_out, uint32_t* masks) {
  if (!net.set) { *err = "weights not packed"; return -1; }
  size_t need = ws_layout(P, P, nullptr, nullptr);
  if (reserve(ws, cap, need, err) != 0) return -1;
  TcWs w{};
  ws_layout(P, P, &w, (uint8_t*)*ws);
  VbParams vp{};
  vp.mode = 1; vp.X = X; vp.kx = d.kx; vp.xyzp = d.xyzp;
  vp.mip = d.mip; vp.fd = d.fd; vp.dirs = d.dirs; vp.app = d.app; vp.dap = d.dap;
  vp.wview_d = net.wview_d; vp.bview = net.bview; vp.n = P; vp.out = w.vbias;
  view_bias_kernel<<<(unsigned)((P + kVbRows - 1) / kVbRows), 128, vb_smem_bytes(d.dap), s>>>(vp);
  *launches += 1;
  TcParams p{};
  p.vbias = w.vbias; p.vb_div = 1;
  p.mode = 1; p.X = X; p.kx = d.kx; p.xyz = d.xyz; p.S = 1;
  p.P = P;
  p.stop_step = stop_layer < 0 ? kLayers - 1 : stop_layer; p.want_feat = feat ? 1 : 0;
  p.raw = w.raw; p.feat_tiles = w.feat; p.error_flag = w.error_flag;
  p.masks = masks;
  p.debug_step = debug_steps > 0 ? debug_steps - 1 : -1; p.debug_out = debug_steps > 0 ? debug_out : nullptr;
  if (launch_tc(net, p, launches, s, err) != 0) return -1;
  cudaError_t e = cudaMemcpyAsync(raw, w.raw, (size_t)P * 4 * sizeof(float), cudaMemcpyDeviceToDevice, s);
  if (e != cudaSuccess) { *err = cudaGetErrorString(e); return -1; }
  if (feat) {
    detile_kernel<<<(unsigned)((P * 32 + 255) / 256), 256, 0, s>>>(w.feat, P, feat, net.fmt);
    *launches += 1;
  }
  return 0;
}

}  // namespace nmr
