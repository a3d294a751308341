// scatt_attn_block (sm_100a): the row-local tail of a self / merge layer as ONE tcgen05 kernel
//
//   h = LayerNorm1(x + ctx Wo^T + bo)            model/keypoint_module.py:62-66,98-102 (out_proj: model/attention.py:74)
//   y = LayerNorm2(h + GELU(h W1^T + b1) W2^T + b2)   model/keypoint_module.py:68-72,104-107, model/layers.py:103-108
//
// for a 128-row tile per CTA (D = 256 model columns, F = 128 n hidden columns).  Three GEMM launches used to move
// ctx, x, h (twice), the F-wide hidden activation (twice) and y through global memory; here only ctx and x come in
// and y goes out - h and the hidden activation never leave the SM.
//
// Warp roles (320 threads): warp 0 TMA producer, warp 1 tcgen05.mma issuer, warps 2-9 epilogue (two per TMEM lane
// quadrant, each owning half of the columns of whatever is being drained).
//
// Tensor memory (512 columns):
//   [0, 256)    accO = ctx Wo^T + x.  The residual x is added BY THE TENSOR CORE: its hi and lo tiles stream through
//               the ring like weights and are multiplied with identity tiles (N = 32 MMAs; exact products, fp32
//               adds) - no register transposition of a row-major residual into the thread-per-row accumulator
//               layout (that pre-initialisation cost 8-10 k cycles per tile in the first version of this kernel).
//               LayerNorm1 (+ bo) reads accO, rewrites it with h + b2 and it becomes the fc2 accumulator, so the
//               second residual add costs nothing either.
//   [256, 384), [384, 512)   two fc1 chunk accumulators (128 hidden columns each).  The epilogue warps apply
//               bias + GELU and write the chunk back IN PLACE as packed 16-bit hi | lo - it is then the TMEM A
//               operand of the fc2 MMAs of that chunk (like P in the attention kernel): the hidden activation
//               never touches shared memory.  While chunk j is being activated the MMA warp computes chunk j + 1.
//               The same columns carry the LayerNorm partial statistics between the two warps of a row.
// Shared memory (227 KB):
//   hA   128 KB: ctx tile (K-major, 128-byte swizzle, 4 k-blocks x hi | lo) = A operand of out_proj; LayerNorm1 then
//               overwrites it with h in the same layout = A operand of fc1; at the end it holds the output boxes the
//               TMA engine stores.
//   ring 3 x 32 KB: one TMA operation each (profiles/r02_ubench_tma_mma_v2.txt: a producer warp issues one operation
//               per ~680 cycles whatever its size, so boxes are as large as the ring allows), streamed in MMA order:
//               Wo [256 x 64] hi / lo planes, x [128 x 64] hi | lo, W1 chunk [128 x 64] hi | lo, W2 [256 x 64] hi / lo.
//   2 KB        two identity tiles [32 x 16] (32-byte swizzle).
// MMA shapes follow the measured rates (same file): N = 256 wherever the whole 256-column output is produced
// (out_proj, fc2; fc2 reads A from TMEM: 139 cycles against the 128-cycle floor), N = 128 for the fc1 chunks
// (shared-memory A: 108 cycles against 64 - the shared-memory read port, not the tensor pipe, is the limit).
// The kernel is persistent: CTA b walks tiles b, b + grid, ...; the ring runs ahead across tile boundaries.
//
// CL = 2 (few row tiles: the small-batch regime): a 2-CTA cluster shares one row tile.  Both CTAs compute out_proj +
// LayerNorm1 (identical results - nothing to exchange), then CTA r takes the hidden chunks j = r, r + 2, ... : half of
// the W1 / W2 stream and half of the fc1 / fc2 MMAs each.  fc2 is thereby split along K: CTA 1 accumulates its partial
// sum on top of zeros, CTA 0 on top of h + b2.  LayerNorm2 and the output are split by COLUMNS: CTA r owns columns
// [128 r, 128 r + 128); each CTA ships the other half of its 128 x 256 fp32 partial sum to the peer through distributed
// shared memory (staged in its own idle hA, one cp.async.bulk shared::cta -> shared::cluster of 8 KB per warp, both
// directions at once), adds what it receives, exchanges the row statistics of its 128 columns (st.shared::cluster into
// the peer's staging zone) and normalises and stores its half.  Hand-shake: mbarriers with cluster-scope release /
// acquire - "landing zone free" (peer_free), the copies' complete_tx (partial_full), "staging read" (xfer_done),
// "statistics written" (stats_full).
#include <cuda.h>

#include <cstdlib>

#include "common.cuh"
#include "tc_epi.cuh"
#include "tc_host.cuh"
#include "tc_ptx.cuh"

#ifndef SCATT_BLOCK_TRACE
#define SCATT_BLOCK_TRACE 0
#endif

namespace scatt {

namespace {

using namespace tc;

constexpr int BM = 128;
constexpr int DM = 256;          // model width (columns of ctx, x, h, y)
constexpr int kMaxF = 4096;      // hidden width limit (a multiple of 128; nothing is sized by it)
constexpr int kEpiWarps = 8;
constexpr int kThreads = 64 + 32 * kEpiWarps;
constexpr int kStages = 3;
constexpr uint32_t kTileBytes = 16384;   // [128 x 64] 16-bit tile
constexpr uint32_t kStageBytes = 32768;  // [256 x 64] one plane, or [128 x 64] hi | lo
constexpr uint32_t kKbBytes = 32768;     // one A k-block: hi tile + lo tile
constexpr uint32_t kHABytes = 4 * kKbBytes;
constexpr uint32_t kTmemCols = 512, kAccO = 0, kAccB = 256;
__device__ constexpr int kOrder[4] = {0, 2, 1, 3};  // k-blocks of h in the order LayerNorm1 completes them (both column halves advance together)
constexpr int kTraceSlots = 160;  // 32-bit stamps: 0..63 phases, 64 + i / 112 + i ring item i issued / taken (i < 48)

struct BlkProblem {
  const float *bo, *g1, *be1, *b1, *b2, *g2, *be2;
  float* y;
  uint16_t* y_planes;
};

struct alignas(64) BlkParams {
  CUtensorMap map_ctx[SCATT_MAX_GROUP];  // [2][M][256]   box 64 x 128 x (1 | 2)
  CUtensorMap map_x[SCATT_MAX_GROUP];    // [2][M][256]   box 64 x 128 x 2
  CUtensorMap map_wo[SCATT_MAX_GROUP];   // [2][256][256] box 64 x 256 x 1
  CUtensorMap map_w1[SCATT_MAX_GROUP];   // [2][F][256]   box 64 x 128 x (1 | 2)
  CUtensorMap map_w2[SCATT_MAX_GROUP];   // [2][256][F]   box 64 x 256 x 1
  CUtensorMap map_y[SCATT_MAX_GROUP];
  CUtensorMap map_p[SCATT_MAX_GROUP];
  BlkProblem prob[SCATT_MAX_GROUP];
  int64_t M;
  int32_t F, nchunk, terms, fmt, groups, tiles_m;
  float eps;
  float qscale;  // MODE 1: q = (h Wq^T + bq) * qscale
};

// shared memory map relative to the (1024-aligned) start of dynamic shared memory
constexpr uint32_t kRingOff = kHABytes;
constexpr uint32_t kIdOff = kRingOff + kStages * kStageBytes;  // 2 identity tiles of 32 rows x 32 B
constexpr uint32_t kBarOff = kIdOff + 2048;
constexpr uint32_t kTraceOff = kBarOff + 256;
constexpr uint32_t kSmemBytes = kTraceOff + (SCATT_BLOCK_TRACE ? kTraceSlots * 4 : 0);
static_assert(kSmemBytes <= 227 * 1024, "attn_block: shared memory map exceeds 227 KB");

// Phase trace (dev builds with -DSCATT_BLOCK_TRACE=1, tools/trace_block.py): CTA 0 stamps the low 32 bits of
// clock64() into a shared array (a global store per stamp would perturb the pipeline) and copies it out at the end.
__device__ long long* g_trace_blk = nullptr;
#define trace(slot)                                                         \
  do {                                                                      \
    if (SCATT_BLOCK_TRACE && tracing) trp[slot] = uint32_t(clock64());      \
  } while (0)

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory"); }

// K-major tile with 32-byte rows (16 K-elements), 32-byte swizzle, 8-row groups 256 B apart
__device__ __forceinline__ uint64_t umma_desc_sw32(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= uint64_t((smem_addr & 0x3FFFFu) >> 4);
  d |= uint64_t(1) << 16;
  d |= uint64_t(256 >> 4) << 32;
  d |= uint64_t(1) << 46;
  d |= uint64_t(6) << 61;
  return d;
}

// LayerNorm statistics of 128 columns held by one thread-row, combined with the partner warp's 128 columns
// (Chan et al.) through two spare TMEM columns; returns (mean, rstd) over the 256 columns.
__device__ __forceinline__ float2 combine_stats(uint32_t tmem_x, int hf, float shift, float s1, float s2, float eps) {
  constexpr float kHalfN = float(DM / 2);
  const float dm = s1 / kHalfN;
  const float my_mean = shift + dm, my_m2 = fmaxf(s2 - s1 * dm, 0.f);
  tc_st2(tmem_x + uint32_t(hf * 2), my_mean, my_m2);
  tc_fence_before();
  epi_bar_sync();
  tc_fence_after();
  float om, oq;
  tc_ld2(tmem_x + uint32_t((hf ^ 1) * 2), om, oq);
  const float mean = 0.5f * (my_mean + om);
  const float da = my_mean - mean, db = om - mean;
  const float m2 = my_m2 + oq + kHalfN * (da * da + db * db);
  return make_float2(mean, rsqrtf(m2 / float(DM) + eps));
}

// v[32] += p[0..32) for a warp-uniform global pointer (per-column parameters: L1-resident broadcast loads)
__device__ __forceinline__ void add_cols_g(float* v, const float* __restrict__ p) {
#pragma unroll
  for (int j = 0; j < 32; j += 4) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p + j));
    v[j] += t.x, v[j + 1] += t.y, v[j + 2] += t.z, v[j + 3] += t.w;
  }
}

__device__ __forceinline__ uint32_t map_to_peer(uint32_t local_addr, uint32_t peer_rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_addr), "r"(peer_rank));
  return remote;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t remote_bar) {  // releases this thread's earlier (remote) writes
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote_bar) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {  // acquire at cluster scope
  #pragma unroll 1  // the compiler otherwise unrolls every spin loop four-fold: 40 % of the fused tail kernel was wait code
  for (uint32_t it = 0; it < kWaitLimit; ++it) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return;
  }
  __trap();
}

// MODE 0: the whole layer tail (above).  MODE 1 (scatt_attn_out_q, the causal layer in front of a merge layer, reference
// model/keypoint_module.py:62-66 followed by the q_proj of model/attention.py:99-101): out_proj + residual + LayerNorm as in
// mode 0, h is ALSO stored to global memory (TMA stores straight out of the operand tiles in hA), and the "fc1 chunks" are
// the merge layer's q projection - bias, q scaling, hi / lo split and TMA stores instead of GELU + fc2.  No second LayerNorm,
// and in a 2-CTA cluster nothing to exchange: CTA r computes q columns [128 r, 128 r + 128).  Two ring stages; the third
// stage's 32 KB are the epilogue warps' output boxes.
template <int FMT, int CL, int MODE = 0>
__global__ void __launch_bounds__(kThreads, 1) attn_block_kernel(const __grid_constant__ BlkParams P) {
  constexpr uint32_t kNst = MODE == 1 ? 2u : uint32_t(kStages);  // ring stages in use
  extern __shared__ __align__(1024) uint8_t sm[];
  const uint32_t base = smem_u32(sm);
  const uint32_t hA = base, ring = base + kRingOff, bar0 = base + kBarOff;
  auto full_bar = [&](uint32_t s) { return bar0 + 8u * s; };
  auto empty_bar = [&](uint32_t s) { return bar0 + 8u * (kStages + s); };
  const uint32_t ctx_bar = bar0 + 8u * 2 * kStages;  // [4], one per ctx k-block
  const uint32_t oproj_full = ctx_bar + 32u, h_ready = oproj_full + 8u;  // h_ready [2]: k-blocks {0, 2} / {1, 3} of h are in hA
  const uint32_t fc1_full = h_ready + 16u;  // [2]
  const uint32_t g_ready = fc1_full + 16u;  // [2]
  const uint32_t out_full = g_ready + 16u, tile_done = out_full + 8u;
  const uint32_t peer_free = tile_done + 8u, partial_full = peer_free + 8u, xfer_done = partial_full + 8u, stats_full = xfer_done + 8u;
  const uint32_t tmem_ptr_addr = stats_full + 8u;
  volatile uint32_t* trp = reinterpret_cast<volatile uint32_t*>(sm + kTraceOff);
  const bool tracing = SCATT_BLOCK_TRACE && g_trace_blk != nullptr && blockIdx.x == 0;
  (void)trp, (void)tracing;

  const int warp = scatt_warp_idx(), lane = threadIdx.x & 31;
  const int total_tiles = P.tiles_m * P.groups;
  // CL = 2: the CTAs of a cluster walk the same tiles; CTA `rank` owns the hidden chunks rank, rank + 2, ...
  const uint32_t rank = CL > 1 ? cluster_ctarank() : 0u;
  const int first_tile = CL > 1 ? int(blockIdx.x) / CL : int(blockIdx.x), tile_step = CL > 1 ? int(gridDim.x) / CL : int(gridDim.x);
  const int nchunk = (P.nchunk - int(rank) + CL - 1) / CL;  // local chunks; local chunk i is global chunk rank + i * CL
  auto chunk_of = [&](int i) { return int(rank) + i * CL; };
  const bool a_lo = P.terms >= 2, b_lo = P.terms >= 3;

  if ((base & 1023u) != 0u) __trap();  // the swizzled tiles need the 1024-byte alignment the declaration asks for
  if (tracing) {
    for (int i = threadIdx.x; i < kTraceSlots; i += kThreads) trp[i] = 0;
  }
  if (threadIdx.x == 0) {
    for (uint32_t s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (uint32_t k = 0; k < 4; ++k) mbar_init(ctx_bar + 8u * k, 1);
    mbar_init(oproj_full, 1);
    mbar_init(h_ready, 32 * kEpiWarps);
    mbar_init(h_ready + 8u, 32 * kEpiWarps);
    for (uint32_t b = 0; b < 2; ++b) {
      mbar_init(fc1_full + 8u * b, 1);
      mbar_init(g_ready + 8u * b, 32 * kEpiWarps);
    }
    mbar_init(out_full, 1);
    mbar_init(tile_done, 1);
    mbar_init(peer_free, 1);
    mbar_init(partial_full, 1);
    mbar_init(xfer_done, 1);
    mbar_init(stats_full, 4 * 32);  // one arrival per row: the peer's warps of column half 0
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int g = 0; g < P.groups; ++g) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&P.map_ctx[g]) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&P.map_x[g]) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&P.map_wo[g]) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&P.map_w1[g]) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&P.map_w2[g]) : "memory");
    }
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_addr), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp >= 2) {
    // identity tiles (K-major, 32-byte rows = 16 K-elements, 32-byte swizzle): tile 0 = [I16; 0], tile 1 = [0; I16].
    // D[:, 32 c .. 32 c + 32) += X[:, 16 k-elements] * tile^T puts 16 columns of X into the lower / upper half.
    const int tid = threadIdx.x - 64;
    uint32_t* id = reinterpret_cast<uint32_t*>(sm + kIdOff);
    for (int i = tid; i < 512; i += 32 * kEpiWarps) id[i] = 0u;
    epi_bar_sync();
    if (tid < 32) {
      const int tile = tid >> 4, k = tid & 15, n = tile * 16 + k;  // element (row n, column k) of tile `tile`
      const uint32_t off = uint32_t(tile * 1024 + (n >> 3) * 256 + (n & 7) * 32 + (((k >> 3) ^ ((n >> 2) & 1)) << 4) + (k & 7) * 2);
      *reinterpret_cast<uint16_t*>(sm + kIdOff + off) = FMT == SCATT_PLANE_F16 ? uint16_t(0x3C00) : uint16_t(0x3F80);
    }
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if constexpr (CL > 1) cluster_sync_all();  // the peer's mbarriers exist before anybody arrives on them remotely
  pdl_launch_dependents();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(sm + (tmem_ptr_addr - base));
  if (threadIdx.x == 0) trace(0);

  if (warp == 0) {  // ================================================= TMA producer
    uint32_t it = 0;  // ring items issued so far (the ring runs across tile boundaries)
    auto put = [&](const CUtensorMap* map, int c0, int c1, int c2, uint32_t bytes) {  // one TMA operation into the next stage
      const uint32_t s = it % kNst;
      mbar_wait(empty_bar(s), ((it / kNst) & 1u) ^ 1u);
      if (elect_one()) {
        mbar_expect_tx(full_bar(s), bytes);
        tma_load_3d(ring + s * kStageBytes, map, full_bar(s), c0, c1, c2);
        if (it < 48) trace(64 + int(it));
      }
      __syncwarp();
      ++it;
    };
    int ti = 0;
    for (int t = first_tile; t < total_tiles; t += tile_step, ++ti) {
      const int g = t / P.tiles_m, m0 = (t % P.tiles_m) * BM;
      auto put_wo = [&](int kb) {
        put(&P.map_wo[g], kb * 64, 0, 0, kStageBytes);
        if (b_lo) put(&P.map_wo[g], kb * 64, 0, 1, kStageBytes);
      };
      // the first Wo k-block is a static weight: it may be fetched before the activations exist
      put_wo(0);
      if (lane == 0 && ti == 0) trace(15);
      if (ti == 0) pdl_wait();                                     // ctx / x come from the preceding kernels
      else mbar_wait(tile_done, uint32_t(ti - 1) & 1u);            // the previous tile's output boxes have left hA
      if (elect_one()) {
        for (int kb = 0; kb < 4; ++kb) {
          mbar_expect_tx(ctx_bar + 8u * kb, a_lo ? kKbBytes : kTileBytes);
          tma_load_3d(hA + kb * kKbBytes, &P.map_ctx[g], ctx_bar + 8u * kb, kb * 64, m0, 0);  // hi | lo in one box when a_lo
        }
      }
      __syncwarp();
      if (lane == 0 && ti == 0) trace(16);
      for (int kb = 0; kb < 4; ++kb) put(&P.map_x[g], kb * 64, m0, 0, kStageBytes);
      for (int kb = 1; kb < 4; ++kb) put_wo(kb);
      auto fc1_items = [&](int i) {
        const int j = chunk_of(i);
        for (int q = 0; q < 4; ++q) put(&P.map_w1[g], kOrder[q] * 64, j * 128, 0, b_lo ? kStageBytes : kTileBytes);  // the order LayerNorm1 finishes them in
      };
      auto fc2_items = [&](int i) {
        const int j = chunk_of(i);
        for (int kb = 0; kb < 2; ++kb) {
          put(&P.map_w2[g], (2 * j + kb) * 64, 0, 0, kStageBytes);
          if (b_lo) put(&P.map_w2[g], (2 * j + kb) * 64, 0, 1, kStageBytes);
        }
      };
      if (nchunk > 0) {
        fc1_items(0);
        for (int i = 1; i < nchunk; ++i) {
          fc1_items(i);
          if (MODE == 0) fc2_items(i - 1);
        }
        if (MODE == 0) fc2_items(nchunk - 1);
      }
      if (lane == 0 && ti == 0) trace(17);
    }
  } else if (warp == 1) {  // ========================================== MMA issuer
    const uint32_t idesc_base = (1u << 4) | (uint32_t(FMT) << 7) | (uint32_t(FMT) << 10) | (uint32_t(BM >> 4) << 24);
    const uint32_t idesc256 = idesc_base | (uint32_t(256 >> 3) << 17), idesc128 = idesc_base | (uint32_t(128 >> 3) << 17);
    const uint32_t idesc32 = idesc_base | (uint32_t(32 >> 3) << 17);
    const uint64_t id0 = umma_desc_sw32(base + kIdOff), id1 = umma_desc_sw32(base + kIdOff + 1024u);
    uint32_t it = 0;
    uint32_t gphase = 0;  // bit b: parity of the next completion of g_ready[b]
    auto take = [&]() -> uint32_t {  // waits for the next ring item; returns its stage
      const uint32_t s = it % kNst;
      mbar_wait(full_bar(s), (it / kNst) & 1u);
      if (lane == 0 && it < 48) trace(112 + int(it));
      ++it;
      tc_fence_after();
      return s;
    };
    int ti = 0;
    for (int t = first_tile; t < total_tiles; t += tile_step, ++ti) {
      const uint32_t tpar = uint32_t(ti) & 1u;
      const uint32_t dO = tmem + kAccO;
      // ---- out_proj: accO[128 x 256] = ctx Wo^T (N = 256 MMAs), + x through the identity tiles (N = 32 MMAs).
      // accO may be overwritten: ctx k-block 0 lands only after tile_done of the previous tile, i.e. after its
      // LayerNorm2 has read the accumulator.
      auto issue_wo = [&](int kb) {
        mbar_wait(ctx_bar + 8u * kb, tpar);
        tc_fence_after();
        if (lane == 0 && ti == 0) trace(10 + kb);
        const uint64_t ah = umma_desc_sw128(hA + kb * kKbBytes), al = umma_desc_sw128(hA + kb * kKbBytes + kTileBytes);
        {
          const uint32_t s = take();
          const uint64_t bh = umma_desc_sw128(ring + s * kStageBytes);
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              const uint64_t adv = uint64_t(kk * 2);
              tc_mma_f16(dO, ah + adv, bh + adv, idesc256, (kb | kk) ? 1u : 0u);  // the tile's very first MMA starts accO
              if (a_lo) tc_mma_f16(dO, al + adv, bh + adv, idesc256, 1);
            }
            tc_commit(empty_bar(s));
          }
          __syncwarp();
        }
        if (b_lo) {
          const uint32_t s = take();
          const uint64_t bl = umma_desc_sw128(ring + s * kStageBytes);
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) tc_mma_f16(dO, ah + uint64_t(kk * 2), bl + uint64_t(kk * 2), idesc256, 1);
            tc_commit(empty_bar(s));
          }
          __syncwarp();
        }
      };
      issue_wo(0);
      for (int kb = 0; kb < 4; ++kb) {  // x[:, 64 kb .. 64 kb + 64) hi | lo: 16 columns per MMA against an identity tile
        const uint32_t s = take();
        const uint64_t xh = umma_desc_sw128(ring + s * kStageBytes), xl = umma_desc_sw128(ring + s * kStageBytes + kTileBytes);
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint32_t d = dO + uint32_t(kb * 64 + (kk >> 1) * 32);
            const uint64_t idd = (kk & 1) ? id1 : id0;
            tc_mma_f16(d, xh + uint64_t(kk * 2), idd, idesc32, 1);
            tc_mma_f16(d, xl + uint64_t(kk * 2), idd, idesc32, 1);
          }
          tc_commit(empty_bar(s));
        }
        __syncwarp();
      }
      for (int kb = 1; kb < 4; ++kb) issue_wo(kb);
      if (elect_one()) tc_commit(oproj_full);
      __syncwarp();
      if (lane == 0 && ti == 0) trace(2);
      // h arrives in hA in two halves: LayerNorm1 finishes k-blocks {0, 2} (columns [0, 64) and [128, 192)) half way through
      // its second pass - the first chunk's MMAs over those start under the rest of the pass
      auto issue_fc1 = [&](int j, bool first) {  // accB[j & 1][128 x 128] = h W1[chunk j]^T
        const uint32_t d = tmem + kAccB + uint32_t((j & 1) * 128);
        if (lane == 0 && ti == 0 && j < 8) trace(20 + j);
        for (int q = 0; q < 4; ++q) {
          const int kb = kOrder[q];
          if (first && (q == 0 || q == 2)) {
            mbar_wait(h_ready + 8u * uint32_t(q >> 1), tpar);
            tc_fence_after();
            if (lane == 0 && ti == 0 && q == 2) trace(3);
          }
          const uint32_t s = take();
          const uint64_t ah = umma_desc_sw128(hA + kb * kKbBytes), al = umma_desc_sw128(hA + kb * kKbBytes + kTileBytes);
          const uint64_t bh = umma_desc_sw128(ring + s * kStageBytes), bl = umma_desc_sw128(ring + s * kStageBytes + kTileBytes);
          if (elect_one()) {
            uint32_t acc = q > 0 ? 1u : 0u;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              const uint64_t adv = uint64_t(kk * 2);
              if (b_lo) {
                tc_mma_f16(d, ah + adv, bl + adv, idesc128, acc);
                acc = 1;
              }
              if (a_lo) {
                tc_mma_f16(d, al + adv, bh + adv, idesc128, acc);
                acc = 1;
              }
              tc_mma_f16(d, ah + adv, bh + adv, idesc128, acc);
              acc = 1;
            }
            tc_commit(empty_bar(s));
            if (q == 3) tc_commit(fc1_full + 8u * uint32_t(j & 1));
          }
          __syncwarp();
        }
      };
      auto issue_fc2 = [&](int j, bool last) {  // accO += g[chunk j] W2[:, chunk j]^T, A = g from TMEM, N = 256
        const uint32_t b = uint32_t(j & 1);
        mbar_wait(g_ready + 8u * b, (gphase >> b) & 1u);
        gphase ^= 1u << b;
        tc_fence_after();
        if (lane == 0 && ti == 0 && j < 8) trace(30 + j);
        const uint32_t gbase = tmem + kAccB + b * 128u;
        for (int kb = 0; kb < 2; ++kb) {
          {
            const uint32_t s = take();
            const uint64_t bh = umma_desc_sw128(ring + s * kStageBytes);
            if (elect_one()) {
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) {
                const uint32_t g_hi = gbase + uint32_t(kb * 64 + kk * 16);  // k-step 4 kb + kk: hi pairs, lo pairs 8 columns on
                tc_mma_f16_ts(dO, g_hi, bh + uint64_t(kk * 2), idesc256, 1);
                if (a_lo) tc_mma_f16_ts(dO, g_hi + 8u, bh + uint64_t(kk * 2), idesc256, 1);
              }
              tc_commit(empty_bar(s));
              if (last && kb == 1 && !b_lo) tc_commit(out_full);
            }
            __syncwarp();
          }
          if (b_lo) {
            const uint32_t s = take();
            const uint64_t bl = umma_desc_sw128(ring + s * kStageBytes);
            if (elect_one()) {
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) tc_mma_f16_ts(dO, gbase + uint32_t(kb * 64 + kk * 16), bl + uint64_t(kk * 2), idesc256, 1);
              tc_commit(empty_bar(s));
              if (last && kb == 1) tc_commit(out_full);
            }
            __syncwarp();
          }
        }
      };
      // tcgen05.mma executes in issue order: fc1(j + 1) can be issued before fc2(j - 1) has read its operand chunk
      // because it writes the OTHER accumulator, and fc1(j + 2) after fc2(j) reuses that chunk's columns safely
      if (nchunk > 0) {
        issue_fc1(0, true);
        for (int i = 1; i < nchunk; ++i) {
          issue_fc1(i, false);
          if (MODE == 0) issue_fc2(i - 1, false);
        }
        if (MODE == 0) issue_fc2(nchunk - 1, true);
      } else {  // a cluster CTA without a hidden chunk (F = 128): its partial sum is the zero it started from
        mbar_wait(h_ready, tpar);
        mbar_wait(h_ready + 8u, tpar);
        tc_fence_after();
        if (elect_one()) tc_commit(out_full);
        __syncwarp();
      }
      if (lane == 0 && ti == 0) trace(4);
    }
  } else {  // ========================================================= epilogue warps 2..9
    const int quad = warp & 3, hf = (warp - 2) >> 2;
    const int row = quad * 32 + lane;
    const uint32_t lane_addr = uint32_t(quad * 32) << 16;
    const uint32_t tmem_x = tmem + kAccB + lane_addr;  // statistics exchange columns (accB is idle around both LayerNorms)
    const int tid = threadIdx.x - 64;
    uint32_t fphase = 0;  // bit b: parity of the next completion of fc1_full[b]
    int ti = 0, cur_g = -1;
    for (int t = first_tile; t < total_tiles; t += tile_step, ++ti) {
      const uint32_t tpar = uint32_t(ti) & 1u;
      const int g = t / P.tiles_m;
      const int64_t m0 = int64_t(t % P.tiles_m) * BM;
      const BlkProblem& Q = P.prob[g];
      const int64_t row0 = m0 + quad * 32;
      if (g != cur_g) {
        // Per-column parameters are read through L1 as warp-uniform loads (no shared memory left for a copy): pull
        // their lines in while the out_proj MMAs run, or every 32-column chunk pays a cold L2 round trip
        // (LayerNorm1 took 9.5 k instead of 5 k cycles on a CTA's first tile).  One 128-byte line per thread.
        const float* vec[6] = {Q.bo, Q.g1, Q.be1, Q.b2, Q.g2, Q.be2};
        if (tid < 48) asm volatile("prefetch.global.L1 [%0];" ::"l"(vec[tid >> 3] + (tid & 7) * 32));
        else if (tid - 48 < P.F / 32) asm volatile("prefetch.global.L1 [%0];" ::"l"(Q.b1 + (tid - 48) * 32));
        cur_g = g;
      }

      // ---- LayerNorm1 over accO + bo -> h: planes into hA (fc1's A operand), h + b2 back into accO (fc2's accumulator)
      mbar_wait(oproj_full, tpar);
      tc_fence_after();
      if (tid == 0 && ti == 0) trace(5);
      {
        float v[32];
        float shift = 0.f;
        float2 s1 = make_float2(0.f, 0.f), s2 = s1;
        // pass 1: + bo (written back, so that the second pass does not fetch it again), shifted sums; packed fp32 arithmetic
#pragma unroll 1
        for (int i = 0; i < 4; ++i) {
          const int cl = hf * 128 + i * 32;
          tc_ld32(tmem + kAccO + lane_addr + cl, v);
          add_cols_g(v, Q.bo + cl);
          if (i == 0) shift = v[0];
          const float2 ns = make_float2(-shift, -shift);
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            const float2 d = __fadd2_rn(make_float2(v[j], v[j + 1]), ns);
            s1 = __fadd2_rn(s1, d);
            s2 = __ffma2_rn(d, d, s2);
          }
          tc_st32(tmem + kAccO + lane_addr + cl, v);
        }
        if (tid == 0 && ti == 0) trace(18);
        const float2 mr = combine_stats(tmem_x, hf, shift, s1.x + s1.y, s2.x + s2.y, P.eps);
        if (tid == 0 && ti == 0) trace(19);
        const float2 nm = make_float2(-mr.x, -mr.x), rs = make_float2(mr.y, mr.y);
        uint8_t* hrow = sm + (row >> 3) * 1024 + (row & 7) * 128;
#pragma unroll 1
        for (int i = 0; i < 4; ++i) {
          const int cl = hf * 128 + i * 32;
          tc_ld32(tmem + kAccO + lane_addr + cl, v);
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 gg = __ldg(reinterpret_cast<const float4*>(Q.g1 + cl + j));
            const float4 bb = __ldg(reinterpret_cast<const float4*>(Q.be1 + cl + j));
            const float2 y0 = __ffma2_rn(__fmul2_rn(__fadd2_rn(make_float2(v[j], v[j + 1]), nm), rs), make_float2(gg.x, gg.y), make_float2(bb.x, bb.y));
            const float2 y1 = __ffma2_rn(__fmul2_rn(__fadd2_rn(make_float2(v[j + 2], v[j + 3]), nm), rs), make_float2(gg.z, gg.w), make_float2(bb.z, bb.w));
            v[j] = y0.x, v[j + 1] = y0.y, v[j + 2] = y1.x, v[j + 3] = y1.y;
          }
          // K-major 128-byte-swizzled A tiles: k-block cl / 64, 16-byte chunk (cl % 64) / 8 + q of row `row`
          uint8_t* kbp = hrow + (cl >> 6) * kKbBytes;
          const int ch0 = (cl & 63) >> 3;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 hi, lo;
            split8<FMT>(make_float4(v[8 * q], v[8 * q + 1], v[8 * q + 2], v[8 * q + 3]),
                        make_float4(v[8 * q + 4], v[8 * q + 5], v[8 * q + 6], v[8 * q + 7]), hi, lo);
            const uint32_t off = uint32_t(((ch0 + q) ^ (row & 7)) << 4);
            *reinterpret_cast<uint4*>(kbp + off) = hi;
            *reinterpret_cast<uint4*>(kbp + kTileBytes + off) = lo;
          }
          if constexpr (MODE == 0) {
            if (rank == 0) {  // fc2 accumulates on top of h + b2 (CTA 1 of a cluster: on top of zero, its sum is added later)
              add_cols_g(v, Q.b2 + cl);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = 0.f;
            }
            tc_st32(tmem + kAccO + lane_addr + cl, v);
          }
          if (i & 1) {  // k-blocks {0, 2} after the first two chunks, {1, 3} after the last two
            fence_proxy_async();  // generic-proxy writes of h -> visible to the tensor core's operand reads
            tc_fence_before();
            mbar_arrive(h_ready + 8u * uint32_t(i >> 1));
          }
        }
      }
      if (tid == 0 && ti == 0) trace(6);

      if constexpr (MODE == 1) {
        // ---- h -> global memory, straight out of the operand tiles (hi tile | lo tile of a k-block = one 64 x 128 x 2 box);
        // CTA r of a cluster stores k-blocks 2 r and 2 r + 1 (both CTAs hold the same h)
        if (tid == 0) {
          mbar_wait(h_ready, tpar);
          mbar_wait(h_ready + 8u, tpar);  // every warp's writes (each followed by fence.proxy.async) have been made
          const int kb0 = CL > 1 ? 2 * int(rank) : 0, nkb = CL > 1 ? 2 : 4;
          for (int kb = kb0; kb < kb0 + nkb; ++kb) tma_store_3d(&P.map_y[g], hA + uint32_t(kb) * kKbBytes, kb * 64, int(m0), 0);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        // ---- q chunks: (acc + bq) * qscale -> hi / lo planes, 32 columns per pair of TMA stores out of this warp's 4 KB box
        // (the third ring stage is not used as a ring in this mode)
        uint8_t* qbox = sm + kRingOff + 2 * kStageBytes + uint32_t(warp - 2) * 4096u;
        const uint32_t qbox_addr = ring + 2 * kStageBytes + uint32_t(warp - 2) * 4096u;
        const float2 qs2 = make_float2(P.qscale, P.qscale);
#pragma unroll 1
        for (int j = 0; j < nchunk; ++j) {
          const uint32_t b = uint32_t(j & 1);
          mbar_wait(fc1_full + 8u * b, (fphase >> b) & 1u);
          fphase ^= 1u << b;
          tc_fence_after();
          if (tid == 0 && ti == 0 && j < 8) trace(40 + j);
          const uint32_t ca = tmem + kAccB + b * 128u + lane_addr + uint32_t(hf * 64);
          const int col0 = chunk_of(j) * 128 + hf * 64;
#pragma unroll 1
          for (int i = 0; i < 2; ++i) {
            float v[32];
            tc_ld32(ca + uint32_t(32 * i), v);
            add_cols_g(v, Q.b1 + col0 + 32 * i);
#pragma unroll
            for (int e = 0; e < 32; e += 2) {
              const float2 t = __fmul2_rn(make_float2(v[e], v[e + 1]), qs2);
              v[e] = t.x, v[e + 1] = t.y;
            }
            if (j + i > 0) {  // the box's previous contents must have been read out
              if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
              __syncwarp();
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              uint4 hi, lo;
              split8<FMT>(make_float4(v[8 * e], v[8 * e + 1], v[8 * e + 2], v[8 * e + 3]),
                          make_float4(v[8 * e + 4], v[8 * e + 5], v[8 * e + 6], v[8 * e + 7]), hi, lo);
              const uint32_t off = uint32_t(lane * 64 + ((e ^ ((lane >> 1) & 3)) << 4));
              *reinterpret_cast<uint4*>(qbox + off) = hi;
              *reinterpret_cast<uint4*>(qbox + 2048 + off) = lo;
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              tma_store_3d(&P.map_p[g], qbox_addr, col0 + 32 * i, int(row0), 0);
              tma_store_3d(&P.map_p[g], qbox_addr + 2048u, col0 + 32 * i, int(row0), 1);
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
          }
          if (tid == 0 && ti == 0 && j < 8) trace(50 + j);
        }
        // hA takes the next tile's ctx (and the boxes their next contents) once everything has been read out
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
        tc_fence_before();
        epi_bar_sync();
        if (tid == 0) mbar_arrive(tile_done);
        if (tid == 0 && ti == 0) trace(8);
        continue;
      }
      // ---- hidden chunks: accB[j & 1] <- packed hi | lo of GELU(acc + b1), in place (64 columns per warp)
#pragma unroll 1
      for (int j = 0; j < nchunk; ++j) {
        const uint32_t b = uint32_t(j & 1);
        mbar_wait(fc1_full + 8u * b, (fphase >> b) & 1u);
        fphase ^= 1u << b;
        tc_fence_after();
        if (tid == 0 && ti == 0 && j < 8) trace(40 + j);
        const uint32_t ca = tmem + kAccB + b * 128u + lane_addr + uint32_t(hf * 64);
        const float* bias = Q.b1 + chunk_of(j) * 128 + hf * 64;
        // 32 columns at a time, rewritten in place: the 16 fp32 columns of MMA k-step q become 8 words of hi pairs
        // followed by 8 words of lo pairs.  A rolled loop on purpose: the fully unrolled 64-column body was 23 KB of
        // straight-line code that every warp fetched cold (stall_no_inst was the top stall reason of this phase).
#pragma unroll 1
        for (int i = 0; i < 2; ++i) {
          float v[32];
          tc_ld32(ca + uint32_t(32 * i), v);
          add_cols_g(v, bias + 32 * i);
          uint32_t w[32];
#pragma unroll
          for (int q = 0; q < 2; ++q) {
#pragma unroll
            for (int e = 0; e < 8; ++e)
              split_pair_rt(gelu_fast(v[16 * q + 2 * e]), gelu_fast(v[16 * q + 2 * e + 1]), FMT, w[16 * q + e], w[16 * q + 8 + e]);
          }
          tc_st32(ca + uint32_t(32 * i), reinterpret_cast<const float*>(w));
        }
        tc_fence_before();
        mbar_arrive(g_ready + 8u * b);
        if (tid == 0 && ti == 0 && j < 8) trace(50 + j);
      }

      // ---- LayerNorm2 over accO -> y (TMA stores out of per-warp boxes in hA: all MMAs of the tile have retired)
      mbar_wait(out_full, tpar);
      tc_fence_after();
      if (tid == 0 && ti == 0) trace(7);
      if constexpr (CL > 1) {
        // Both CTAs hold a partial sum of the whole 128 x 256 tile (CTA 0 on top of h + b2, CTA 1 on top of zero).
        // CTA r owns columns [128 r, 128 r + 128) of LayerNorm2 and of the output: each CTA ships the OTHER half of
        // its partial sum to the peer (64 KB each way, at the same time - the one-way 128 KB transfer took 9 k
        // cycles while CTA 1 idled), both normalise and store 64 KB.  hA (idle: every MMA of this CTA has retired)
        // is cut in two: [0, 64 KB) landing zone - the peer's copies, later this CTA's output boxes; [64 KB, 128 KB)
        // staging zone - the boxes the peer's columns are copied from, later the peer's row statistics.
        const uint32_t peer = rank ^ 1u, wq = uint32_t(warp - 2);
        const uint32_t land = wq * 8192u, stag = 65536u + wq * 8192u;  // this warp's two 32 x 32 fp32 boxes in each zone
        float v0[32], v1[32];
        {
          const uint32_t sc = peer * 128u + uint32_t(hf) * 64u;  // the peer's columns this warp ships
          tc_ld32(tmem + kAccO + lane_addr + sc, v0);
          tc_ld32(tmem + kAccO + lane_addr + sc + 32u, v1);
          uint8_t* box = sm + stag;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint32_t o = uint32_t(lane * 128 + ((j ^ (lane & 7)) << 4));
            *reinterpret_cast<float4*>(box + o) = make_float4(v0[4 * j], v0[4 * j + 1], v0[4 * j + 2], v0[4 * j + 3]);
            *reinterpret_cast<float4*>(box + 4096 + o) = make_float4(v1[4 * j], v1[4 * j + 1], v1[4 * j + 2], v1[4 * j + 3]);
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (tid == 0 && ti == 0) trace(58);
        if (tid == 0) {
          mbar_expect_tx(partial_full, 65536u);  // 8 bulk copies of 8 KB complete it
          mbar_arrive_remote(map_to_peer(peer_free, peer));  // this CTA's landing zone may be written
        }
        mbar_wait_cluster(peer_free, tpar);
        if (tid == 0 && ti == 0) trace(59);
        if (lane == 0) {
          asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                           map_to_peer(hA + land, peer)),
                       "r"(hA + stag), "r"(8192u), "r"(map_to_peer(partial_full, peer))
                       : "memory");
        }
        const uint32_t oc = rank * 128u + uint32_t(hf) * 64u;  // this warp's 64 owned columns
        tc_ld32(tmem + kAccO + lane_addr + oc, v0);
        tc_ld32(tmem + kAccO + lane_addr + oc + 32u, v1);
        mbar_wait_cluster(partial_full, tpar);
        // the peer's partial sum of the owned columns has landed, hence the peer's staging boxes have been read
        if (tid == 0) mbar_arrive_remote(map_to_peer(xfer_done, peer));
        {
          const uint8_t* box = sm + land;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint32_t o = uint32_t(lane * 128 + ((j ^ (lane & 7)) << 4));
            const float4 a = *reinterpret_cast<const float4*>(box + o), b = *reinterpret_cast<const float4*>(box + 4096 + o);
            v0[4 * j] += a.x, v0[4 * j + 1] += a.y, v0[4 * j + 2] += a.z, v0[4 * j + 3] += a.w;
            v1[4 * j] += b.x, v1[4 * j + 1] += b.y, v1[4 * j + 2] += b.z, v1[4 * j + 3] += b.w;
          }
        }
        if (tid == 0 && ti == 0) trace(60);
        // statistics: 64 columns per thread -> the CTA's 128 (partner warp, through TMEM) -> the row's 256 (peer CTA)
        const float shift = v0[0];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float d0 = v0[j] - shift, d1 = v1[j] - shift;
          s1 += d0 + d1;
          s2 = fmaf(d0, d0, fmaf(d1, d1, s2));
        }
        const float dm = s1 * (1.0f / 64.0f);
        const float my_mean = shift + dm, my_m2 = fmaxf(s2 - s1 * dm, 0.f);
        tc_st2(tmem_x + uint32_t(hf * 2), my_mean, my_m2);
        tc_fence_before();
        epi_bar_sync();
        tc_fence_after();
        float om, oq;
        tc_ld2(tmem_x + uint32_t((hf ^ 1) * 2), om, oq);
        const float mean_c = 0.5f * (my_mean + om);
        const float m2_c = my_m2 + oq + 64.0f * ((my_mean - mean_c) * (my_mean - mean_c) + (om - mean_c) * (om - mean_c));
        // the peer's staging zone is free (its boxes were read before this CTA's partial_full completed): row statistics go there
        if (hf == 0) {
          st_peer_f32x2(hA + 65536u + uint32_t(row) * 8u, peer, mean_c, m2_c);
          mbar_arrive_remote(map_to_peer(stats_full, peer));
        }
        mbar_wait_cluster(stats_full, tpar);
        float2 ps;  // written by the peer through DSMEM: the acquire above orders this read behind it
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(ps.x), "=f"(ps.y) : "r"(hA + 65536u + uint32_t(row) * 8u) : "memory");
        const float mean = 0.5f * (mean_c + ps.x);
        const float m2 = m2_c + ps.y + 128.0f * ((mean_c - mean) * (mean_c - mean) + (ps.x - mean) * (ps.x - mean));
        const float rstd = rsqrtf(m2 * (1.0f / float(DM)) + P.eps);
        if (tid == 0 && ti == 0) trace(61);
        const uint32_t poff = Q.y ? 4096u : 0u;
        const float2 nm2 = make_float2(-mean, -mean), rs2 = make_float2(rstd, rstd);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          float* v = i == 0 ? v0 : v1;
          const int cl = int(oc) + i * 32;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 gg = __ldg(reinterpret_cast<const float4*>(Q.g2 + cl + j));
            const float4 bb = __ldg(reinterpret_cast<const float4*>(Q.be2 + cl + j));
            const float2 y0 = __ffma2_rn(__fmul2_rn(__fadd2_rn(make_float2(v[j], v[j + 1]), nm2), rs2), make_float2(gg.x, gg.y), make_float2(bb.x, bb.y));
            const float2 y1 = __ffma2_rn(__fmul2_rn(__fadd2_rn(make_float2(v[j + 2], v[j + 3]), nm2), rs2), make_float2(gg.z, gg.w), make_float2(bb.z, bb.w));
            v[j] = y0.x, v[j + 1] = y0.y, v[j + 2] = y1.x, v[j + 3] = y1.y;
          }
          // output boxes in this warp's landing boxes (read above): planes only - one 4 KB box pair per chunk; with an
          // fp32 output as well the 8 KB buffer is reused after the first chunk's boxes have been read out
          const uint32_t buf = Q.y ? 0u : uint32_t(i) * 4096u;
          if (Q.y && i == 1) {
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            __syncwarp();
          }
          uint8_t* box = sm + land + buf;
          if (Q.y) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              *reinterpret_cast<float4*>(box + lane * 128 + ((j ^ (lane & 7)) << 4)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          }
          if (Q.y_planes) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 hi, lo;
              split8<FMT>(make_float4(v[8 * j], v[8 * j + 1], v[8 * j + 2], v[8 * j + 3]),
                          make_float4(v[8 * j + 4], v[8 * j + 5], v[8 * j + 6], v[8 * j + 7]), hi, lo);
              const uint32_t off = uint32_t(lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4));
              *reinterpret_cast<uint4*>(box + poff + off) = hi;
              *reinterpret_cast<uint4*>(box + poff + 2048 + off) = lo;
            }
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            const uint32_t src = hA + land + buf;
            if (Q.y) tma_store_2d(&P.map_y[g], src, cl, int(row0));
            if (Q.y_planes) {
              tma_store_3d(&P.map_p[g], src + poff, cl, int(row0), 0);
              tma_store_3d(&P.map_p[g], src + poff + 2048u, cl, int(row0), 1);
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
        if (tid == 0 && ti == 0) trace(62);
        // hA takes the next tile's ctx once the output boxes have been read out and the peer has copied the staging boxes
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
        mbar_wait_cluster(xfer_done, tpar);
        tc_fence_before();
        epi_bar_sync();
        if (tid == 0) mbar_arrive(tile_done);
        if (tid == 0 && ti == 0) trace(8);
        continue;
      }
      {
        float v[32];
        float shift = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll 1
        for (int i = 0; i < 4; ++i) {
          tc_ld32(tmem + kAccO + lane_addr + hf * 128 + i * 32, v);
          if (i == 0) shift = v[0];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float d = v[j] - shift;
            s1 += d;
            s2 = fmaf(d, d, s2);
          }
        }
        const float2 mr = combine_stats(tmem_x, hf, shift, s1, s2, P.eps);
        const float2 nm2 = make_float2(-mr.x, -mr.x), rs2 = make_float2(mr.y, mr.y);
        const uint32_t obase = uint32_t(warp - 2) * 16384u;  // 2 x 8 KB: [fp32 box 4 KB][hi 2 KB][lo 2 KB]
        const uint32_t poff = Q.y ? 4096u : 0u;
#pragma unroll 1
        for (int i = 0; i < 4; ++i) {
          const int cl = hf * 128 + i * 32;
          tc_ld32(tmem + kAccO + lane_addr + cl, v);
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 gg = __ldg(reinterpret_cast<const float4*>(Q.g2 + cl + j));
            const float4 bb = __ldg(reinterpret_cast<const float4*>(Q.be2 + cl + j));
            const float2 y0 = __ffma2_rn(__fmul2_rn(__fadd2_rn(make_float2(v[j], v[j + 1]), nm2), rs2), make_float2(gg.x, gg.y), make_float2(bb.x, bb.y));
            const float2 y1 = __ffma2_rn(__fmul2_rn(__fadd2_rn(make_float2(v[j + 2], v[j + 3]), nm2), rs2), make_float2(gg.z, gg.w), make_float2(bb.z, bb.w));
            v[j] = y0.x, v[j + 1] = y0.y, v[j + 2] = y1.x, v[j + 3] = y1.y;
          }
          const uint32_t buf = uint32_t(i & 1) * 8192u;
          if (i >= 2) {  // the box written two chunks ago must have been read out by the TMA engine
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            __syncwarp();
          }
          uint8_t* box = sm + obase + buf;
          if (Q.y) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              *reinterpret_cast<float4*>(box + lane * 128 + ((j ^ (lane & 7)) << 4)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          }
          if (Q.y_planes) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 hi, lo;
              split8<FMT>(make_float4(v[8 * j], v[8 * j + 1], v[8 * j + 2], v[8 * j + 3]),
                          make_float4(v[8 * j + 4], v[8 * j + 5], v[8 * j + 6], v[8 * j + 7]), hi, lo);
              const uint32_t off = uint32_t(lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4));
              *reinterpret_cast<uint4*>(box + poff + off) = hi;
              *reinterpret_cast<uint4*>(box + poff + 2048 + off) = lo;
            }
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            const uint32_t src = hA + obase + buf;
            if (Q.y) tma_store_2d(&P.map_y[g], src, cl, int(row0));
            if (Q.y_planes) {
              tma_store_3d(&P.map_p[g], src + poff, cl, int(row0), 0);
              tma_store_3d(&P.map_p[g], src + poff + 2048u, cl, int(row0), 1);
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
        // the boxes must have been READ out of hA before the next tile's ctx lands there (or the CTA exits)
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
        epi_bar_sync();
        if (tid == 0) mbar_arrive(tile_done);
      }
      if (tid == 0 && ti == 0) trace(8);
    }
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (CL > 1) cluster_sync_all();  // no CTA leaves while its peer may still touch its shared memory
  if (tracing) {
    for (int i = threadIdx.x; i < kTraceSlots; i += kThreads) g_trace_blk[i] = (long long)trp[i];
  }
  if (warp == 1) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");
  }
}

}  // namespace

// 0 = pick by the number of row tiles, 1 | 2 = force the schedule (SCATT_BLOCK_CLUSTER or scatt_debug_set_block_cluster: A/B runs, tests)
static std::atomic<int> g_block_cluster{[] {
  const char* e = std::getenv("SCATT_BLOCK_CLUSTER");
  return e ? std::atoi(e) : 0;
}()};
int debug_set_block_cluster(int cl) {
  SCATT_REQUIRE(cl >= 0 && cl <= 2, "debug_set_block_cluster: 0 (automatic), 1 or 2");
  g_block_cluster.store(cl, std::memory_order_relaxed);
  return SCATT_OK;
}

int debug_set_trace_block(void* dev_buf) {
  long long* p = reinterpret_cast<long long*>(dev_buf);
  SCATT_CUDA(cudaMemcpyToSymbol(g_trace_blk, &p, sizeof(p)));
  return SCATT_OK;
}

bool attn_block_supported(int64_t M, int D, int F) { return D == DM && F >= 128 && F <= kMaxF && F % 128 == 0 && M >= 1 && M < (int64_t(1) << 31); }

int launch_attn_block(const scatt_block_problem* p, int group, int64_t M, int D, int F, float eps, int fmt, int terms, cudaStream_t s) {
  const int cluster_override = g_block_cluster.load(std::memory_order_relaxed);
  SCATT_REQUIRE(terms >= 1 && terms <= 3, "attn_block: terms must be 1, 2 or 3");
  if (M == 0) return SCATT_OK;
  SCATT_REQUIRE(attn_block_supported(M, D, F), "attn_block: needs D = %d and F a multiple of 128 in [128, %d] (got D=%d F=%d)", DM, kMaxF, D, F);
  BlkParams P{};
  P.M = M, P.F = F, P.nchunk = F / 128, P.terms = terms, P.fmt = fmt, P.groups = group, P.eps = eps;
  P.tiles_m = int((M + BM - 1) / BM);
  for (int i = 0; i < group; ++i) {
    const scatt_block_problem& a = p[i];
    SCATT_REQUIRE(a.ctx_planes && a.residual_planes && a.wo_planes && a.w1_planes && a.w2_planes, "attn_block: problem %d lacks an operand", i);
    SCATT_REQUIRE(a.bo && a.ln1_g && a.ln1_b && a.b1 && a.b2 && a.ln2_g && a.ln2_b, "attn_block: problem %d lacks a bias or LayerNorm parameter", i);
    SCATT_REQUIRE(a.y || a.y_planes, "attn_block: no output");
    const uintptr_t al = reinterpret_cast<uintptr_t>(a.bo) | reinterpret_cast<uintptr_t>(a.ln1_g) | reinterpret_cast<uintptr_t>(a.ln1_b) |
                         reinterpret_cast<uintptr_t>(a.b1) | reinterpret_cast<uintptr_t>(a.b2) | reinterpret_cast<uintptr_t>(a.ln2_g) |
                         reinterpret_cast<uintptr_t>(a.ln2_b);
    SCATT_REQUIRE((al & 15) == 0, "attn_block: biases and LayerNorm parameters must be 16-byte aligned");
    int rc = encode_planes_map(&P.map_ctx[i], a.ctx_planes, M, DM, BM, fmt, terms >= 2 ? 2 : 1);
    if (rc == SCATT_OK) rc = encode_planes_map(&P.map_x[i], a.residual_planes, M, DM, BM, fmt, 2);
    if (rc == SCATT_OK) rc = encode_planes_map(&P.map_wo[i], a.wo_planes, DM, DM, 256, fmt, 1);
    if (rc == SCATT_OK) rc = encode_planes_map(&P.map_w1[i], a.w1_planes, F, DM, 128, fmt, terms >= 3 ? 2 : 1);
    if (rc == SCATT_OK) rc = encode_planes_map(&P.map_w2[i], a.w2_planes, DM, F, 256, fmt, 1);
    if (rc == SCATT_OK) rc = encode_out_maps(&P.map_y[i], &P.map_p[i], a.y, DM, a.y_planes, M, DM, fmt);
    if (rc != SCATT_OK) return rc;
    P.prob[i] = BlkProblem{a.bo, a.ln1_g, a.ln1_b, a.b1, a.b2, a.ln2_g, a.ln2_b, a.y, reinterpret_cast<uint16_t*>(a.y_planes)};
  }
  static PerDeviceFlag attr_done;
  if (!attr_done.load()) {
    SCATT_CUDA(cudaFuncSetAttribute(attn_block_kernel<SCATT_PLANE_F16, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kSmemBytes)));
    SCATT_CUDA(cudaFuncSetAttribute(attn_block_kernel<SCATT_PLANE_BF16, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kSmemBytes)));
    SCATT_CUDA(cudaFuncSetAttribute(attn_block_kernel<SCATT_PLANE_F16, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kSmemBytes)));
    SCATT_CUDA(cudaFuncSetAttribute(attn_block_kernel<SCATT_PLANE_BF16, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kSmemBytes)));
    attr_done.store(true);
  }
  const int tiles = P.tiles_m * group;
  // few row tiles (small batches): a 2-CTA cluster per tile, each CTA half of the hidden chunks - twice the SMs share the
  // weight stream and the fc1 / fc2 MMAs; from 75 tiles on every SM has a tile of its own anyway
  const int cl = cluster_override > 0 ? cluster_override : ((tiles <= 74 && P.nchunk >= 2) ? 2 : 1);
  if (cl == 2) {
    dim3 grid(unsigned(2 * (tiles < 74 ? tiles : 74)));
    if (fmt == SCATT_PLANE_F16) (void)launch_kernel_cluster(attn_block_kernel<SCATT_PLANE_F16, 2>, grid, dim3(kThreads), kSmemBytes, s, 2, P);
    else (void)launch_kernel_cluster(attn_block_kernel<SCATT_PLANE_BF16, 2>, grid, dim3(kThreads), kSmemBytes, s, 2, P);
  } else {
    dim3 grid(unsigned(tiles < 148 ? tiles : 148));
    if (fmt == SCATT_PLANE_F16) (void)launch_kernel(attn_block_kernel<SCATT_PLANE_F16, 1>, grid, dim3(kThreads), kSmemBytes, s, P);
    else (void)launch_kernel(attn_block_kernel<SCATT_PLANE_BF16, 1>, grid, dim3(kThreads), kSmemBytes, s, P);
  }
  const int rc = after_launch("attn_block_kernel");
  set_last_kernel("attn_block_kernel<%d, %d, 0>", fmt, cl);
  return rc;
}

bool attn_out_q_supported(int64_t M, int D, int N) { return D == DM && N == DM && M >= 1 && M < (int64_t(1) << 31); }

// MODE 1 of attn_block_kernel: h = LayerNorm(x + ctx Wo^T + bo) -> h_planes, q = (h Wq^T + bq) * q_scale -> q_planes
int launch_attn_out_q(const scatt_outq_problem* p, int group, int64_t M, int D, int N, float eps, float q_scale, int fmt, int terms,
                      cudaStream_t s) {
  const int cluster_override = g_block_cluster.load(std::memory_order_relaxed);
  SCATT_REQUIRE(terms >= 1 && terms <= 3, "attn_out_q: terms must be 1, 2 or 3");
  if (M == 0) return SCATT_OK;
  SCATT_REQUIRE(attn_out_q_supported(M, D, N), "attn_out_q: needs D = N = %d (got D=%d N=%d)", DM, D, N);
  BlkParams P{};
  P.M = M, P.F = N, P.nchunk = N / 128, P.terms = terms, P.fmt = fmt, P.groups = group, P.eps = eps, P.qscale = q_scale;
  P.tiles_m = int((M + BM - 1) / BM);
  for (int i = 0; i < group; ++i) {
    const scatt_outq_problem& a = p[i];
    SCATT_REQUIRE(a.ctx_planes && a.residual_planes && a.wo_planes && a.wq_planes && a.h_planes && a.q_planes, "attn_out_q: problem %d lacks an operand or an output", i);
    SCATT_REQUIRE(a.bo && a.ln_g && a.ln_b && a.bq, "attn_out_q: problem %d lacks a bias or LayerNorm parameter", i);
    const uintptr_t al = reinterpret_cast<uintptr_t>(a.bo) | reinterpret_cast<uintptr_t>(a.ln_g) | reinterpret_cast<uintptr_t>(a.ln_b) | reinterpret_cast<uintptr_t>(a.bq);
    SCATT_REQUIRE((al & 15) == 0, "attn_out_q: biases and LayerNorm parameters must be 16-byte aligned");
    int rc = encode_planes_map(&P.map_ctx[i], a.ctx_planes, M, DM, BM, fmt, terms >= 2 ? 2 : 1);
    if (rc == SCATT_OK) rc = encode_planes_map(&P.map_x[i], a.residual_planes, M, DM, BM, fmt, 2);
    if (rc == SCATT_OK) rc = encode_planes_map(&P.map_wo[i], a.wo_planes, DM, DM, 256, fmt, 1);
    if (rc == SCATT_OK) rc = encode_planes_map(&P.map_w1[i], a.wq_planes, N, DM, 128, fmt, terms >= 3 ? 2 : 1);
    if (rc == SCATT_OK) rc = encode_planes_map(&P.map_y[i], a.h_planes, M, DM, BM, fmt, 2);  // h: k-block boxes, hi tile | lo tile
    if (rc == SCATT_OK) rc = encode_out_maps(nullptr, &P.map_p[i], nullptr, 0, a.q_planes, M, N, fmt);
    if (rc != SCATT_OK) return rc;
    P.map_w2[i] = P.map_wo[i];  // never used in this mode; a valid descriptor for the prefetch
    P.prob[i] = BlkProblem{a.bo, a.ln_g, a.ln_b, a.bq, a.bo, a.bo, a.bo, nullptr, reinterpret_cast<uint16_t*>(a.q_planes)};
  }
  static PerDeviceFlag attr_done;
  if (!attr_done.load()) {
    SCATT_CUDA(cudaFuncSetAttribute(attn_block_kernel<SCATT_PLANE_F16, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kSmemBytes)));
    SCATT_CUDA(cudaFuncSetAttribute(attn_block_kernel<SCATT_PLANE_BF16, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kSmemBytes)));
    SCATT_CUDA(cudaFuncSetAttribute(attn_block_kernel<SCATT_PLANE_F16, 2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kSmemBytes)));
    SCATT_CUDA(cudaFuncSetAttribute(attn_block_kernel<SCATT_PLANE_BF16, 2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kSmemBytes)));
    attr_done.store(true);
  }
  const int tiles = P.tiles_m * group;
  const int cl = cluster_override > 0 ? cluster_override : (tiles <= 74 ? 2 : 1);
  if (cl == 2) {
    dim3 grid(unsigned(2 * (tiles < 74 ? tiles : 74)));
    if (fmt == SCATT_PLANE_F16) (void)launch_kernel_cluster(attn_block_kernel<SCATT_PLANE_F16, 2, 1>, grid, dim3(kThreads), kSmemBytes, s, 2, P);
    else (void)launch_kernel_cluster(attn_block_kernel<SCATT_PLANE_BF16, 2, 1>, grid, dim3(kThreads), kSmemBytes, s, 2, P);
  } else {
    dim3 grid(unsigned(tiles < 148 ? tiles : 148));
    if (fmt == SCATT_PLANE_F16) (void)launch_kernel(attn_block_kernel<SCATT_PLANE_F16, 1, 1>, grid, dim3(kThreads), kSmemBytes, s, P);
    else (void)launch_kernel(attn_block_kernel<SCATT_PLANE_BF16, 1, 1>, grid, dim3(kThreads), kSmemBytes, s, P);
  }
  const int rc = after_launch("attn_block_kernel");
  set_last_kernel("attn_block_kernel<%d, %d, 1>", fmt, cl);
  return rc;
}

}  // namespace scatt
