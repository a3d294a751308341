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
//   [0, 256)    accO: out_proj accumulator, pre-initialised with bo + x by the epilogue warps while the operands
//               are in flight; after LayerNorm1 the same columns are rewritten with h + b2 and become the fc2
//               accumulator, so both residual adds cost no extra pass
//   [256, 384), [384, 512)   two fc1 chunk accumulators (128 hidden columns each).  The epilogue warps apply
//               bias + GELU and write the chunk back IN PLACE as packed 16-bit hi | lo - it is then the TMEM A
//               operand of the fc2 MMAs of that chunk (like P in the attention kernel): the hidden activation
//               never touches shared memory.  While chunk j is being activated the MMA warp computes chunk j + 1.
// Shared memory:
//   hA   128 KB: ctx tile (K-major, 128-byte swizzle, 4 k-blocks x hi | lo) = A operand of out_proj; LayerNorm1 then
//               overwrites it with h in the same layout = A operand of fc1; at the end it holds the output boxes the
//               TMA engine stores.  The last 32 KB double as the transposition tiles of the accumulator
//               pre-initialisation (ctx k-block 3 is fetched when that is done).
//   ring 5 x 16 KB: weight tiles [128 rows x 64 K] of one plane, streamed in MMA order
//               (Wo, then W1 chunk 0, W1 chunk 1, W2 chunk 0, W1 chunk 2, W2 chunk 1, ...), full / empty mbarriers.
// Per tile the tensor pipe sees 2 * 128 * (D*D + 2*D*F) * terms flops = 43 k cycles at F = 768, terms = 3; the
// weights (D*D + 2*D*F) * 4 B = 1.75 MB stream from L2 at ~42 B / clk.
// The kernel is persistent: CTA b walks tiles b, b + grid, ...; the weight ring runs ahead across tile boundaries.
#include <cuda.h>

#include "common.cuh"
#include "tc_epi.cuh"
#include "tc_host.cuh"
#include "tc_ptx.cuh"

namespace scatt {

namespace {

using namespace tc;

constexpr int BM = 128;
constexpr int DM = 256;          // model width (columns of ctx, x, h, y)
constexpr int kMaxF = 1024;      // widest hidden layer the shared-memory map provides for
constexpr int kEpiWarps = 8;
constexpr int kThreads = 64 + 32 * kEpiWarps;
constexpr int kSlots = 5;
constexpr uint32_t kSlotBytes = 16384;   // [128 x 64] 16-bit tile
constexpr uint32_t kKbBytes = 32768;     // one A k-block: hi tile + lo tile
constexpr uint32_t kHABytes = 4 * kKbBytes;
constexpr uint32_t kStageOff = 3 * kKbBytes;  // pre-initialisation transposition tiles (inside ctx k-block 3)
constexpr int kStageLd = 20;
constexpr uint32_t kStageWarpBytes = 32 * kStageLd * 4;
constexpr uint32_t kTmemCols = 512, kAccO = 0, kAccB = 256;

struct BlkProblem {
  const float *bo, *g1, *be1, *b1, *b2, *g2, *be2;
  const uint16_t* res_planes;
  float* y;
  uint16_t* y_planes;
};

struct alignas(64) BlkParams {
  CUtensorMap map_ctx[SCATT_MAX_GROUP];
  CUtensorMap map_wo[SCATT_MAX_GROUP];
  CUtensorMap map_w1[SCATT_MAX_GROUP];
  CUtensorMap map_w2[SCATT_MAX_GROUP];
  CUtensorMap map_y[SCATT_MAX_GROUP];
  CUtensorMap map_p[SCATT_MAX_GROUP];
  BlkProblem prob[SCATT_MAX_GROUP];
  int64_t M;
  int32_t F, nchunk, terms, fmt, groups, tiles_m;
  float eps;
};

// shared memory map relative to the 1024-aligned base
constexpr uint32_t kRingOff = kHABytes;
constexpr uint32_t kBarOff = kRingOff + kSlots * kSlotBytes;
constexpr uint32_t kColOff = kBarOff + 256;                       // float[6 * 256 + kMaxF]
constexpr uint32_t kStatsOff = kColOff + (6 * DM + kMaxF) * 4;    // float2[2][128]
constexpr uint32_t kSmemBytes = kStatsOff + 2 * BM * 8 + 1024;    // + alignment slack
static_assert(kSmemBytes <= 227 * 1024, "attn_block: shared memory map exceeds 227 KB");
static_assert(kStageOff + kEpiWarps * kStageWarpBytes <= kHABytes, "staging tiles must fit in ctx k-block 3");

__device__ long long* g_trace_blk = nullptr;
__device__ __forceinline__ void trace(int slot) {
  if (g_trace_blk != nullptr && blockIdx.x == 0) g_trace_blk[slot] = clock64();
}

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory"); }

__device__ __forceinline__ void tc_mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// LayerNorm statistics of 128 columns held by one thread-row, combined with the partner warp's 128 columns
// (Chan et al.); returns (mean, rstd) over the 256 columns.
__device__ __forceinline__ float2 combine_stats(float2* stats, int hf, int row, float shift, float s1, float s2, float eps) {
  constexpr float kHalfN = float(DM / 2);
  const float dm = s1 / kHalfN;
  const float my_mean = shift + dm, my_m2 = fmaxf(s2 - s1 * dm, 0.f);
  stats[hf * BM + row] = make_float2(my_mean, my_m2);
  epi_bar_sync();
  const float2 other = stats[(hf ^ 1) * BM + row];
  const float mean = 0.5f * (my_mean + other.x);
  const float da = my_mean - mean, db = other.x - mean;
  const float m2 = my_m2 + other.y + kHalfN * (da * da + db * db);
  return make_float2(mean, rsqrtf(m2 / float(DM) + eps));
}

template <int FMT>
__global__ void __launch_bounds__(kThreads, 1) attn_block_kernel(const __grid_constant__ BlkParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  const uint32_t hA = base, ring = base + kRingOff, bar0 = base + kBarOff;
  auto full_bar = [&](uint32_t s) { return bar0 + 8u * s; };
  auto empty_bar = [&](uint32_t s) { return bar0 + 8u * (kSlots + s); };
  const uint32_t ctx_bar = bar0 + 8u * 2 * kSlots;  // [4], one per ctx k-block
  const uint32_t acc_init_bar = ctx_bar + 32u, oproj_full = acc_init_bar + 8u, h_ready = oproj_full + 8u;
  const uint32_t fc1_full = h_ready + 8u;   // [2]
  const uint32_t g_ready = fc1_full + 16u;  // [2]
  const uint32_t out_full = g_ready + 16u, tile_done = out_full + 8u, tmem_ptr_addr = tile_done + 8u;
  float* col = reinterpret_cast<float*>(sm + kColOff);  // bo | g1 | be1 | b2 | g2 | be2 | b1[F]
  float2* stats = reinterpret_cast<float2*>(sm + kStatsOff);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nchunk = P.nchunk;
  const int total_tiles = P.tiles_m * P.groups;
  const bool a_lo = P.terms >= 2, b_lo = P.terms >= 3;

  if (threadIdx.x == 0) {
    for (uint32_t s = 0; s < kSlots; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (uint32_t k = 0; k < 4; ++k) mbar_init(ctx_bar + 8u * k, 1);
    mbar_init(acc_init_bar, 32 * kEpiWarps);
    mbar_init(oproj_full, 1);
    mbar_init(h_ready, 32 * kEpiWarps);
    for (uint32_t b = 0; b < 2; ++b) {
      mbar_init(fc1_full + 8u * b, 1);
      mbar_init(g_ready + 8u * b, 32 * kEpiWarps);
    }
    mbar_init(out_full, 1);
    mbar_init(tile_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int g = 0; g < P.groups; ++g) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&P.map_ctx[g]) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&P.map_wo[g]) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&P.map_w1[g]) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&P.map_w2[g]) : "memory");
    }
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_addr), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_launch_dependents();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(sm + (tmem_ptr_addr - base));
  if (threadIdx.x == 0) trace(0);

  if (warp == 0) {  // ================================================= TMA producer
    uint32_t it = 0;  // ring items issued so far (the ring runs across tile boundaries)
    auto put = [&](const CUtensorMap* map, int k0, int n0) {  // weight rows [n0, n0 + 128) x K [k0, k0 + 64): hi (, lo)
      for (int pl = 0; pl < (b_lo ? 2 : 1); ++pl, ++it) {
        const uint32_t s = it % kSlots;
        mbar_wait(empty_bar(s), ((it / kSlots) & 1u) ^ 1u);
        if (elect_one()) {
          mbar_expect_tx(full_bar(s), kSlotBytes);
          tma_load_3d(ring + s * kSlotBytes, map, full_bar(s), k0, n0, pl);
        }
        __syncwarp();
      }
    };
    auto put_ctx = [&](int g, int m0, int kb) {
      if (elect_one()) {
        mbar_expect_tx(ctx_bar + 8u * kb, a_lo ? kKbBytes : kSlotBytes);
        tma_load_3d(hA + kb * kKbBytes, &P.map_ctx[g], ctx_bar + 8u * kb, kb * 64, m0, 0);
        if (a_lo) tma_load_3d(hA + kb * kKbBytes + kSlotBytes, &P.map_ctx[g], ctx_bar + 8u * kb, kb * 64, m0, 1);
      }
      __syncwarp();
    };
    int ti = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++ti) {
      const int g = t / P.tiles_m, m0 = (t % P.tiles_m) * BM;
      // the first Wo k-block is a static weight: it may be fetched before the activations exist
      put(&P.map_wo[g], 0, 0);
      put(&P.map_wo[g], 0, 128);
      if (ti == 0) pdl_wait();                                     // ctx / x come from the preceding kernels
      else mbar_wait(tile_done, uint32_t(ti - 1) & 1u);            // the previous tile's output boxes have left hA
      for (int kb = 0; kb < 3; ++kb) put_ctx(g, m0, kb);
      mbar_wait(acc_init_bar, uint32_t(ti) & 1u);                  // the transposition tiles inside k-block 3 are idle
      put_ctx(g, m0, 3);
      for (int kb = 1; kb < 4; ++kb) {
        put(&P.map_wo[g], kb * 64, 0);
        put(&P.map_wo[g], kb * 64, 128);
      }
      auto fc1_items = [&](int j) {
        for (int kb = 0; kb < 4; ++kb) put(&P.map_w1[g], kb * 64, j * 128);
      };
      auto fc2_items = [&](int j) {
        for (int kb = 0; kb < 2; ++kb) {
          put(&P.map_w2[g], (2 * j + kb) * 64, 0);
          put(&P.map_w2[g], (2 * j + kb) * 64, 128);
        }
      };
      fc1_items(0);
      for (int j = 1; j < nchunk; ++j) {
        fc1_items(j);
        fc2_items(j - 1);
      }
      fc2_items(nchunk - 1);
    }
  } else if (warp == 1) {  // ========================================== MMA issuer
    const uint32_t idesc = (1u << 4) | (uint32_t(FMT) << 7) | (uint32_t(FMT) << 10) | (uint32_t(128 >> 3) << 17) | (uint32_t(BM >> 4) << 24);
    uint32_t it = 0;
    uint32_t gphase = 0;  // bit b: parity of the next completion of g_ready[b]
    // waits for the next weight tile(s) of the ring; returns the slots of the hi and lo planes
    auto take = [&](uint32_t& s_hi, uint32_t& s_lo) {
      s_hi = it % kSlots;
      mbar_wait(full_bar(s_hi), (it / kSlots) & 1u);
      ++it;
      s_lo = s_hi;
      if (b_lo) {
        s_lo = it % kSlots;
        mbar_wait(full_bar(s_lo), (it / kSlots) & 1u);
        ++it;
      }
      tc_fence_after();
    };
    int ti = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++ti) {
      const uint32_t tpar = uint32_t(ti) & 1u;
      mbar_wait(acc_init_bar, tpar);  // accO holds bo + x
      tc_fence_after();
      if (lane == 0 && ti == 0) trace(1);
      // ---- out_proj: accO[128 x 256] += ctx Wo^T, two 128-column halves per k-block
      for (int kb = 0; kb < 4; ++kb) {
        mbar_wait(ctx_bar + 8u * kb, tpar);
        tc_fence_after();
        const uint64_t ah = umma_desc_sw128(hA + kb * kKbBytes), al = umma_desc_sw128(hA + kb * kKbBytes + kSlotBytes);
        for (int nh = 0; nh < 2; ++nh) {
          uint32_t s_hi, s_lo;
          take(s_hi, s_lo);
          const uint64_t bh = umma_desc_sw128(ring + s_hi * kSlotBytes), bl = umma_desc_sw128(ring + s_lo * kSlotBytes);
          const uint32_t d = tmem + kAccO + uint32_t(nh * 128);
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              const uint64_t adv = uint64_t(kk * 2);
              if (b_lo) tc_mma_f16(d, ah + adv, bl + adv, idesc, 1);
              if (a_lo) tc_mma_f16(d, al + adv, bh + adv, idesc, 1);
              tc_mma_f16(d, ah + adv, bh + adv, idesc, 1);
            }
            tc_commit(empty_bar(s_hi));
            if (b_lo) tc_commit(empty_bar(s_lo));
            if (kb == 3 && nh == 1) tc_commit(oproj_full);
          }
          __syncwarp();
        }
      }
      if (lane == 0 && ti == 0) trace(2);
      mbar_wait(h_ready, tpar);  // h is in hA (A operand of fc1), h + b2 in accO
      tc_fence_after();
      if (lane == 0 && ti == 0) trace(3);
      auto issue_fc1 = [&](int j) {  // accB[j & 1][128 x 128] = h W1[chunk j]^T
        const uint32_t d = tmem + kAccB + uint32_t((j & 1) * 128);
        for (int kb = 0; kb < 4; ++kb) {
          uint32_t s_hi, s_lo;
          take(s_hi, s_lo);
          const uint64_t ah = umma_desc_sw128(hA + kb * kKbBytes), al = umma_desc_sw128(hA + kb * kKbBytes + kSlotBytes);
          const uint64_t bh = umma_desc_sw128(ring + s_hi * kSlotBytes), bl = umma_desc_sw128(ring + s_lo * kSlotBytes);
          if (elect_one()) {
            uint32_t acc = kb > 0 ? 1u : 0u;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              const uint64_t adv = uint64_t(kk * 2);
              if (b_lo) {
                tc_mma_f16(d, ah + adv, bl + adv, idesc, acc);
                acc = 1;
              }
              if (a_lo) {
                tc_mma_f16(d, al + adv, bh + adv, idesc, acc);
                acc = 1;
              }
              tc_mma_f16(d, ah + adv, bh + adv, idesc, acc);
              acc = 1;
            }
            tc_commit(empty_bar(s_hi));
            if (b_lo) tc_commit(empty_bar(s_lo));
            if (kb == 3) tc_commit(fc1_full + 8u * uint32_t(j & 1));
          }
          __syncwarp();
        }
      };
      auto issue_fc2 = [&](int j, bool last) {  // accO += g[chunk j] W2[:, chunk j]^T, A = g from TMEM
        const uint32_t b = uint32_t(j & 1);
        mbar_wait(g_ready + 8u * b, (gphase >> b) & 1u);
        gphase ^= 1u << b;
        tc_fence_after();
        const uint32_t gbase = tmem + kAccB + b * 128u;
        for (int kb = 0; kb < 2; ++kb) {
          for (int nh = 0; nh < 2; ++nh) {
            uint32_t s_hi, s_lo;
            take(s_hi, s_lo);
            const uint64_t bh = umma_desc_sw128(ring + s_hi * kSlotBytes), bl = umma_desc_sw128(ring + s_lo * kSlotBytes);
            const uint32_t d = tmem + kAccO + uint32_t(nh * 128);
            if (elect_one()) {
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) {
                const uint64_t adv = uint64_t(kk * 2);
                const uint32_t g_hi = gbase + uint32_t(kb * 64 + kk * 8), g_lo = g_hi + 32u;
                if (b_lo) tc_mma_ts(d, g_hi, bl + adv, idesc, 1);
                if (a_lo) tc_mma_ts(d, g_lo, bh + adv, idesc, 1);
                tc_mma_ts(d, g_hi, bh + adv, idesc, 1);
              }
              tc_commit(empty_bar(s_hi));
              if (b_lo) tc_commit(empty_bar(s_lo));
              if (last && kb == 1 && nh == 1) tc_commit(out_full);
            }
            __syncwarp();
          }
        }
      };
      // tcgen05.mma executes in issue order: fc1(j + 1) can be issued before fc2(j - 1) has read its operand chunk
      // because it writes the OTHER accumulator, and fc1(j + 2) after fc2(j) reuses that chunk's columns safely
      issue_fc1(0);
      for (int j = 1; j < nchunk; ++j) {
        issue_fc1(j);
        issue_fc2(j - 1, false);
      }
      issue_fc2(nchunk - 1, true);
      if (lane == 0 && ti == 0) trace(4);
    }
  } else {  // ========================================================= epilogue warps 2..9
    const int quad = warp & 3, hf = (warp - 2) >> 2;
    const int row = quad * 32 + lane;
    const uint32_t lane_addr = uint32_t(quad * 32) << 16;
    const int tid = threadIdx.x - 64;
    float* stage = reinterpret_cast<float*>(sm + kStageOff) + (warp - 2) * (kStageWarpBytes / 4);
    const float *c_bo = col, *c_g1 = col + DM, *c_be1 = col + 2 * DM, *c_b2 = col + 3 * DM, *c_g2 = col + 4 * DM,
                *c_be2 = col + 5 * DM, *c_b1 = col + 6 * DM;
    uint32_t fphase = 0;  // bit b: parity of the next completion of fc1_full[b]
    int cur_g = -1, ti = 0;
    const int sub = lane >> 2, q4 = lane & 3;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++ti) {
      const uint32_t tpar = uint32_t(ti) & 1u;
      const int g = t / P.tiles_m;
      const int64_t m0 = int64_t(t % P.tiles_m) * BM;
      const BlkProblem& Q = P.prob[g];
      if (g != cur_g) {  // per-column parameters of this problem -> shared memory (static weights: before pdl_wait)
        if (cur_g >= 0) epi_bar_sync();
        col[tid] = Q.bo[tid], col[DM + tid] = Q.g1[tid], col[2 * DM + tid] = Q.be1[tid];
        col[3 * DM + tid] = Q.b2[tid], col[4 * DM + tid] = Q.g2[tid], col[5 * DM + tid] = Q.be2[tid];
        for (int i = tid; i < P.F; i += 32 * kEpiWarps) col[6 * DM + i] = Q.b1[i];
        epi_bar_sync();
        cur_g = g;
      }
      if (ti == 0) pdl_wait();
      const int64_t row0 = m0 + quad * 32;
      const int rows_valid = int(min(int64_t(32), max(int64_t(0), P.M - row0)));

      // ---- accO <- bo + x  (x = the layer input as split planes; coalesced fetch, transposed through `stage`)
      {
        const int64_t ps = P.M * int64_t(DM);
        float4 r[2][8];
        auto fetch = [&](int c0, float4(&dst)[8]) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int rr = 8 * i + sub;
            dst[i] = make_float4(0.f, 0.f, 0.f, 0.f), dst[4 + i] = dst[i];
            if (rr < rows_valid) {
              const uint16_t* p = Q.res_planes + (row0 + rr) * DM + c0 + q4 * 8;
              dst[i] = *reinterpret_cast<const float4*>(p);
              dst[4 + i] = *reinterpret_cast<const float4*>(p + ps);
            }
          }
        };
        fetch(hf * 128, r[0]);
#pragma unroll 2
        for (int i = 0; i < 4; ++i) {
          const int cl = hf * 128 + i * 32;
          if (i + 1 < 4) fetch(cl + 32, r[(i + 1) & 1]);
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0.f;
          add_cols(v, c_bo + cl);
          const float4(&rr)[8] = r[i & 1];
#pragma unroll
          for (int h = 0; h < 2; ++h) {  // 16-column halves through the transposition tile
            if ((q4 >> 1) == h) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const float2 p0 = unpack_pair(rr[k].x, rr[4 + k].x, FMT), p1 = unpack_pair(rr[k].y, rr[4 + k].y, FMT);
                const float2 p2 = unpack_pair(rr[k].z, rr[4 + k].z, FMT), p3 = unpack_pair(rr[k].w, rr[4 + k].w, FMT);
                float* dst = stage + (8 * k + sub) * kStageLd + (q4 & 1) * 8;
                *reinterpret_cast<float4*>(dst) = make_float4(p0.x, p0.y, p1.x, p1.y);
                *reinterpret_cast<float4*>(dst + 4) = make_float4(p2.x, p2.y, p3.x, p3.y);
              }
            }
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              const float4 tt = *reinterpret_cast<const float4*>(stage + lane * kStageLd + j);
              v[h * 16 + j] += tt.x, v[h * 16 + j + 1] += tt.y, v[h * 16 + j + 2] += tt.z, v[h * 16 + j + 3] += tt.w;
            }
            __syncwarp();
          }
          tc_st32(tmem + kAccO + lane_addr + cl, v);
        }
        tc_fence_before();
        mbar_arrive(acc_init_bar);
      }

      // ---- LayerNorm1 over accO -> h: planes into hA (fc1's A operand), h + b2 back into accO (fc2's accumulator)
      mbar_wait(oproj_full, tpar);
      tc_fence_after();
      if (tid == 0 && ti == 0) trace(5);
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
        const float2 mr = combine_stats(stats, hf, row, shift, s1, s2, P.eps);
        uint8_t* hrow = sm + (row >> 3) * 1024 + (row & 7) * 128;
#pragma unroll 1
        for (int i = 0; i < 4; ++i) {
          const int cl = hf * 128 + i * 32;
          tc_ld32(tmem + kAccO + lane_addr + cl, v);
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 gg = *reinterpret_cast<const float4*>(c_g1 + cl + j);
            const float4 bb = *reinterpret_cast<const float4*>(c_be1 + cl + j);
            v[j] = (v[j] - mr.x) * mr.y * gg.x + bb.x;
            v[j + 1] = (v[j + 1] - mr.x) * mr.y * gg.y + bb.y;
            v[j + 2] = (v[j + 2] - mr.x) * mr.y * gg.z + bb.z;
            v[j + 3] = (v[j + 3] - mr.x) * mr.y * gg.w + bb.w;
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
            *reinterpret_cast<uint4*>(kbp + kSlotBytes + off) = lo;
          }
          add_cols(v, c_b2 + cl);
          tc_st32(tmem + kAccO + lane_addr + cl, v);
        }
        fence_proxy_async();  // generic-proxy writes of h -> visible to the tensor core's operand reads
        tc_fence_before();
        mbar_arrive(h_ready);
      }
      if (tid == 0 && ti == 0) trace(6);

      // ---- hidden chunks: accB[j & 1] <- packed hi | lo of GELU(acc + b1), in place (64 columns per warp)
#pragma unroll 1
      for (int j = 0; j < nchunk; ++j) {
        const uint32_t b = uint32_t(j & 1);
        mbar_wait(fc1_full + 8u * b, (fphase >> b) & 1u);
        fphase ^= 1u << b;
        tc_fence_after();
        const uint32_t ca = tmem + kAccB + b * 128u + lane_addr + uint32_t(hf * 64);
        float va[32], vb[32];
        tc_ld32(ca, va);
        tc_ld32(ca + 32u, vb);
        const float* bias = c_b1 + j * 128 + hf * 64;
        add_cols(va, bias);
        add_cols(vb, bias + 32);
        uint32_t whi[32], wlo[32];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          split_pair_rt(gelu_fast(va[2 * e]), gelu_fast(va[2 * e + 1]), FMT, whi[e], wlo[e]);
          split_pair_rt(gelu_fast(vb[2 * e]), gelu_fast(vb[2 * e + 1]), FMT, whi[16 + e], wlo[16 + e]);
        }
        tc_st32(ca, reinterpret_cast<const float*>(whi));        // K elements [64 hf, 64 hf + 64): hi pairs
        tc_st32(ca + 32u, reinterpret_cast<const float*>(wlo));  // ... and their lo pairs
        tc_fence_before();
        mbar_arrive(g_ready + 8u * b);
      }

      // ---- LayerNorm2 over accO -> y (TMA stores out of per-warp boxes in hA: all MMAs of the tile have retired)
      mbar_wait(out_full, tpar);
      tc_fence_after();
      if (tid == 0 && ti == 0) trace(7);
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
        const float2 mr = combine_stats(stats, hf, row, shift, s1, s2, P.eps);
        const uint32_t obase = uint32_t(warp - 2) * 16384u;  // 2 x 8 KB: [fp32 box 4 KB][hi 2 KB][lo 2 KB]
        const uint32_t poff = Q.y ? 4096u : 0u;
#pragma unroll 1
        for (int i = 0; i < 4; ++i) {
          const int cl = hf * 128 + i * 32;
          tc_ld32(tmem + kAccO + lane_addr + cl, v);
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 gg = *reinterpret_cast<const float4*>(c_g2 + cl + j);
            const float4 bb = *reinterpret_cast<const float4*>(c_be2 + cl + j);
            v[j] = (v[j] - mr.x) * mr.y * gg.x + bb.x;
            v[j + 1] = (v[j + 1] - mr.x) * mr.y * gg.y + bb.y;
            v[j + 2] = (v[j + 2] - mr.x) * mr.y * gg.z + bb.z;
            v[j + 3] = (v[j + 3] - mr.x) * mr.y * gg.w + bb.w;
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
  if (warp == 1) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");
  }
}

}  // namespace

int debug_set_trace_block(void* dev_buf) {
  long long* p = reinterpret_cast<long long*>(dev_buf);
  SCATT_CUDA(cudaMemcpyToSymbol(g_trace_blk, &p, sizeof(p)));
  return SCATT_OK;
}

bool attn_block_supported(int64_t M, int D, int F) { return D == DM && F >= 128 && F <= kMaxF && F % 128 == 0 && M >= 1 && M < (int64_t(1) << 31); }

int launch_attn_block(const scatt_block_problem* p, int group, int64_t M, int D, int F, float eps, int fmt, int terms, cudaStream_t s) {
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
    SCATT_REQUIRE((reinterpret_cast<uintptr_t>(a.residual_planes) & 15) == 0, "attn_block: residual planes must be 16-byte aligned");
    int rc = encode_planes_map(&P.map_ctx[i], a.ctx_planes, M, DM, BM, fmt);
    if (rc == SCATT_OK) rc = encode_planes_map(&P.map_wo[i], a.wo_planes, DM, DM, 128, fmt);
    if (rc == SCATT_OK) rc = encode_planes_map(&P.map_w1[i], a.w1_planes, F, DM, 128, fmt);
    if (rc == SCATT_OK) rc = encode_planes_map(&P.map_w2[i], a.w2_planes, DM, F, 128, fmt);
    if (rc == SCATT_OK) rc = encode_out_maps(&P.map_y[i], &P.map_p[i], a.y, DM, a.y_planes, M, DM, fmt);
    if (rc != SCATT_OK) return rc;
    P.prob[i] = BlkProblem{a.bo, a.ln1_g, a.ln1_b, a.b1, a.b2, a.ln2_g, a.ln2_b, reinterpret_cast<const uint16_t*>(a.residual_planes),
                           a.y, reinterpret_cast<uint16_t*>(a.y_planes)};
  }
  static std::atomic<bool> attr_done[64];  // per device: the attribute belongs to the device's context
  int dev = 0;
  SCATT_CUDA(cudaGetDevice(&dev));
  if (!attr_done[dev & 63].load()) {
    SCATT_CUDA(cudaFuncSetAttribute(attn_block_kernel<SCATT_PLANE_F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kSmemBytes)));
    SCATT_CUDA(cudaFuncSetAttribute(attn_block_kernel<SCATT_PLANE_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kSmemBytes)));
    attr_done[dev & 63].store(true);
  }
  const int tiles = P.tiles_m * group;
  dim3 grid(unsigned(tiles < 148 ? tiles : 148));
  if (fmt == SCATT_PLANE_F16) (void)launch_kernel(attn_block_kernel<SCATT_PLANE_F16>, grid, dim3(kThreads), kSmemBytes, s, P);
  else (void)launch_kernel(attn_block_kernel<SCATT_PLANE_BF16>, grid, dim3(kThreads), kSmemBytes, s, P);
  return after_launch("attn_block_kernel");
}

}  // namespace scatt
