// tcgen05 engine of scatt_linear (sm_100a):
//   y = epilogue(x W^T + bias),  x and W given as 16-bit hi/lo split planes.
//
// One CTA computes a 128 x BN output tile.  Warp roles (320 threads):
//   warp 0     TMA producer   cp.async.bulk.tensor (3-D maps: K x rows x plane, 128B swizzle)
//   warp 1     TMEM allocator + single-thread tcgen05.mma issuer (accumulators in TMEM)
//   warps 2-9  epilogue: two warps per TMEM lane quadrant, each owning half of the columns
// A `stages`-deep smem ring is handed between producer and issuer with
// full/empty mbarriers; the issuer signals the epilogue through a TMEM-full
// mbarrier (tcgen05.commit).  With `terms` = 3 every K step issues
// hi*hi + lo*hi + hi*lo so products are fp32-grade while running on the fp16 /
// bf16 tensor pipe; `terms` = 1 is the plain 16-bit product.
//
// What keeps the epilogue short (it used to be 4x the main loop, see
// profiles/r01_linear_phase_trace.txt):
//   * accumulator pre-initialisation: while TMA/MMA start up, the (otherwise
//     idle) epilogue warps write `bias + residual` into the TMEM accumulator
//     with tcgen05.st, so the residual stream is fetched off the critical
//     path and the MMAs simply accumulate on top of it;
//   * per-column parameters (bias, gamma, beta) are staged in shared memory once;
//   * every global access goes through a per-warp staging tile, so a warp
//     instruction touches whole row segments (64-byte fp32 / 32-byte plane
//     pieces) instead of 32 different rows;
//   * LayerNorm is fused when the CTA owns the whole row (N == BN == 256): one
//     statistics pass over TMEM (Chan-combined between the two column halves),
//     one normalise-and-store pass.
//
// Schedules (picked per launch in launch_linear_tc from M, N, the group size and the epilogue):
//   linear_tc_ln_cluster_kernel<FMT, CL>   LayerNorm rows of 128 CL = 256 / 512 / 1024 columns over CL-CTA clusters
//                                          (DSMEM statistics exchange), while the clusters fit one wave; the residual
//                                          tile is staged by TMA (in the ring slot that frees first at the end of the
//                                          K loop when three 64 KB stages are in play)
//   linear_tc_kernel<256, 1, FMT>          LayerNorm with one CTA per 128 x 256 row tile (large batches; residual
//                                          stream optionally in split planes only)
//   linear_tc_sub2_kernel<FMT>             one-wave grids of 256-column tiles as two 128-column sub-tiles: the first
//                                          sub-tile's epilogue overlaps the second's MMAs
//   linear_tc_kernel<128 | 64, 0, FMT>     one-wave grids of narrower tiles
//   linear_tc_persist_kernel<FMT, BN, SLIM> multi-wave grids: one CTA per SM walks tiles, two TMEM accumulators
//   linear_tc_dual_kernel<FMT>             (SCATT_PERSIST=0 builds) two CTAs per SM instead
// Compile-time switches for A/B builds (tools/ab.sh): SCATT_RES_STAGED, SCATT_RES_IN_RING, SCATT_SUB2,
// SCATT_PERSIST, SCATT_PERSIST_WIDE (all default to 1).
#include <cuda.h>

#include <cstdlib>
#include <mutex>

#include "common.cuh"
#include "tc_ptx.cuh"
#include "tc_epi.cuh"
#include "tc_host.cuh"

#ifndef SCATT_RES_IN_RING
#define SCATT_RES_IN_RING 1
#endif
#ifndef SCATT_PERSIST
#define SCATT_PERSIST 1
#endif
#ifndef SCATT_PERSIST_WIDE
#define SCATT_PERSIST_WIDE 1
#endif
#ifndef SCATT_SUB2
#define SCATT_SUB2 1
#endif
#ifndef SCATT_RES_STAGED
#define SCATT_RES_STAGED 1
#endif

namespace scatt {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 x 2 B = one 128-byte swizzle row
__host__ __device__ constexpr int num_kb_of(int K) { return (K + BK - 1) / BK; }
constexpr int kEpiWarps = 8;   // default: two epilogue warps per TMEM lane quadrant (EW template parameter)

struct TcProblem {
  const float* bias;
  const float* residual;
  const float* ln_g;
  const float* ln_b;
  float* y;
  uint16_t* y_planes;
  const uint16_t* res_planes;  // residual as [2][M][N] split planes (used when `residual` is null)
};

struct alignas(64) TcParams {
  CUtensorMap map_a[SCATT_MAX_GROUP];
  CUtensorMap map_b[SCATT_MAX_GROUP];
  CUtensorMap map_y[SCATT_MAX_GROUP];   // fp32 output  [M][ldy]      (box 32 x 32, 128B swizzle)
  CUtensorMap map_p[SCATT_MAX_GROUP];   // split planes [2][M][N]     (box 32 x 32 x 1, 64B swizzle)
  CUtensorMap map_r[SCATT_MAX_GROUP];
  CUtensorMap map_rp[SCATT_MAX_GROUP];  // residual kept as split planes [2][M][N] (box 32 x 32 x 1, 64B swizzle)   // fp32 residual [M][ldres]   (box 32 x 32, 128B swizzle; cluster LayerNorm kernels)
  TcProblem prob[SCATT_MAX_GROUP];
  scatt_epilogue ep;
  int64_t M, ldres, ldy;
  int32_t N, K, stages, terms, fmt, fused_ln, pre_init, res_staged, res_in_ring, res_planes, groups;
  int32_t kb0[SCATT_MAX_GROUP];  // split-K launches: first 64-element k-block of problem slot g (K is then the slot's share)
};

// Optional phase trace (dev tool, tools/trace_linear.py): when set, CTA (0,0,0)
// stores clock64() stamps of its pipeline phases.
__device__ long long* g_trace = nullptr;
__device__ __forceinline__ void trace(int slot) {
  if (g_trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) g_trace[slot] = clock64();
}

using namespace tc;

__device__ __forceinline__ void epi_bar_sync() {  // the 8 epilogue warps only
  asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");  // LayerNorm kernels always run EW = 8
}

// ------------------------------------------------------------------ epilogue
// One thread owns one output row (TMEM lane); columns arrive 32 at a time.
// Global I/O goes through a per-warp staging tile of 32 rows x 16 fp32 (rows
// padded to 20 words) so that a warp instruction moves 8 rows x 64 B.
constexpr int kEpiLd = 20;
constexpr int kEpiWarpBytes = 32 * kEpiLd * 4;

struct EpiCtx {
  float* stage;            // this warp's staging tile
  const float* col_bias;   // smem copies of the per-column parameters of this CTA's BN columns
  const float* col_g;
  const float* col_b;
  int64_t row0;            // first global row of this warp's 32-row slab
  int rows_valid;          // rows of the slab that exist (M tail)
  int lane;
  uint32_t out_stage;      // shared address of this warp's 2 x 8 KB output boxes (TMA store sources)
  uint8_t* out_stage_gen;  // same, generic pointer
  const CUtensorMap* map_y;
  const CUtensorMap* map_p;
  int stores;              // output boxes handed to TMA so far (selects the buffer)
  uint32_t nbuf, buf_stride;  // staging buffers of this warp and their distance (2 x 8 KB in the idle operand ring;
                              // two-sub-tile kernels: 8 KB of dedicated staging = 1 x 8 KB or 2 x 4 KB)
  const uint8_t* res_box;  // this warp's residual boxes (32 x 32 fp32, 128B swizzle, one per column chunk) or null
};

// registers <- 32 x 32 fp32 tile of a row-major matrix (two 16-column halves), coalesced
__device__ __forceinline__ void tile_fetch(const EpiCtx& E, const float* __restrict__ g, int64_t ld, int c0, float4 (&r)[8]) {
  const int sub = E.lane >> 2, c4 = (E.lane & 3) * 4;
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int row = 8 * i + sub;
      r[h * 4 + i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row < E.rows_valid) r[h * 4 + i] = *reinterpret_cast<const float4*>(g + (E.row0 + row) * ld + c0 + h * 16 + c4);
    }
}
// v[32] (thread-per-row) += the fetched tile, transposed through the staging tile
__device__ __forceinline__ void tile_add(const EpiCtx& E, const float4 (&r)[8], float* v) {
  const int sub = E.lane >> 2, c4 = (E.lane & 3) * 4;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
#pragma unroll
    for (int i = 0; i < 4; ++i) *reinterpret_cast<float4*>(E.stage + (8 * i + sub) * kEpiLd + c4) = r[h * 4 + i];
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 16; j += 4) {
      const float4 t = *reinterpret_cast<const float4*>(E.stage + E.lane * kEpiLd + j);
      v[h * 16 + j] += t.x, v[h * 16 + j + 1] += t.y, v[h * 16 + j + 2] += t.z, v[h * 16 + j + 3] += t.w;
    }
    __syncwarp();
  }
}

// The same for a residual kept as 16-bit split planes [2][M][N]: r[i] <- 8 hi halves, r[4 + i] <- 8 lo halves
// of row 8 i + lane / 4, columns c0 + 8 (lane % 4) ... (raw bits in the float4 registers).
__device__ __forceinline__ void tile_fetch_planes(const EpiCtx& E, const uint16_t* __restrict__ g, int64_t plane_stride, int64_t ld,
                                                  int c0, float4 (&r)[8]) {
  const int sub = E.lane >> 2, q = E.lane & 3;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = 8 * i + sub;
    r[i] = make_float4(0.f, 0.f, 0.f, 0.f), r[4 + i] = r[i];
    if (row < E.rows_valid) {
      const uint16_t* p = g + (E.row0 + row) * ld + c0 + q * 8;
      r[i] = *reinterpret_cast<const float4*>(p);
      r[4 + i] = *reinterpret_cast<const float4*>(p + plane_stride);
    }
  }
}
// v[32] (thread-per-row) += hi + lo of the fetched plane tile, transposed through the staging tile
__device__ __forceinline__ void tile_add_planes(const EpiCtx& E, const float4 (&r)[8], float* v, int fmt) {
  const int sub = E.lane >> 2, q = E.lane & 3;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    if ((q >> 1) == h) {  // this lane's 8 columns belong to the 16-column half being transposed
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 p0 = unpack_pair(r[i].x, r[4 + i].x, fmt), p1 = unpack_pair(r[i].y, r[4 + i].y, fmt);
        const float2 p2 = unpack_pair(r[i].z, r[4 + i].z, fmt), p3 = unpack_pair(r[i].w, r[4 + i].w, fmt);
        float* dst = E.stage + (8 * i + sub) * kEpiLd + (q & 1) * 8;
        *reinterpret_cast<float4*>(dst) = make_float4(p0.x, p0.y, p1.x, p1.y);
        *reinterpret_cast<float4*>(dst + 4) = make_float4(p2.x, p2.y, p3.x, p3.y);
      }
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 16; j += 4) {
      const float4 t = *reinterpret_cast<const float4*>(E.stage + E.lane * kEpiLd + j);
      v[h * 16 + j] += t.x, v[h * 16 + j + 1] += t.y, v[h * 16 + j + 2] += t.z, v[h * 16 + j + 3] += t.w;
    }
    __syncwarp();
  }
}

// v[32] (thread-per-row) += row `lane` of a 32 x 32 fp32 box the TMA engine wrote with the 128-byte swizzle
__device__ __forceinline__ void box_add(const EpiCtx& E, const uint8_t* box, float* v) {
  const int r = E.lane;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 t = *reinterpret_cast<const float4*>(box + r * 128 + ((j ^ (r & 7)) << 4));
    const float2 a = __fadd2_rn(make_float2(v[4 * j], v[4 * j + 1]), make_float2(t.x, t.y));
    const float2 b = __fadd2_rn(make_float2(v[4 * j + 2], v[4 * j + 3]), make_float2(t.z, t.w));
    v[4 * j] = a.x, v[4 * j + 1] = a.y, v[4 * j + 2] = b.x, v[4 * j + 3] = b.y;
  }
}

// The same for a residual staged as split planes: hi box (32 rows x 64 B, 64-byte swizzle) at +0, lo box at +2048
__device__ __forceinline__ void box_add_planes(const EpiCtx& E, const uint8_t* box, float* v, int fmt) {
  const int r = E.lane;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t off = r * 64 + ((j ^ ((r >> 1) & 3)) << 4);
    const float4 h = *reinterpret_cast<const float4*>(box + off), l = *reinterpret_cast<const float4*>(box + 2048 + off);
    const float2 p0 = unpack_pair(h.x, l.x, fmt), p1 = unpack_pair(h.y, l.y, fmt);
    const float2 p2 = unpack_pair(h.z, l.z, fmt), p3 = unpack_pair(h.w, l.w, fmt);
    v[8 * j] += p0.x, v[8 * j + 1] += p0.y, v[8 * j + 2] += p1.x, v[8 * j + 3] += p1.y;
    v[8 * j + 4] += p2.x, v[8 * j + 5] += p2.y, v[8 * j + 6] += p3.x, v[8 * j + 7] += p3.y;
  }
}

__device__ __forceinline__ void act_vec32(float* v, int act) {
  if (act == SCATT_ACT_GELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = gelu_fast(v[j]);
  } else if (act == SCATT_ACT_RELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
  }
}

// act_post (none / ReLU) -> clamp -> outputs.  Each thread writes its row of the 32 x 32 chunk into
// this warp's swizzled output boxes (fp32: 128-byte rows; planes: 64-byte rows) and one lane hands the
// boxes to the TMA engine (cp.async.bulk.tensor store): the SM's LSU store path (~16 B/clk measured)
// is bypassed, rows past M and columns past N are clipped by the tensor map.
template <int FMT>
__device__ __forceinline__ void chunk_store(const TcParams& P, const TcProblem& Q, EpiCtx& E, float* v, int c0) {
  if (P.ep.act_post == SCATT_ACT_RELU) {  // GELU is only ever a pre-activation on this path (checked on the host)
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
  }
  if (P.ep.clamp > 0.f) {
    const float c = P.ep.clamp;
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fminf(fmaxf(v[j], -c), c);
  }
  const uint32_t buf = (E.nbuf == 2 ? uint32_t(E.stores & 1) : 0u) * E.buf_stride;
  if (E.stores >= int(E.nbuf)) {  // the box last written into this buffer must have been read out by the TMA engine
    if (E.lane == 0) {
      if (E.nbuf == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    __syncwarp();
  }
  uint8_t* box = E.out_stage_gen + buf;
  const uint32_t poff = Q.y ? 4096u : 0u;  // the plane boxes follow the fp32 box when both are written
  const int r = E.lane;
  if (Q.y) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      *reinterpret_cast<float4*>(box + r * 128 + ((j ^ (r & 7)) << 4)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  }
  if (Q.y_planes) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint4 hi, lo;
      split8<FMT>(make_float4(v[8 * j], v[8 * j + 1], v[8 * j + 2], v[8 * j + 3]),
                  make_float4(v[8 * j + 4], v[8 * j + 5], v[8 * j + 6], v[8 * j + 7]), hi, lo);
      const uint32_t off = r * 64 + ((j ^ ((r >> 1) & 3)) << 4);
      *reinterpret_cast<uint4*>(box + poff + off) = hi;
      *reinterpret_cast<uint4*>(box + poff + 2048 + off) = lo;
    }
  }
  fence_proxy_async();
  __syncwarp();
  if (E.lane == 0) {
    const int row = int(E.row0);
    if (Q.y) tma_store_2d(E.map_y, E.out_stage + buf, c0, row);
    if (Q.y_planes) {
      tma_store_3d(E.map_p, E.out_stage + buf + poff, c0, row, 0);
      tma_store_3d(E.map_p, E.out_stage + buf + poff + 2048, c0, row, 1);
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }
  ++E.stores;
}

// (acc [+ bias]) * colscale -> act_pre -> [+ residual]; returns true if v was modified
__device__ __forceinline__ bool chunk_pre(const TcParams& P, const TcProblem& Q, const EpiCtx& E, float* v, int cl, int c0,
                                          bool late_res) {
  bool modified = false;
  if (!P.pre_init && Q.bias) {
    add_cols(v, E.col_bias + cl);
    modified = true;
  }
  if (c0 < P.ep.scale_cols) {  // scale_cols is a multiple of 32 (checked on the host)
    const float s = P.ep.scale;
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] *= s;
    modified = true;
  }
  if (P.ep.act_pre != SCATT_ACT_NONE) {
    act_vec32(v, P.ep.act_pre);
    modified = true;
  }
  if (late_res) {
    if (E.res_box != nullptr) {  // staged by TMA (cluster LayerNorm kernels): chunk cl / 32 of this CTA's columns
      if (P.res_planes) box_add_planes(E, E.res_box + (cl & 63) / 32 * 4096, v, P.fmt);
      else box_add(E, E.res_box + (cl & 63) / 32 * 4096, v);
    } else {
      float4 r[8];
      tile_fetch(E, Q.residual, P.ldres, c0, r);
      tile_add(E, r, v);
    }
    modified = true;
  }
  return modified;
}

// Runs on the epilogue warps while TMA / MMA start: accumulator <- bias + residual.
// PL: the residual stream lives in split planes only (two straight-line copies of the routine, picked once).
template <int BN, int EW, bool PL>
__device__ __forceinline__ void acc_pre_init_impl(const TcParams& P, const TcProblem& Q, const EpiCtx& E, uint32_t tmem_acc, int n0,
                                                  int half) {
  constexpr int kMine = BN / 32 / (EW / 4);  // 32-column chunks per warp
  float4 r[2][8];
  const int64_t ps = P.M * int64_t(P.N);
  auto fetch = [&](int c0, float4 (&dst)[8]) {
    if constexpr (PL) tile_fetch_planes(E, Q.res_planes, ps, P.N, c0, dst);
    else tile_fetch(E, Q.residual, P.ldres, c0, dst);
  };
  int c = half * kMine;
  if (n0 + c * 32 < P.N) fetch(n0 + c * 32, r[0]);
#pragma unroll 2
  for (int i = 0; i < kMine; ++i) {
    const int cl = (half * kMine + i) * 32;
    if (n0 + cl >= P.N) break;
    if (i + 1 < kMine && n0 + cl + 32 < P.N) fetch(n0 + cl + 32, r[(i + 1) & 1]);
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = 0.f;
    if (Q.bias) add_cols(v, E.col_bias + cl);
    if constexpr (PL) tile_add_planes(E, r[i & 1], v, P.fmt);
    else tile_add(E, r[i & 1], v);
    tc_st32(tmem_acc + cl, v);
  }
}
// Only the large-batch LayerNorm kernel (one CTA per 128-row tile, LN == 1) carries the plane-residual copy: the
// small-batch kernels are latency-bound and measurably slower with the extra code in them (B=8 step +3 %).
template <int BN, int EW, bool ALLOW_PL>
__device__ __forceinline__ void acc_pre_init(const TcParams& P, const TcProblem& Q, const EpiCtx& E, uint32_t tmem_acc, int n0,
                                             int half) {
  if constexpr (ALLOW_PL) {
    if (Q.residual == nullptr) {
      acc_pre_init_impl<BN, EW, true>(P, Q, E, tmem_acc, n0, half);
      return;
    }
  }
  acc_pre_init_impl<BN, EW, false>(P, Q, E, tmem_acc, n0, half);
}

// LN: 0 = no LayerNorm in this kernel, 1 = the CTA owns the whole row (N == BN), 2 / 4 / 8 = the row is split
// over the LN CTAs of a cluster (N == LN * BN), which exchange per-row partial statistics through DSMEM.
template <int BN, int LN, int FMT, int EW>
__device__ __forceinline__ void epilogue_rows(const TcParams& P, const TcProblem& Q, EpiCtx& E, uint32_t tmem_acc, int n0,
                                              int half, float2* stats, uint32_t xstats_addr, int row_in_tile) {
  const scatt_epilogue& ep = P.ep;
  constexpr int kMine = BN / 32 / (EW / 4);
  static_assert(LN == 0 || EW == 8, "the fused LayerNorm combines exactly two column halves per CTA");
  static_assert(LN < 2 || kMine == 2, "the staged residual boxes are indexed for two chunks per warp");
  float v[32];
  const bool late_res_any = ep.residual_mode != SCATT_RES_NONE && !P.pre_init;

  if constexpr (LN == 0) {
#pragma unroll 1
    for (int i = 0; i < kMine; ++i) {
      const int cl = (half * kMine + i) * 32, c0 = n0 + cl;
      if (c0 >= P.N) break;  // N is a multiple of 32; warp-uniform
      tc_ld32(tmem_acc + cl, v);
      chunk_pre(P, Q, E, v, cl, c0, late_res_any);  // no LayerNorm: residual before == after
      chunk_store<FMT>(P, Q, E, v, c0);
    }
  } else {
    // LayerNorm over the N columns of the row.  Pass 1: statistics of this warp's half of the
    // CTA's columns (shifted sums), combined with the other half (Chan et al.) and, for LN == 2,
    // with the peer CTA's half of the row.
    constexpr float kHalfN = float(BN / 2);
    const bool late_res = late_res_any && ep.residual_mode == SCATT_RES_BEFORE_LN;
    float shift = 0.f;
    float2 s1p = make_float2(0.f, 0.f), s2p = s1p;
#pragma unroll 1
    for (int i = 0; i < kMine; ++i) {
      const int cl = (half * kMine + i) * 32;
      tc_ld32(tmem_acc + cl, v);
      if (chunk_pre(P, Q, E, v, cl, n0 + cl, late_res)) tc_st32(tmem_acc + cl, v);
      if (i == 0) shift = v[0];
      const float2 ns = make_float2(-shift, -shift);
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        const float2 d = __fadd2_rn(make_float2(v[j], v[j + 1]), ns);
        s1p = __fadd2_rn(s1p, d);
        s2p = __ffma2_rn(d, d, s2p);
      }
    }
    const float s1 = s1p.x + s1p.y, s2 = s2p.x + s2p.y;
    const float dm = s1 / kHalfN;
    const float my_mean = shift + dm, my_m2 = fmaxf(s2 - s1 * dm, 0.f);
    stats[half * BM + row_in_tile] = make_float2(my_mean, my_m2);
    epi_bar_sync();
    if (threadIdx.x == 64) trace(6);
    const float2 other = stats[(half ^ 1) * BM + row_in_tile];
    float mean = 0.5f * (my_mean + other.x);
    const float da = my_mean - mean, db = other.x - mean;
    float m2 = my_m2 + other.y + kHalfN * (da * da + db * db);
    if constexpr (LN >= 2) {
      // the row's LN * BN columns live in the LN CTAs of the cluster: every CTA sends its (mean, m2) to every
      // peer (both column halves hold the CTA's statistics; half h serves the peers of parity h)
      const uint32_t rank = cluster_ctarank();
      cluster_wait();  // phase 1 (armed at kernel start): the peer CTAs are running
#pragma unroll
      for (uint32_t p = 0; p < uint32_t(LN); ++p)
        if ((p & 1u) == uint32_t(half) && p != rank) st_peer_f32x2(xstats_addr + 8u * (rank * BM + row_in_tile), p, mean, m2);
      cluster_sync_all();  // phase 2: every thread of every CTA takes part (warps 0 / 1 after their roles)
      float2 part[LN];
#pragma unroll
      for (uint32_t p = 0; p < uint32_t(LN); ++p) {
        part[p] = make_float2(mean, m2);
        if (p != rank)
          asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(part[p].x), "=f"(part[p].y) : "r"(xstats_addr + 8u * (p * BM + row_in_tile)) : "memory");
      }
      float msum = part[0].x, qsum = part[0].y;
#pragma unroll
      for (int p = 1; p < LN; ++p) msum += part[p].x, qsum += part[p].y;
      mean = msum * (1.0f / float(LN));
      float dsum = 0.f;
#pragma unroll
      for (int p = 0; p < LN; ++p) {
        const float d = part[p].x - mean;
        dsum = fmaf(d, d, dsum);
      }
      m2 = qsum + float(BN) * dsum;
    }
    const float var = m2 / float(LN >= 2 ? LN * BN : BN);
    const float rstd = rsqrtf(var + ep.ln_eps);
    const bool res_after = ep.residual_mode == SCATT_RES_AFTER_LN;
    const bool res_global = res_after && E.res_box == nullptr;  // cluster kernels find the residual staged in smem
    const float2 nm2 = make_float2(-mean, -mean), rs2 = make_float2(rstd, rstd);  // packed fp32 arithmetic: half the issue slots
    auto normalise = [&](int cl) {
      tc_ld32(tmem_acc + cl, v);
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 g = *reinterpret_cast<const float4*>(E.col_g + cl + j);
        const float4 b = *reinterpret_cast<const float4*>(E.col_b + cl + j);
        const float2 y0 = __ffma2_rn(__fmul2_rn(__fadd2_rn(make_float2(v[j], v[j + 1]), nm2), rs2), make_float2(g.x, g.y), make_float2(b.x, b.y));
        const float2 y1 = __ffma2_rn(__fmul2_rn(__fadd2_rn(make_float2(v[j + 2], v[j + 3]), nm2), rs2), make_float2(g.z, g.w), make_float2(b.z, b.w));
        v[j] = y0.x, v[j + 1] = y0.y, v[j + 2] = y1.x, v[j + 3] = y1.y;
      }
    };
    if constexpr (LN >= 2) {
      // Small-batch kernels run this code once or twice per CTA: it is kept ROLLED (one copy of the normalise /
      // residual / store code instead of two) - instruction fetches were 28-30 % of the stall samples of these
      // kernels (profiles/r02_ncu_step_kernels.txt), every line of an unrolled copy is a cold miss.
#pragma unroll 1
      for (int i = 0; i < kMine; ++i) {
        const int cl = (half * kMine + i) * 32;
        normalise(cl);
        if (res_global) {  // only with SCATT_RES_STAGED=0 builds
          float4 r1[8];
          tile_fetch(E, Q.residual, P.ldres, n0 + cl, r1);
          tile_add(E, r1, v);
        } else if (res_after && P.res_planes) {
          box_add_planes(E, E.res_box + i * 4096, v, P.fmt);
        } else if (res_after) {
          box_add(E, E.res_box + i * 4096, v);
        }
        chunk_store<FMT>(P, Q, E, v, n0 + cl);
      }
    } else {
      float4 r[2][8];
      if (res_global) tile_fetch(E, Q.residual, P.ldres, n0 + half * kMine * 32, r[0]);
#pragma unroll 2
      for (int i = 0; i < kMine; ++i) {
        const int cl = (half * kMine + i) * 32;
        if (res_global && i + 1 < kMine) tile_fetch(E, Q.residual, P.ldres, n0 + cl + 32, r[(i + 1) & 1]);
        normalise(cl);
        if (res_global) tile_add(E, r[i & 1], v);
        else if (res_after && P.res_planes) box_add_planes(E, E.res_box + i * 4096, v, P.fmt);
        else if (res_after) box_add(E, E.res_box + i * 4096, v);
        chunk_store<FMT>(P, Q, E, v, n0 + cl);
      }
    }
  }
}

// NSUB = 2: the CTA owns two adjacent 128 x BN sub-tiles with separate TMEM accumulators and walks them back
// to back (the A tile is fetched twice, from L2 the second time): the epilogue of the first sub-tile - bound by
// the SM's ~27 B/clk store path - runs while the MMAs of the second are issued.  One tile per CTA cannot overlap
// the two phases, and at small batches there is no second CTA on the SM to do it.
template <int BN, int LN, int FMT, int EW, int NSUB = 1>
__device__ __forceinline__ void linear_tc_body(const TcParams& P) {
  static_assert(NSUB == 1 || LN == 0, "sub-tiles are for kernels without a fused LayerNorm");
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [stages][A_hi | A_lo | B_hi | B_lo] tiles, barriers, column parameters, LN partials, staging tiles
  constexpr uint32_t kABytes = BM * 128, kBBytes = BN * 128;
  const bool need_a_lo = P.terms >= 2, need_b_lo = P.terms >= 3;
  const uint32_t kBOff = kABytes * (need_a_lo ? 2 : 1);  // B tiles follow the A plane(s) of a stage
  const uint32_t kStageBytes = kBOff + kBBytes * (need_b_lo ? 2 : 1);
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const int stages = P.stages;
  const uint32_t ring_bytes = NSUB > 1 ? uint32_t(stages) * kStageBytes : max(uint32_t(stages) * kStageBytes, uint32_t(EW) * 16384u);
  constexpr uint32_t kOutStage = NSUB > 1 ? uint32_t(EW) * 8192u : 0u;  // dedicated output staging (the ring stays busy)
  // cluster LayerNorm kernels: the CTA's residual tile (128 x 128 fp32 as 32 x 32 boxes, 64 KB) is staged by TMA
  constexpr uint32_t kResBytes = LN >= 2 ? uint32_t(EW) * 2u * 4096u : 0u;
  const uint32_t res_base = base + ring_bytes;
  const uint32_t ostage_base = res_base + ((P.res_staged && !P.res_in_ring) ? kResBytes : 0u);
  // res_in_ring (three 64 KB stages): the residual tile lands in the ring slot that is freed first at the end of
  // the K loop, the output boxes take the other two slots - three operand stages AND the staged residual fit
  const uint32_t res_slot = uint32_t(num_kb_of(P.K) % 3);
  const uint32_t res_addr = P.res_in_ring ? base + res_slot * kStageBytes : res_base;
  const uint32_t bar_base = ostage_base + kOutStage;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (stages + s); };
  const uint32_t tmem_full_bar = bar_base + 16u * stages;  // NSUB barriers, one per accumulator
  const uint32_t acc_init_bar = tmem_full_bar + 8u * NSUB;
  const uint32_t res_bar = acc_init_bar + 8u;
  const uint32_t tmem_ptr_addr = res_bar + 8u;
  constexpr int CW = BN * NSUB;                                       // columns of this CTA
  const uint32_t col_base = (tmem_ptr_addr + 4u + 15u) & ~15u;       // float[3][CW]
  const uint32_t stats_base = col_base + 3u * CW * 4u;               // float2[2][BM]
  const uint32_t xstats_base = stats_base + 2u * BM * 8u;            // float2[LN][BM], slot p written by peer CTA p (LN >= 2)
  const uint32_t stage_base = xstats_base + BM * 8u * uint32_t(LN >= 2 ? LN : 1);  // kEpiWarps staging tiles
  auto gen = [&](uint32_t a) { return smem_raw + (a - raw); };
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(gen(tmem_ptr_addr));

  const int warp = scatt_warp_idx(), lane = threadIdx.x & 31;
  const int g = blockIdx.z;
  const int n0 = blockIdx.x * CW;
  const int64_t m0 = int64_t(blockIdx.y) * BM;
  const int num_kb = (P.K + BK - 1) / BK;
  const TcProblem& Q = P.prob[g];

  if (threadIdx.x == 0) trace(0);
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int sub = 0; sub < NSUB; ++sub) mbar_init(tmem_full_bar + 8u * sub, 1);
    mbar_init(acc_init_bar, 32 * EW);
    mbar_init(res_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&P.map_a[g]) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&P.map_b[g]) : "memory");
    if (Q.y) asm volatile("prefetch.tensormap [%0];" ::"l"(&P.map_y[g]) : "memory");
    if (Q.y_planes) asm volatile("prefetch.tensormap [%0];" ::"l"(&P.map_p[g]) : "memory");
    if (LN >= 2 && P.res_staged) asm volatile("prefetch.tensormap [%0];" ::"l"(P.res_planes ? &P.map_rp[g] : &P.map_r[g]) : "memory");
  }
  if (warp == 1) {  // TMEM allocation (whole warp, .sync.aligned)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_addr), "r"(uint32_t(CW))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp >= 2) {  // per-column parameters of this CTA's columns -> shared memory
    float* col = reinterpret_cast<float*>(gen(col_base));
    for (int i = threadIdx.x - 64; i < CW; i += 32 * EW) {
      const bool in = n0 + i < P.N;
      col[i] = (in && Q.bias) ? Q.bias[n0 + i] : 0.f;
      col[CW + i] = (in && Q.ln_g) ? Q.ln_g[n0 + i] : 0.f;
      col[2 * CW + i] = (in && Q.ln_b) ? Q.ln_b[n0 + i] : 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // LN >= 2: the peer CTAs must be resident before anybody writes into their smem.  Only arrive here; the
  // matching wait sits right before the remote store, so nobody stalls on the peer's start-up.
  if constexpr (LN >= 2) cluster_arrive();
  // Everything above touched only this CTA's resources and static weights; activations follow stream order.
  pdl_launch_dependents();
  pdl_wait();
  const uint32_t tmem_acc = *tmem_ptr_gen;
  if (threadIdx.x == 0) trace(1);

  // Producer and issuer warps walk their loops warp-uniformly; the single-thread instructions (TMA,
  // tcgen05.mma, tcgen05.commit) are issued by an elected lane so they compile to bare UTMALDG / UTCHMMA.
  if (warp == 0) {  // ---------------- TMA producer
    const uint32_t tx = kStageBytes;
    if constexpr (LN >= 2) {
      if (P.res_staged && !P.res_in_ring) {  // residual boxes in epilogue-warp order: warp (quad, half), chunk i -> rows quad*32, cols (half*2+i)*32
        if (elect_one()) {
          mbar_expect_tx(res_bar, kResBytes);
          for (int b = 0; b < EW * 2; ++b) {
            const int w = b >> 1, i = b & 1, quad = (w + 2) & 3, half = w >> 2;
            if (P.res_planes) {
              tma_load_3d(res_base + uint32_t(b) * 4096u, &P.map_rp[g], res_bar, n0 + (half * 2 + i) * 32, int(m0) + quad * 32, 0);
              tma_load_3d(res_base + uint32_t(b) * 4096u + 2048u, &P.map_rp[g], res_bar, n0 + (half * 2 + i) * 32, int(m0) + quad * 32, 1);
            } else {
              tma_load_2d(res_base + uint32_t(b) * 4096u, &P.map_r[g], res_bar, n0 + (half * 2 + i) * 32, int(m0) + quad * 32);
            }
          }
        }
        __syncwarp();
      }
    }
    for (int it = 0; it < NSUB * num_kb; ++it) {  // sub-tile after sub-tile through one continuous ring
      const int kb = NSUB > 1 ? it % num_kb : it, nb = n0 + (NSUB > 1 ? it / num_kb : 0) * BN;
      const int s = it % stages;
      mbar_wait(empty_bar(s), ((it / stages) & 1) ^ 1);
      const uint32_t st = base + s * kStageBytes;
      if (elect_one()) {
        mbar_expect_tx(full_bar(s), tx);
        const int kc = (P.kb0[g] + kb) * BK;
        tma_load_3d(st, &P.map_a[g], full_bar(s), kc, int(m0), 0);
        if (need_a_lo) tma_load_3d(st + kABytes, &P.map_a[g], full_bar(s), kc, int(m0), 1);
        tma_load_3d(st + kBOff, &P.map_b[g], full_bar(s), kc, nb, 0);
        if (need_b_lo) tma_load_3d(st + kBOff + kBBytes, &P.map_b[g], full_bar(s), kc, nb, 1);
      }
      __syncwarp();
      if (it == 0 && lane == 0) trace(2);
    }
    if constexpr (LN >= 2) {
      if (P.res_staged && P.res_in_ring) {  // the slot the next operand stage would take: free once its MMAs retired
        const int it = num_kb, s = it % stages;
        mbar_wait(empty_bar(s), ((it / stages) & 1) ^ 1);
        if (elect_one()) {
          mbar_expect_tx(res_bar, kResBytes);
          for (int b = 0; b < EW * 2; ++b) {
            const int w = b >> 1, i = b & 1, quad = (w + 2) & 3, half = w >> 2;
            if (P.res_planes) {
              tma_load_3d(res_addr + uint32_t(b) * 4096u, &P.map_rp[g], res_bar, n0 + (half * 2 + i) * 32, int(m0) + quad * 32, 0);
              tma_load_3d(res_addr + uint32_t(b) * 4096u + 2048u, &P.map_rp[g], res_bar, n0 + (half * 2 + i) * 32, int(m0) + quad * 32, 1);
            } else {
              tma_load_2d(res_addr + uint32_t(b) * 4096u, &P.map_r[g], res_bar, n0 + (half * 2 + i) * 32, int(m0) + quad * 32);
            }
          }
        }
        __syncwarp();
      }
    }
    if constexpr (LN >= 2) {
      __syncwarp();
      cluster_wait();      // phase 1
      cluster_sync_all();  // phase 2: statistics exchange point of the epilogue warps
    }
  } else if (warp == 1) {  // ---------------- MMA issuer
    // instruction descriptor: D=f32, A/B = f16|bf16, both K-major, N, M=128
    const uint32_t idesc = (1u << 4) | (uint32_t(FMT) << 7) | (uint32_t(FMT) << 10) | (uint32_t(BN >> 3) << 17) |
                           (uint32_t(BM >> 4) << 24);
    uint32_t accumulate = 0;
    if (P.pre_init) {  // the epilogue warps have put bias + residual into the accumulator
      mbar_wait(acc_init_bar, 0);
      tc_fence_after();
      accumulate = 1;
    }
    for (int it = 0; it < NSUB * num_kb; ++it) {
      const int kb = NSUB > 1 ? it % num_kb : it, sub = NSUB > 1 ? it / num_kb : 0;
      const int s = it % stages;
      mbar_wait(full_bar(s), (it / stages) & 1);
      if (it == 0 && lane == 0) trace(3);
      tc_fence_after();
      const uint32_t st = base + s * kStageBytes;
      const uint64_t a_hi = umma_desc_sw128(st), a_lo = umma_desc_sw128(st + kABytes);
      const uint64_t b_hi = umma_desc_sw128(st + kBOff), b_lo = umma_desc_sw128(st + kBOff + kBBytes);
      const uint32_t acc_addr = tmem_acc + uint32_t(sub * BN);
      if (NSUB > 1 && kb == 0) accumulate = 0;  // a fresh accumulator per sub-tile (no pre-initialisation with NSUB > 1)
      if (elect_one()) {
        uint32_t acc = accumulate;
#pragma unroll
        for (int kk = 0; kk < BK / 16; ++kk) {
          const uint64_t adv = uint64_t(kk * 32 >> 4);  // 16 elements x 2 B along K inside the swizzle row
          if (need_b_lo) {
            tc_mma_f16(acc_addr, a_hi + adv, b_lo + adv, idesc, acc);
            acc = 1;
          }
          if (need_a_lo) {
            tc_mma_f16(acc_addr, a_lo + adv, b_hi + adv, idesc, acc);
            acc = 1;
          }
          tc_mma_f16(acc_addr, a_hi + adv, b_hi + adv, idesc, acc);
          acc = 1;
        }
        tc_commit(empty_bar(s));  // smem slot reusable once these MMAs retire
        if (kb == num_kb - 1) tc_commit(tmem_full_bar + 8u * sub);  // this accumulator is complete
      }
      __syncwarp();
      accumulate = 1;
    }
    if (lane == 0) trace(4);
    if constexpr (LN >= 2) {
      __syncwarp();
      cluster_wait();
      cluster_sync_all();
    }
  } else {  // ---------------- epilogue warps 2..9
    const int quad = warp & 3;          // TMEM lane quadrant this warp may access
    const int half = (warp - 2) >> 2;   // which half of the columns
    const float* col = reinterpret_cast<const float*>(gen(col_base));
    EpiCtx E;
    E.stage = reinterpret_cast<float*>(gen(stage_base)) + (warp - 2) * (kEpiWarpBytes / 4);
    E.col_bias = col, E.col_g = col + CW, E.col_b = col + 2 * CW;
    E.row0 = m0 + quad * 32;
    E.rows_valid = int(min(int64_t(32), max(int64_t(0), P.M - E.row0)));
    E.lane = lane;
    // output boxes live in the (by then idle) operand ring: 2 x 8 KB per epilogue warp; with sub-tiles the ring
    // stays busy: 8 KB of dedicated staging per warp, split in two when only one kind of output is written
    E.out_stage = NSUB > 1 ? ostage_base + uint32_t(warp - 2) * 8192u : base + uint32_t(warp - 2) * 16384u;
    if (LN >= 2 && P.res_in_ring)  // four warps in each of the two slots the residual does not occupy
      E.out_stage = base + ((res_slot + 1u + (uint32_t(warp - 2) >> 2)) % 3u) * kStageBytes + (uint32_t(warp - 2) & 3u) * 16384u;
    E.nbuf = 2, E.buf_stride = 8192u;
    if (NSUB > 1) {
      const bool both = Q.y != nullptr && Q.y_planes != nullptr;
      E.nbuf = both ? 1u : 2u, E.buf_stride = 4096u;
    }
    E.out_stage_gen = gen(E.out_stage);
    E.map_y = &P.map_y[g];
    E.map_p = &P.map_p[g];
    E.stores = 0;
    E.res_box = nullptr;
    if constexpr (LN >= 2) {
      if (P.res_staged) E.res_box = gen(res_addr + uint32_t(warp - 2) * 2u * 4096u);
    }
    const uint32_t my_tmem = tmem_acc + (uint32_t(quad * 32) << 16);
    if (P.pre_init) {
      acc_pre_init<BN, EW, LN == 1>(P, Q, E, my_tmem, n0, half);
      tc_fence_before();
      mbar_arrive(acc_init_bar);
      if (threadIdx.x == 64) trace(9);
    }
    if constexpr (LN >= 2) {
      if (P.res_staged) mbar_wait(res_bar, 0);  // lands before (dedicated region) or about when (ring slot) the accumulator completes
    }
#pragma unroll 1
    for (int sub = 0; sub < NSUB; ++sub) {
      mbar_wait(tmem_full_bar + 8u * sub, 0);
      if (threadIdx.x == 64 && sub == 0) trace(5);
      tc_fence_after();
      if (NSUB > 1) E.col_bias = col + sub * BN;  // this sub-tile's slice of the staged column parameters
      epilogue_rows<BN, LN, FMT, EW>(P, Q, E, my_tmem + uint32_t(sub * BN), n0 + sub * BN, half,
                                     reinterpret_cast<float2*>(gen(stats_base)), xstats_base, quad * 32 + lane);
    }
    // the boxes must have been READ out of shared memory before the CTA exits; the global writes drain behind it
    // (they are part of the grid's memory operations: the next kernel's griddepcontrol.wait / stream order covers them)
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    __syncwarp();
    if (threadIdx.x == 64) trace(7);
  }

  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) trace(8);
  if (warp == 1) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"(uint32_t(CW)) : "memory");
  }
}

template <int BN, int LN, int FMT>
__global__ void __launch_bounds__(64 + 32 * kEpiWarps, 1) linear_tc_kernel(const __grid_constant__ TcParams P) {
  linear_tc_body<BN, LN, FMT, kEpiWarps>(P);
}

// One-wave grids of 256-column tiles run them as two 128-column sub-tiles (NSUB = 2, see linear_tc_body).
template <int FMT>
__global__ void __launch_bounds__(64 + 32 * kEpiWarps, 1) linear_tc_sub2_kernel(const __grid_constant__ TcParams P) {
  linear_tc_body<128, 0, FMT, kEpiWarps, 2>(P);
}

// ------------------------------------------------------------------ persistent kernel (multi-wave grids)
// One CTA per SM walks 128 x 128 output tiles (tile = cta, cta + grid, ...; column tiles fastest, so the CTAs
// of a moment share their A tiles in L2).  Two TMEM accumulators alternate: while the eight epilogue warps drain
// tile i (bias / scaling / GELU / residual / clamp, TMA stores out of dedicated staging), the MMA warp already
// accumulates tile i + 1 from an operand ring that runs ahead across tile boundaries.  tmem_full / tmem_empty
// mbarriers hand the accumulators back and forth.  No LayerNorm here (its kernels own whole rows).
// SLIM (BN = 256): 128 x 256 tiles for launches without a residual that write one kind of output - 4 KB of
// output staging per warp and no transposition tiles leave room for two 96 KB operand stages, i.e. twice the
// MMA work per byte in flight (the 128-wide variant's main loop waits on TMA latency with its two stages).
template <int FMT, int BN, bool SLIM>
__global__ void __launch_bounds__(64 + 32 * kEpiWarps, 1) linear_tc_persist_kernel(const __grid_constant__ TcParams P) {
  constexpr int EW = kEpiWarps;
  constexpr uint32_t kWarpStage = SLIM ? 4096u : 8192u;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr uint32_t kABytes = BM * 128, kBBytes = BN * 128;
  const bool need_a_lo = P.terms >= 2, need_b_lo = P.terms >= 3;
  const uint32_t kBOff = kABytes * (need_a_lo ? 2 : 1);
  const uint32_t kStageBytes = kBOff + kBBytes * (need_b_lo ? 2 : 1);
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const int stages = P.stages;
  const uint32_t ostage_base = base + uint32_t(stages) * kStageBytes;    // EW x 8 (4) KB output staging
  const uint32_t bar_base = ostage_base + uint32_t(EW) * kWarpStage;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (stages + s); };
  const uint32_t tmem_full_bar = bar_base + 16u * stages;   // [2]
  const uint32_t tmem_empty_bar = tmem_full_bar + 16u;      // [2]
  const uint32_t tmem_ptr_addr = tmem_empty_bar + 16u;
  const uint32_t col_base = (tmem_ptr_addr + 4u + 15u) & ~15u;  // float[2][BN] bias of the tile, double-buffered
  const uint32_t stage_base = col_base + 2u * BN * 4u;          // EW transposition tiles (residual fetch)
  auto gen = [&](uint32_t a) { return smem_raw + (a - raw); };
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(gen(tmem_ptr_addr));

  const int warp = scatt_warp_idx(), lane = threadIdx.x & 31;
  const int num_kb = (P.K + BK - 1) / BK;
  const int tiles_n = (P.N + BN - 1) / BN, tiles_m = int((P.M + BM - 1) / BM);
  const int tiles_per_group = tiles_n * tiles_m, total_tiles = tiles_per_group * int(P.groups);

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tmem_full_bar + 8u * a, 1);
      mbar_init(tmem_empty_bar + 8u * a, 32 * EW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int g = 0; g < int(P.groups); ++g) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&P.map_a[g]) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&P.map_b[g]) : "memory");
    }
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_addr), "r"(uint32_t(2 * BN)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_launch_dependents();
  pdl_wait();
  const uint32_t tmem_acc = *tmem_ptr_gen;

  if (warp == 0) {  // ---------------- TMA producer
    int it = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const int g = t / tiles_per_group, r = t % tiles_per_group;
      const int m0 = (r / tiles_n) * BM, n0 = (r % tiles_n) * BN;
      for (int kb = 0; kb < num_kb; ++kb, ++it) {
        const int s = it % stages;
        mbar_wait(empty_bar(s), ((it / stages) & 1) ^ 1);
        const uint32_t st = base + s * kStageBytes;
        if (elect_one()) {
          mbar_expect_tx(full_bar(s), kStageBytes);
          tma_load_3d(st, &P.map_a[g], full_bar(s), kb * BK, m0, 0);
          if (need_a_lo) tma_load_3d(st + kABytes, &P.map_a[g], full_bar(s), kb * BK, m0, 1);
          tma_load_3d(st + kBOff, &P.map_b[g], full_bar(s), kb * BK, n0, 0);
          if (need_b_lo) tma_load_3d(st + kBOff + kBBytes, &P.map_b[g], full_bar(s), kb * BK, n0, 1);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {  // ---------------- MMA issuer
    const uint32_t idesc = (1u << 4) | (uint32_t(FMT) << 7) | (uint32_t(FMT) << 10) | (uint32_t(BN >> 3) << 17) |
                           (uint32_t(BM >> 4) << 24);
    int it = 0, i = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++i) {
      const int a = i & 1;
      mbar_wait(tmem_empty_bar + 8u * a, ((i >> 1) & 1) ^ 1);  // the epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t acc_addr = tmem_acc + uint32_t(a * BN);
      for (int kb = 0; kb < num_kb; ++kb, ++it) {
        const int s = it % stages;
        mbar_wait(full_bar(s), (it / stages) & 1);
        tc_fence_after();
        const uint32_t st = base + s * kStageBytes;
        const uint64_t a_hi = umma_desc_sw128(st), a_lo = umma_desc_sw128(st + kABytes);
        const uint64_t b_hi = umma_desc_sw128(st + kBOff), b_lo = umma_desc_sw128(st + kBOff + kBBytes);
        if (elect_one()) {
          uint32_t acc = kb > 0 ? 1u : 0u;
#pragma unroll
          for (int kk = 0; kk < BK / 16; ++kk) {
            const uint64_t adv = uint64_t(kk * 32 >> 4);
            if (need_b_lo) {
              tc_mma_f16(acc_addr, a_hi + adv, b_lo + adv, idesc, acc);
              acc = 1;
            }
            if (need_a_lo) {
              tc_mma_f16(acc_addr, a_lo + adv, b_hi + adv, idesc, acc);
              acc = 1;
            }
            tc_mma_f16(acc_addr, a_hi + adv, b_hi + adv, idesc, acc);
            acc = 1;
          }
          tc_commit(empty_bar(s));
          if (kb == num_kb - 1) tc_commit(tmem_full_bar + 8u * a);
        }
        __syncwarp();
      }
    }
  } else {  // ---------------- epilogue warps 2..9
    const int quad = warp & 3, half = (warp - 2) >> 2;
    float* col = reinterpret_cast<float*>(gen(col_base));
    EpiCtx E;
    E.stage = SLIM ? nullptr : reinterpret_cast<float*>(gen(stage_base)) + (warp - 2) * (kEpiWarpBytes / 4);
    E.col_g = E.col_b = nullptr;
    E.lane = lane;
    E.out_stage = ostage_base + uint32_t(warp - 2) * kWarpStage;
    E.out_stage_gen = gen(E.out_stage);
    E.stores = 0;
    E.res_box = nullptr;
    int i = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++i) {
      const int g = t / tiles_per_group, r = t % tiles_per_group;
      const int64_t m0 = int64_t(r / tiles_n) * BM;
      const int n0 = (r % tiles_n) * BN, a = i & 1;
      const TcProblem& Q = P.prob[g];
      const bool both = Q.y != nullptr && Q.y_planes != nullptr;
      E.nbuf = (both || SLIM) ? 1u : 2u, E.buf_stride = 4096u;
      E.map_y = &P.map_y[g], E.map_p = &P.map_p[g];
      E.row0 = m0 + quad * 32;
      E.rows_valid = int(min(int64_t(32), max(int64_t(0), P.M - E.row0)));
      // this tile's bias slice -> col[a] (double-buffered: the other half may still be read by slower warps)
      if constexpr (SLIM) {
        E.col_bias = Q.bias + n0;  // read in place (warp-uniform, L1-resident): no room for a staged copy
      } else {
        for (int c = threadIdx.x - 64; c < BN; c += 32 * EW) col[a * BN + c] = (n0 + c < P.N && Q.bias) ? Q.bias[n0 + c] : 0.f;
        epi_bar_sync();
        E.col_bias = col + a * BN;
      }
      mbar_wait(tmem_full_bar + 8u * a, (i >> 1) & 1);
      tc_fence_after();
      const uint32_t my_tmem = tmem_acc + uint32_t(a * BN) + (uint32_t(quad * 32) << 16);
      {  // epilogue_rows<LN = 0>, with the accumulator released right after its last TMEM read
        constexpr int kMine = BN / 32 / (EW / 4);
        const bool late_res = !SLIM && P.ep.residual_mode != SCATT_RES_NONE;
        int nvalid = (P.N - n0 - half * kMine * 32 + 31) / 32;  // chunks of this warp left of N
        nvalid = nvalid < 0 ? 0 : (nvalid > kMine ? kMine : nvalid);
        if (nvalid == 0) {
          tc_fence_before();
          mbar_arrive(tmem_empty_bar + 8u * a);
        }
        float v[32];
#pragma unroll 1
        for (int c = 0; c < nvalid; ++c) {
          const int cl = (half * kMine + c) * 32, c0 = n0 + cl;
          tc_ld32(my_tmem + cl, v);
          if (c == nvalid - 1) {
            tc_fence_before();
            mbar_arrive(tmem_empty_bar + 8u * a);
          }
          chunk_pre(P, Q, E, v, cl, c0, late_res);
          chunk_store<FMT>(P, Q, E, v, c0);
        }
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"(uint32_t(2 * BN)) : "memory");
  }
}

// Multi-wave grids (large batches): a 128 x 128 tile per CTA with one epilogue warp per TMEM quadrant and a
// single-stage operand ring needs < 100 KB of shared memory, 128 TMEM columns and 192 threads, so TWO CTAs are
// resident per SM and the hardware overlaps one CTA's epilogue with the other's TMA / MMA phase - the
// overlap a persistent kernel would get from double-buffered accumulators, without the tile scheduler.
template <int FMT>
__global__ void __launch_bounds__(64 + 32 * 4, 2) linear_tc_dual_kernel(const __grid_constant__ TcParams P) {
  linear_tc_body<128, 0, FMT, 4>(P);
}

// LayerNorm GEMM as CL-CTA clusters (one cluster per 128-row tile, BN = 128 per CTA, N = 128 CL = 256 / 512 /
// 1024) for the small-batch regime: CL times the CTAs of a one-CTA-per-row-tile launch share its memory
// traffic, row statistics exchanged through DSMEM.  The cluster shape is a launch attribute.
template <int FMT, int CL>
__global__ void __launch_bounds__(64 + 32 * kEpiWarps, 1) linear_tc_ln_cluster_kernel(const __grid_constant__ TcParams P) {
  linear_tc_body<128, CL, FMT, kEpiWarps>(P);
}

}  // namespace

// ------------------------------------------------------------------ host side
using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                              const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeFn get_encode() {
  static EncodeFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeFn>(p);
  });
  return fn;
}

int encode_planes_map(CUtensorMap* map, const void* planes, int64_t rows, int K, int box_rows, int fmt, int box_planes) {
  EncodeFn enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return SCATT_ERR_CUDA;
  }
  const cuuint64_t dims[3] = {cuuint64_t(K), cuuint64_t(rows), 2};
  const cuuint64_t strides[2] = {cuuint64_t(K) * 2, cuuint64_t(rows) * cuuint64_t(K) * 2};
  const cuuint32_t box[3] = {BK, cuuint32_t(box_rows), cuuint32_t(box_planes)};
  const cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, fmt == SCATT_PLANE_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3,
                   const_cast<void*>(planes), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld K=%d box_rows=%d)", int(r), (long long)rows, K,
              box_rows);
    return SCATT_ERR_CUDA;
  }
  return SCATT_OK;
}

// output maps: 32-row x 32-column boxes written by one epilogue warp
int encode_out_maps(CUtensorMap* map_y, CUtensorMap* map_p, float* y, int64_t ldy, void* planes, int64_t M, int N, int fmt) {
  EncodeFn enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return SCATT_ERR_CUDA;
  }
  const cuuint32_t estr[3] = {1, 1, 1};
  if (y) {
    const cuuint64_t dims[2] = {cuuint64_t(N), cuuint64_t(M)};
    const cuuint64_t strides[1] = {cuuint64_t(ldy) * 4};
    const cuuint32_t box[2] = {32, 32};
    CUresult r = enc(map_y, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, y, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeTiled(y) failed with CUresult %d", int(r));
      return SCATT_ERR_CUDA;
    }
  }
  if (planes) {
    const cuuint64_t dims[3] = {cuuint64_t(N), cuuint64_t(M), 2};
    const cuuint64_t strides[2] = {cuuint64_t(N) * 2, cuuint64_t(M) * cuuint64_t(N) * 2};
    const cuuint32_t box[3] = {32, 32, 1};
    CUresult r = enc(map_p, fmt == SCATT_PLANE_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3,
                     planes, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                     CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeTiled(planes out) failed with CUresult %d", int(r));
      return SCATT_ERR_CUDA;
    }
  }
  return SCATT_OK;
}

namespace {

template <int BN, int LN, int FMT>
int launch_bn_fmt(TcParams& P, int group, cudaStream_t s) {
  const uint32_t kStageBytes = BM * 128 * (P.terms >= 2 ? 2 : 1) + BN * 128 * (P.terms >= 3 ? 2 : 1);
  const int num_kb = (P.K + BK - 1) / BK;
  // TMA-staged residual tile: inside the ring when three 64 KB stages are in play, else in its own 64 KB
  P.res_in_ring = (LN >= 2 && P.res_staged && SCATT_RES_IN_RING && kStageBytes == 65536u && num_kb >= 3) ? 1 : 0;
  const size_t res_bytes = (LN >= 2 && P.res_staged && !P.res_in_ring) ? size_t(kEpiWarps) * 2 * 4096 : 0;
  int stages = int((197u * 1024u - res_bytes) / kStageBytes);
  if (stages > num_kb) stages = num_kb;
  if (stages > 8) stages = 8;
  if (stages < 1) stages = 1;
  P.stages = stages;
  size_t ring = size_t(stages) * kStageBytes;
  if (ring < size_t(kEpiWarps) * 16384) ring = size_t(kEpiWarps) * 16384;  // the epilogue's output boxes reuse the ring
  const size_t smem = ring + res_bytes + 1024 /*align slack*/ + 16 * stages + 64 + 3 * BN * 4 + (2 + (LN >= 2 ? LN : 1)) * BM * 8 + kEpiWarps * kEpiWarpBytes;
  static PerDeviceFlag attr_done;
  dim3 grid((P.N + BN - 1) / BN, unsigned((P.M + BM - 1) / BM), group);
  if constexpr (LN >= 2) {
    if (!attr_done.load()) {
      SCATT_CUDA(cudaFuncSetAttribute(linear_tc_ln_cluster_kernel<FMT, LN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      attr_done.store(true);
    }
    (void)launch_kernel_cluster(linear_tc_ln_cluster_kernel<FMT, LN>, grid, dim3(64 + 32 * kEpiWarps), smem, s, LN, P);
    const int rc = after_launch("linear_tc_ln_cluster_kernel");
    set_last_kernel("linear_tc_ln_cluster_kernel<%d, %d>", FMT, LN);
    return rc;
  } else {
    if (!attr_done.load()) {
      SCATT_CUDA(cudaFuncSetAttribute(linear_tc_kernel<BN, LN, FMT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      attr_done.store(true);
    }
    (void)launch_kernel(linear_tc_kernel<BN, LN, FMT>, grid, dim3(64 + 32 * kEpiWarps), smem, s, P);
    const int rc = after_launch("linear_tc_kernel");
    set_last_kernel("linear_tc_kernel<%d, %d, %d>", BN, LN, FMT);
    return rc;
  }
}

template <int FMT>
int launch_sub2_fmt(TcParams& P, int group, cudaStream_t s) {
  constexpr int BN = 128, NSUB = 2;
  const uint32_t kStageBytes = BM * 128 * (P.terms >= 2 ? 2 : 1) + BN * 128 * (P.terms >= 3 ? 2 : 1);
  const size_t ostage = size_t(kEpiWarps) * 8192;
  const int num_kb = (P.K + BK - 1) / BK;
  int stages = int((200u * 1024u - ostage) / kStageBytes);
  if (stages > NSUB * num_kb) stages = NSUB * num_kb;
  if (stages > 8) stages = 8;
  if (stages < 1) stages = 1;
  P.stages = stages;
  const size_t smem = size_t(stages) * kStageBytes + ostage + 1024 + 16 * stages + 64 + 8 * NSUB + 3 * BN * NSUB * 4 + 3 * BM * 8 +
                      kEpiWarps * kEpiWarpBytes;
  static PerDeviceFlag attr_done;
  if (!attr_done.load()) {
    SCATT_CUDA(cudaFuncSetAttribute(linear_tc_sub2_kernel<FMT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_done.store(true);
  }
  dim3 grid((P.N + BN * NSUB - 1) / (BN * NSUB), unsigned((P.M + BM - 1) / BM), group);
  (void)launch_kernel(linear_tc_sub2_kernel<FMT>, grid, dim3(64 + 32 * kEpiWarps), smem, s, P);
  const int rc = after_launch("linear_tc_sub2_kernel");
  set_last_kernel("linear_tc_sub2_kernel<%d>", FMT);
  return rc;
}

template <int FMT, int BN, bool SLIM>
int launch_persist_fmt(TcParams& P, int group, cudaStream_t s) {
  const uint32_t kStageBytes = BM * 128 * (P.terms >= 2 ? 2 : 1) + BN * 128 * (P.terms >= 3 ? 2 : 1);
  const size_t fixed = size_t(kEpiWarps) * (SLIM ? 4096 : 8192) + 1024 + 256 + (SLIM ? 0 : 2 * BN * 4 + kEpiWarps * kEpiWarpBytes);
  int stages = int((227u * 1024u - fixed - 128) / kStageBytes);
  if (stages > 8) stages = 8;
  if (stages < 1) stages = 1;
  P.stages = stages;
  P.groups = group;
  P.pre_init = 0;
  const size_t smem = size_t(stages) * kStageBytes + fixed + 16 * stages;
  static PerDeviceFlag attr_done;
  if (!attr_done.load()) {
    SCATT_CUDA(cudaFuncSetAttribute(linear_tc_persist_kernel<FMT, BN, SLIM>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_done.store(true);
  }
  const int64_t tiles = int64_t((P.N + BN - 1) / BN) * ((P.M + BM - 1) / BM) * group;
  dim3 grid(unsigned(tiles < 148 ? tiles : 148));
  (void)launch_kernel(linear_tc_persist_kernel<FMT, BN, SLIM>, grid, dim3(64 + 32 * kEpiWarps), smem, s, P);
  const int rc = after_launch("linear_tc_persist_kernel");
  set_last_kernel("linear_tc_persist_kernel<%d, %d, %d>", FMT, BN, SLIM ? 1 : 0);
  return rc;
}

template <int FMT>
int launch_dual_fmt(TcParams& P, int group, cudaStream_t s) {
  constexpr int BN = 128, EW = 4;
  const uint32_t kStageBytes = BM * 128 * (P.terms >= 2 ? 2 : 1) + BN * 128 * (P.terms >= 3 ? 2 : 1);
  int stages = int((64u * 1024u) / kStageBytes);
  const int num_kb = (P.K + BK - 1) / BK;
  if (stages > num_kb) stages = num_kb;
  if (stages < 1) stages = 1;
  P.stages = stages;
  size_t ring = size_t(stages) * kStageBytes;
  if (ring < size_t(EW) * 16384) ring = size_t(EW) * 16384;
  const size_t smem = ring + 1024 + 16 * stages + 64 + 3 * BN * 4 + 3 * BM * 8 + EW * kEpiWarpBytes;
  static PerDeviceFlag attr_done;
  if (!attr_done.load()) {
    SCATT_CUDA(cudaFuncSetAttribute(linear_tc_dual_kernel<FMT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
    attr_done.store(true);
  }
  dim3 grid((P.N + BN - 1) / BN, unsigned((P.M + BM - 1) / BM), group);
  (void)launch_kernel(linear_tc_dual_kernel<FMT>, grid, dim3(64 + 32 * EW), smem, s, P);
  const int rc = after_launch("linear_tc_dual_kernel");
  set_last_kernel("linear_tc_dual_kernel<%d>", FMT);
  return rc;
}

template <int BN, int LN>
int launch_bn(TcParams& P, int group, cudaStream_t s) {
  return P.fmt == SCATT_PLANE_F16 ? launch_bn_fmt<BN, LN, SCATT_PLANE_F16>(P, group, s)
                                  : launch_bn_fmt<BN, LN, SCATT_PLANE_BF16>(P, group, s);
}

}  // namespace

int debug_set_trace(void* dev_buf) {
  long long* p = reinterpret_cast<long long*>(dev_buf);
  SCATT_CUDA(cudaMemcpyToSymbol(g_trace, &p, sizeof(p)));
  return SCATT_OK;
}

// CTAs that share one LayerNorm row in the tcgen05 engine: 1 = one CTA owns the row (N = 256), 2 / 4 / 8 = a
// cluster of 128-column CTAs (N = 256 in the small-batch regime; N = 512 / 1024 while the clusters fit one
// wave of the 148 SMs), 0 = the LayerNorm runs as a separate row-wise launch.
int linear_tc_ln_cluster(int64_t M, int N, int group, int layer_norm) {
  if (!layer_norm) return 0;
  const int64_t row_tiles = ((M + BM - 1) / BM) * group;
  if (N == 256) return row_tiles <= 74 ? 2 : 1;
  if ((N == 512 || N == 1024) && row_tiles * (N / 128) <= 148) return N / 128;
  return 0;
}

// Split-K plan.  A GEMM over few row tiles with a long K loop (the fusion block at small batches: M = B T' = 400 rows,
// K = 1024 / 3072) leaves most SMs idle while each CTA walks 16 - 48 k-blocks.  K is cut into S slices that run as S
// problem slots of ONE plain GEMM launch (same operands, k-block offset per slot, fp32 partial sums into a caller
// workspace); the row-wise reduce kernel adds them in slot order and runs the whole epilogue (bias, activation,
// residual, LayerNorm, planes).  Returns S, or 1 when the launch stays as it is.
int linear_tc_splitk(int64_t M, int N, int K, int group) {
  static const bool off = [] { const char* e = std::getenv("SCATT_SPLITK"); return e && e[0] == '0'; }();
  if (off || group != 1 || N > 1024 || N % 128 != 0 || K < 1024 || K % BK != 0 || M < 1) return 1;
  const int64_t tiles = int64_t(N / 128) * ((M + BM - 1) / BM);
  const int num_kb = K / BK;
  for (int S = SCATT_MAX_GROUP; S >= 3; --S)
    if (num_kb % S == 0 && tiles * S <= 148 && num_kb / S >= 4) return S;
  return 1;
}

size_t linear_tc_workspace_bytes(int64_t M, int N, int K, int group) {
  const int S = linear_tc_splitk(M, N, K, group);
  return S > 1 ? size_t(S) * size_t(M) * size_t(N) * sizeof(float) : 0;
}

static int launch_linear_tc_impl(const scatt_linear_problem* p, int group, int64_t M, int N, int K, int64_t ldres, int64_t ldy,
                                 const scatt_epilogue& ep, int fmt, int terms, cudaStream_t s, int Kmap, const int* kb0);

int launch_linear_tc(const scatt_linear_problem* p, int group, int64_t M, int N, int K, int64_t ldres, int64_t ldy,
                     const scatt_epilogue& ep, int fmt, int terms, void* workspace, size_t workspace_bytes, cudaStream_t s) {
  const int S = linear_tc_splitk(M, N, K, group);
  if (S > 1 && workspace && workspace_bytes >= linear_tc_workspace_bytes(M, N, K, group)) {
    SCATT_REQUIRE(p[0].x_planes && p[0].w_planes && (p[0].y || p[0].y_planes), "linear(tcgen05): problem 0 lacks split planes or an output");
    SCATT_REQUIRE(!ep.layer_norm || (p[0].ln_g && p[0].ln_b), "linear(tcgen05): LayerNorm needs gamma and beta");
    SCATT_REQUIRE(ep.residual_mode == SCATT_RES_NONE || p[0].residual || p[0].residual_planes, "linear(tcgen05): residual missing");
    scatt_linear_problem sp[SCATT_MAX_GROUP] = {};
    int kb0[SCATT_MAX_GROUP] = {};
    for (int i = 0; i < S; ++i) {
      sp[i].x_planes = p[0].x_planes, sp[i].w_planes = p[0].w_planes;
      sp[i].y = reinterpret_cast<float*>(workspace) + size_t(i) * size_t(M) * size_t(N);
      kb0[i] = i * (K / BK / S);
    }
    const scatt_epilogue plain{};
    const int rc = launch_linear_tc_impl(sp, S, M, N, K / S, N, N, plain, fmt, terms, s, K, kb0);
    if (rc != SCATT_OK) return rc;
    char gemm_symbol[128];
    snprintf(gemm_symbol, sizeof(gemm_symbol), "%s", scatt_last_kernel());
    const int rc2 = launch_rowwise_splitk(reinterpret_cast<const float*>(workspace), S, p[0], M, N, ldres, ldy, ep, fmt, s);
    set_last_kernel("%s", gemm_symbol);  // the call's algorithmic flops belong to the GEMM (bench.py keys them by this symbol)
    return rc2;
  }
  return launch_linear_tc_impl(p, group, M, N, K, ldres, ldy, ep, fmt, terms, s, K, nullptr);
}

static int launch_linear_tc_impl(const scatt_linear_problem* p, int group, int64_t M, int N, int K, int64_t ldres, int64_t ldy,
                                 const scatt_epilogue& ep, int fmt, int terms, cudaStream_t s, int Kmap, const int* kb0) {
  SCATT_REQUIRE(terms >= 1 && terms <= 3, "linear(tcgen05): terms must be 1, 2 or 3");
  SCATT_REQUIRE(K % 8 == 0 && N % 32 == 0, "linear(tcgen05): K=%d must be a multiple of 8 and N=%d of 32", K, N);
  SCATT_REQUIRE(ep.scale_cols % 32 == 0, "linear(tcgen05): scale_cols must be a multiple of 32");
  SCATT_REQUIRE(ep.act_post != SCATT_ACT_GELU, "linear(tcgen05): GELU is supported as act_pre only");
  SCATT_REQUIRE(ldres % 4 == 0 && ldy % 4 == 0, "linear(tcgen05): row strides must be multiples of 4");
  SCATT_REQUIRE(M < (int64_t(1) << 31), "linear(tcgen05): M too large");
  if (M == 0) return SCATT_OK;
  // Small-batch regime: when the 128-row tiles of all problems fill at most half of the 148 SMs, halve the
  // tile width of N = 256 outputs (LayerNorm then spans a 2-CTA cluster) so twice as many SMs share the
  // memory traffic of the launch.  Rows of 512 / 1024 columns are normalised in the GEMM by 4- / 8-CTA
  // clusters as long as the clusters fit one wave.
  const int64_t row_tiles = ((M + BM - 1) / BM) * group;
  const int ln_cluster = linear_tc_ln_cluster(M, N, group, ep.layer_norm);  // CTAs sharing a LayerNorm row (0: not fused)
  const bool fused_ln = ln_cluster > 0;
  // LayerNorm wider than the clusters reach: GEMM with the pre-norm part of the chain, then the row-wise tail in place.
  const bool split_ln = ep.layer_norm && !fused_ln;
  // Tile width: the candidate whose grid comes closest to one full wave of the 148 SMs without exceeding
  // it (few row tiles -> narrow tiles -> more CTAs sharing the K loop's memory traffic); the widest tile
  // when even that overflows one wave (large batches: fewer re-reads of A).
  int BN = 256;
  if (fused_ln) {
    BN = ln_cluster >= 2 ? 128 : 256;
  } else {
    int64_t best = -1;
    for (int cand : {256, 128, 64}) {
      const int64_t ctas = ((N + cand - 1) / cand) * row_tiles;
      if (ctas <= 148 && ctas > best) best = ctas, BN = cand;
    }
    if (best < 0) BN = 0;  // more than one wave whatever the width: the two-CTAs-per-SM kernel (128-wide tiles)
  }
  const bool dual = BN == 0;
  if (dual) BN = 128;
  bool one_kind = true;  // every problem writes fp32 or planes, not both
  for (int i = 0; i < group; ++i) one_kind = one_kind && !(p[i].y && p[i].y_planes && !split_ln);
  const bool sub2 = !fused_ln && !dual && BN == 256 && SCATT_SUB2;  // 256-wide tiles as two 128-wide sub-tiles
  if (sub2) BN = 128;
  // multi-wave launches without a residual that write one kind of output: 128 x 256 tiles in the persistent kernel
  const bool persist_wide = dual && SCATT_PERSIST && SCATT_PERSIST_WIDE && one_kind && ep.residual_mode == SCATT_RES_NONE && !split_ln && N >= 256;
  if (persist_wide) BN = 256;

  TcParams P{};
  P.ep = ep;
  if (split_ln) {
    P.ep.layer_norm = 0;
    P.ep.act_post = SCATT_ACT_NONE;
    P.ep.clamp = 0.f;
    if (ep.residual_mode == SCATT_RES_AFTER_LN) P.ep.residual_mode = SCATT_RES_NONE;
  }
  P.M = M, P.N = N, P.K = K, P.ldres = ldres, P.ldy = ldy, P.terms = terms, P.fmt = fmt, P.fused_ln = fused_ln ? 1 : 0;
  // The residual can be folded into the accumulator's initial value when nothing non-linear
  // or scaled sits between the GEMM and the add.
  const bool res_early = P.ep.residual_mode == SCATT_RES_BEFORE_LN || (!fused_ln && P.ep.residual_mode == SCATT_RES_AFTER_LN);
  P.pre_init = (res_early && P.ep.act_pre == SCATT_ACT_NONE && P.ep.scale_cols == 0) ? 1 : 0;
  // Cluster LayerNorm kernels take the residual tile through TMA into shared memory (one 64 KB fetch beside the
  // operand loads) and add it in the epilogue: fetched with ld.global - into the accumulator up front or in the
  // epilogue - it cost 3-5 k cycles per launch either way (profiles/r01_linear_phase_trace_v5.txt).
  if (sub2) P.pre_init = 0;  // the accumulators are started by the MMAs (the residual, if any, is added in the epilogue)
  P.res_staged = (SCATT_RES_STAGED && ln_cluster >= 2 && P.ep.residual_mode != SCATT_RES_NONE) ? 1 : 0;
  if (P.res_staged) P.pre_init = 0;
  for (int i = 0; i < group; ++i) {
    SCATT_REQUIRE(p[i].x_planes && p[i].w_planes, "linear(tcgen05): problem %d lacks split planes", i);
    SCATT_REQUIRE(ep.residual_mode == SCATT_RES_NONE || p[i].residual || p[i].residual_planes, "linear(tcgen05): residual missing");
    SCATT_REQUIRE(ep.residual_mode == SCATT_RES_NONE || p[i].residual || (P.pre_init && ln_cluster == 1) || P.res_staged,
                  "linear(tcgen05): a residual given as split planes is taken by the LayerNorm kernels only (cluster kernels: any "
                  "residual mode; N = 256 beyond 74 row tiles: residual before LayerNorm, no pre-activation, no column scaling)");
    SCATT_REQUIRE(!ep.layer_norm || (p[i].ln_g && p[i].ln_b), "linear(tcgen05): LayerNorm needs gamma and beta");
    SCATT_REQUIRE(!split_ln || p[i].y, "linear(tcgen05): LayerNorm that is not fused (scatt_linear_ln_fused) needs y as scratch");
    SCATT_REQUIRE(p[i].y || p[i].y_planes, "linear(tcgen05): no output");
    P.kb0[i] = kb0 ? kb0[i] : 0;
    int rc = encode_planes_map(&P.map_a[i], p[i].x_planes, M, Kmap, BM, fmt);
    if (rc != SCATT_OK) return rc;
    rc = encode_planes_map(&P.map_b[i], p[i].w_planes, N, Kmap, BN, fmt);
    if (rc != SCATT_OK) return rc;
    P.prob[i] = TcProblem{p[i].bias, p[i].residual, p[i].ln_g, p[i].ln_b, p[i].y,
                          split_ln ? nullptr : reinterpret_cast<uint16_t*>(p[i].y_planes),
                          reinterpret_cast<const uint16_t*>(p[i].residual_planes)};
    rc = encode_out_maps(&P.map_y[i], &P.map_p[i], P.prob[i].y, ldy, P.prob[i].y_planes, M, N, fmt);
    if (rc != SCATT_OK) return rc;
    if (P.res_staged) {  // fp32 residual, or the residual stream kept in split planes only
      const bool as_planes = p[i].residual == nullptr;
      SCATT_REQUIRE(i == 0 || as_planes == (P.res_planes != 0), "linear(tcgen05): the residuals of a group must all be fp32 or all planes");
      P.res_planes = as_planes ? 1 : 0;
      if (as_planes) {
        rc = encode_out_maps(nullptr, &P.map_rp[i], nullptr, 0, const_cast<void*>(p[i].residual_planes), M, N, fmt);
      } else {
        SCATT_REQUIRE((reinterpret_cast<uintptr_t>(p[i].residual) & 15) == 0,
                      "linear(tcgen05): the cluster LayerNorm kernels need a 16-byte aligned fp32 residual");
        rc = encode_out_maps(&P.map_r[i], nullptr, const_cast<float*>(p[i].residual), ldres, nullptr, M, N, fmt);
      }
      if (rc != SCATT_OK) return rc;
    }
  }
  int rc;
  if (fused_ln)
    rc = ln_cluster == 1 ? launch_bn<256, 1>(P, group, s)
       : ln_cluster == 2 ? launch_bn<128, 2>(P, group, s)
       : ln_cluster == 4 ? launch_bn<128, 4>(P, group, s) : launch_bn<128, 8>(P, group, s);
  else if (dual && SCATT_PERSIST && persist_wide)
    rc = fmt == SCATT_PLANE_F16 ? launch_persist_fmt<SCATT_PLANE_F16, 256, true>(P, group, s) : launch_persist_fmt<SCATT_PLANE_BF16, 256, true>(P, group, s);
  else if (dual && SCATT_PERSIST)
    rc = fmt == SCATT_PLANE_F16 ? launch_persist_fmt<SCATT_PLANE_F16, 128, false>(P, group, s) : launch_persist_fmt<SCATT_PLANE_BF16, 128, false>(P, group, s);
  else if (dual) rc = fmt == SCATT_PLANE_F16 ? launch_dual_fmt<SCATT_PLANE_F16>(P, group, s) : launch_dual_fmt<SCATT_PLANE_BF16>(P, group, s);
  else if (sub2) rc = fmt == SCATT_PLANE_F16 ? launch_sub2_fmt<SCATT_PLANE_F16>(P, group, s) : launch_sub2_fmt<SCATT_PLANE_BF16>(P, group, s);
  else rc = BN == 256 ? launch_bn<256, 0>(P, group, s) : (BN == 128 ? launch_bn<128, 0>(P, group, s) : launch_bn<64, 0>(P, group, s));
  if (rc != SCATT_OK || !split_ln) return rc;
  return launch_rowwise_linear_tail(p, group, M, N, ldres, ldy, ep, fmt, s);
}

}  // namespace scatt
