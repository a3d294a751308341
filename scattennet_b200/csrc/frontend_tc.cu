// scatt_frontend on the tensor cores (sm_100a): K1 for large batches, planes-only output
//
//   out[br] = LayerNorm_br(W_br c + b_br + pos_br[t + 2])     c = the stream's joints' x or y coordinates of a frame
//   (region gather model/__init__.py:133-142, x / y split model/keypoint_module.py:23-24, CoordinateMapping
//    model/layers.py:118-123, LearningPositionEmbedding model/layers.py:20-30, first_*_norm keypoint_module.py:155-162)
//
// The CUDA-core kernel (rowwise.cu: frontend_kernel) spends ~660 warp instructions per frame and stream on the
// K_s -> 256 mapping and is bound by instruction issue at 23-26 % of the HBM copy rate (profiles/r02_sweep_membound.md).
// Here the mapping is a [128 frames x 32] x [32 x 256] product on tcgen05 (coordinates and weights as fp16 hi / lo
// split planes, three product terms, fp32 accumulation: ~2^-22 relative, the same class as every other GEMM of the
// path), so the SM only gathers, normalises, splits and stores.  The mapping bias rides in the product too: column
// n_joints of the A tile is 1, row n_joints of W^T is the bias.
//
// One CTA works on ONE (stream, branch) pair - its mapping weight (B operand, 32 KB) and LayerNorm parameters are
// staged once - and walks 128-frame tiles of it.  416 threads:
//   warps 0-7   epilogue, two per TMEM lane quadrant (128 columns each), thread = frame.  Pass 1 adds the frame's
//               position row and accumulates shifted sums (values written back to TMEM); pass 2 normalises, splits
//               into hi / lo and hands 32-column boxes (hi box | lo box, 64-byte swizzle, double-buffered) to the TMA
//               engine.  Packed fp32 arithmetic (FADD2 / FMUL2 / FFMA2) throughout: ~7 instructions per element.  The position rows of a warp's 32 frames are consecutive table rows (t = frame mod T), so
//               they arrive as 32 x 32 fp32 TMA boxes through a two-slot ring per warp that runs two chunks ahead,
//               across tile boundaries; thread-per-row global loads (32 cache lines per request) kept the L1 at
//               60-70 % busy in the first version (profiles/r02_ncu_frontend_tc.txt).  Warps whose 32 frames straddle a
//               sequence boundary take the per-thread loads.
//   warp 8      TMEM allocation + tcgen05.mma issue: <= 6 MMAs (N = 256, K = 16) per tile into one of TWO accumulators,
//               so the MMAs of tile i + 1 run under the epilogue of tile i.
//   warps 9-12  loaders: the region gather straight from keypoints[B,T,K,2] (lane = joint, one request per frame),
//               split into fp16 hi / lo and written as K-major, 32-byte-swizzled A tiles (double-buffered) - a tile
//               ahead of the MMAs.
// Measured (profiles/r02_frontend_tc_ablation.txt, B = 256, T = 200: 334 MB of planes): 133 us = 2.5 TB/s = 38 % of the
// HBM copy rate (the CUDA-core kernel: 221 us).  Without position rows, gather and stores the kernel still takes 90 us:
// each element is read from TMEM twice (statistics, then normalisation) at the ~64 B/clk/SM the TMEM read path gives,
// and ~7 issue slots per element on 8 epilogue warps - that, not HBM, is the bound of this version.
// Algorithmic traffic per frame and stream: 8 n_joints bytes read (sector granularity makes that up to 32 n_joints),
// 2 x 256 x 4 bytes written.
#include <cuda.h>

#include <cstdlib>

#include "common.cuh"
#include "tc_epi.cuh"
#include "tc_host.cuh"
#include "tc_ptx.cuh"

namespace scatt {

namespace {

using namespace tc;

constexpr int kFtEpiWarps = 8;
constexpr int kFtMmaWarp = kFtEpiWarps;
constexpr int kFtLoadWarps = 4;
constexpr int kFtThreads = 32 * (kFtEpiWarps + 1 + kFtLoadWarps);
constexpr int kFtD = 256;
constexpr uint32_t kFtATile = 128 * 32;  // [128 frames x 16 joints] fp16, 32-byte rows
constexpr uint32_t kFtWTile = 256 * 32;  // [256 channels x 16 joints]
// shared memory map (relative to the 1024-aligned base)
constexpr uint32_t kFtWOff = 0;                             // [ks 2][plane 2] W tiles
constexpr uint32_t kFtAOff = kFtWOff + 4 * kFtWTile;        // [stage 2][ks 2][plane 2] A tiles
constexpr uint32_t kFtWarpOff = kFtAOff + 2 * 4 * kFtATile; // per epilogue warp: position ring 2 x 4 KB | output box pair 8 KB
constexpr uint32_t kFtWarpBytes = 16384;
constexpr uint32_t kFtColOff = kFtWarpOff + kFtEpiWarps * kFtWarpBytes;  // float[2][256]: gamma, beta
constexpr uint32_t kFtIdxOff = kFtColOff + 2 * kFtD * 4;    // int[32] joint indices
constexpr uint32_t kFtStatOff = kFtIdxOff + 128;            // float2[parity 2][half 2][128] row statistics of the two column halves
constexpr uint32_t kFtBarOff = kFtStatOff + 2 * 2 * 128 * 8;  // a_full[2], a_empty[2], acc_full[2], acc_empty[2], tmem ptr, pos_full[8][2]
constexpr uint32_t kFtSmemBytes = kFtBarOff + 256 + 1024;   // + alignment slack
static_assert(kFtWarpOff % 1024 == 0, "128-byte-swizzled boxes need 1024-byte alignment");
static_assert(kFtSmemBytes <= 227 * 1024, "frontend_tc: shared memory map exceeds 227 KB");

struct FtPair {  // one (stream, branch)
  const int32_t* joint_idx;
  const float* wt;   // [n_joints][256] transposed mapping weight
  const float* bias; // [256]
  const float* pos;  // [max_pos + 2][256]
  const float* ln_g;
  const float* ln_b;
  int32_t n_joints, coord;
};

struct alignas(64) FtParams {
  CUtensorMap map_out[2 * SCATT_MAX_GROUP];  // planes [2][M][256]: box 32 x 32 x 1, 64-byte swizzle
  CUtensorMap map_pos[2 * SCATT_MAX_GROUP];  // fp32 [max_pos + 2][256]: box 32 x 32, 128-byte swizzle
  FtPair pair[2 * SCATT_MAX_GROUP];
  const float* kp;
  int64_t M;
  int32_t T, K, npairs, ctas_per_pair, tiles_m;
  int32_t ablate;  // dev builds (-DSCATT_FT_ABLATE=1, env SCATT_FT_DBG): 1 = no position rows, 2 = no gather loads, 4 = no stores
};
#ifndef SCATT_FT_ABLATE
#define SCATT_FT_ABLATE 0
#endif

// K-major tile with 32-byte rows (16 K-elements), 32-byte swizzle, 8-row groups 256 B apart
__device__ __forceinline__ uint64_t ft_desc_sw32(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= uint64_t((smem_addr & 0x3FFFFu) >> 4);
  d |= uint64_t(1) << 16;
  d |= uint64_t(256 >> 4) << 32;
  d |= uint64_t(1) << 46;
  d |= uint64_t(6) << 61;
  return d;
}
// byte offset of the 16-byte chunk q (8 K-elements) of row r inside such a tile
__device__ __forceinline__ uint32_t ft_chunk_off(int r, int q) {
  return uint32_t((r >> 3) * 256 + (r & 7) * 32 + ((q ^ ((r >> 2) & 1)) << 4));
}

__device__ __forceinline__ void ft_epi_bar() { asm volatile("bar.sync 1, %0;" ::"n"(32 * kFtEpiWarps) : "memory"); }

// hi / lo split of a pair with packed conversions and a packed subtraction
__device__ __forceinline__ void ft_split2(float2 y, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __float22half2_rn(y);
  const float2 back = __half22float2(h);
  const __half2 l = __float22half2_rn(__fadd2_rn(y, make_float2(-back.x, -back.y)));
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

__global__ void __launch_bounds__(kFtThreads, 1) frontend_tc_kernel(const __grid_constant__ FtParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  const uint32_t bar0 = base + kFtBarOff;
  auto a_full = [&](uint32_t s) { return bar0 + 8u * s; };
  auto a_empty = [&](uint32_t s) { return bar0 + 16u + 8u * s; };
  auto acc_full = [&](uint32_t s) { return bar0 + 32u + 8u * s; };
  auto acc_empty = [&](uint32_t s) { return bar0 + 48u + 8u * s; };
  const uint32_t tmem_ptr_addr = bar0 + 64u;
  auto pos_full = [&](uint32_t w, uint32_t s) { return bar0 + 80u + 16u * w + 8u * s; };

  const int warp = scatt_warp_idx(), lane = threadIdx.x & 31;
  const int pi = int(blockIdx.x) / P.ctas_per_pair, slot = int(blockIdx.x) % P.ctas_per_pair;
  const FtPair& Q = P.pair[pi];
  const int nj = Q.n_joints, nks = nj >= 16 ? 2 : 1;  // + 1 column for the bias
  const int ntiles = slot < P.tiles_m ? (P.tiles_m - slot + P.ctas_per_pair - 1) / P.ctas_per_pair : 0;

  if (threadIdx.x == 0) {
    for (uint32_t s = 0; s < 2; ++s) {
      mbar_init(a_full(s), 32 * kFtLoadWarps);
      mbar_init(a_empty(s), 1);
      mbar_init(acc_full(s), 1);
      mbar_init(acc_empty(s), 32 * kFtEpiWarps);
    }
    for (uint32_t w = 0; w < uint32_t(kFtEpiWarps); ++w) {
      mbar_init(pos_full(w, 0), 1);
      mbar_init(pos_full(w, 1), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&P.map_out[pi]) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&P.map_pos[pi]) : "memory");
  }
  if (warp == kFtMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_addr), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  } else {
    // static data of the pair: W^T [nj][256] fp32 (+ the bias as row nj) -> [ks][plane] K-major tiles (zero beyond),
    // LayerNorm parameters, joint indices
    const int tid = warp < kFtMmaWarp ? int(threadIdx.x) : int(threadIdx.x) - 32;
    constexpr int kStagers = kFtThreads - 32;
    for (int item = tid; item < kFtD * 2 * nks; item += kStagers) {
      const int n = item % kFtD, c = item / kFtD, ks = c >> 1, q = c & 1;
      float w[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int k = ks * 16 + q * 8 + e;
        w[e] = k < nj ? Q.wt[k * kFtD + n] : (k == nj ? Q.bias[n] : 0.f);
      }
      uint4 hi, lo;
      split8<SCATT_PLANE_F16>(make_float4(w[0], w[1], w[2], w[3]), make_float4(w[4], w[5], w[6], w[7]), hi, lo);
      const uint32_t off = kFtWOff + uint32_t(ks) * 2u * kFtWTile + ft_chunk_off(n, q);
      *reinterpret_cast<uint4*>(sm + off) = hi;
      *reinterpret_cast<uint4*>(sm + off + kFtWTile) = lo;
    }
    float* col = reinterpret_cast<float*>(sm + kFtColOff);
    for (int i = tid; i < kFtD; i += kStagers) {
      col[i] = Q.ln_g[i];
      col[kFtD + i] = Q.ln_b[i];
    }
    for (int i = tid; i < int(2 * 4 * kFtATile / 16); i += kStagers) reinterpret_cast<uint4*>(sm + kFtAOff)[i] = make_uint4(0u, 0u, 0u, 0u);
    if (tid < 32) reinterpret_cast<int32_t*>(sm + kFtIdxOff)[tid] = tid < nj ? Q.joint_idx[tid] : 0;
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_launch_dependents();
  pdl_wait();  // everything above is static; keypoints / outputs follow stream order
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(sm + kFtBarOff + 64);

  if (warp > kFtMmaWarp) {  // ================================================= loaders: gather + split -> A tiles
    // lane = joint (lane n_joints carries the constant 1 of the bias column), warp w takes frames 32 w .. 32 w + 31 of the
    // tile: one request per frame spans the stream's joints (168 contiguous bytes for a hand) instead of one request
    // per joint spanning 32 frames (32 cache lines) - the L1 tag stage is shared with the epilogue's LDS / STS traffic.
    // The tiles' columns beyond n_joints were zeroed once; each lane writes only its own 2-byte hi and lo elements.
    const int lw = warp - (kFtMmaWarp + 1);
    const int32_t* idx = reinterpret_cast<const int32_t*>(sm + kFtIdxOff);
    const bool mine = lane < nj;
    const float* kpc = P.kp + Q.coord + (mine ? idx[lane] * 2 : 0);
    const uint32_t koff = uint32_t(lane >> 4) * 2u * kFtATile + uint32_t(lane & 7) * 2u;  // k-step tile + position inside the 16-byte chunk
    const int q = (lane >> 3) & 1;
    for (int it = 0; it < ntiles; ++it) {
      const uint32_t s = uint32_t(it) & 1u;
      const int64_t row0 = int64_t(slot + it * P.ctas_per_pair) * 128 + lw * 32;
      if (it >= 2) mbar_wait(a_empty(s), ((uint32_t(it) >> 1) & 1u) ^ 1u);
#pragma unroll 1
      for (int f0 = 0; f0 < 32; f0 += 8) {
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int64_t row = row0 + f0 + e;
          const bool valid = row < P.M && !(SCATT_FT_ABLATE && (P.ablate & 2));
          v[e] = (valid && mine) ? __ldg(kpc + row * int64_t(P.K) * 2) : ((valid && lane == nj) ? 1.0f : 0.f);
        }
        if (lane <= nj) {
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int r = lw * 32 + f0 + e;
            const __half h = __float2half_rn(v[e]);
            const __half l = __float2half_rn(v[e] - __half2float(h));
            const uint32_t off = kFtAOff + s * 4u * kFtATile + koff + ft_chunk_off(r, q);
            *reinterpret_cast<__half*>(sm + off) = h;
            *reinterpret_cast<__half*>(sm + off + kFtATile) = l;
          }
        }
      }
      fence_proxy_async();  // generic-proxy writes -> visible to the tensor core's operand reads
      mbar_arrive(a_full(s));
    }
  } else if (warp == kFtMmaWarp) {  // ========================================== MMA issuer
    const uint32_t idesc = (1u << 4) | (uint32_t(kFtD >> 3) << 17) | (uint32_t(128 >> 4) << 24);  // f16 x f16 -> f32, M 128, N 256
    for (int it = 0; it < ntiles; ++it) {
      const uint32_t s = uint32_t(it) & 1u, ph = (uint32_t(it) >> 1) & 1u;
      if (it >= 2) mbar_wait(acc_empty(s), ph ^ 1u);  // the epilogue has drained this accumulator
      mbar_wait(a_full(s), ph);
      tc_fence_after();
      const uint32_t d = tmem + s * 256u;
      if (elect_one()) {
        uint32_t acc = 0;
        for (int ks = 0; ks < nks; ++ks) {
          const uint32_t a = base + kFtAOff + s * 4u * kFtATile + uint32_t(ks) * 2u * kFtATile;
          const uint32_t w = base + kFtWOff + uint32_t(ks) * 2u * kFtWTile;
          const uint64_t ah = ft_desc_sw32(a), al = ft_desc_sw32(a + kFtATile);
          const uint64_t wh = ft_desc_sw32(w), wl = ft_desc_sw32(w + kFtWTile);
          tc_mma_f16(d, ah, wl, idesc, acc);
          tc_mma_f16(d, al, wh, idesc, 1);
          tc_mma_f16(d, ah, wh, idesc, 1);
          acc = 1;
        }
        tc_commit(a_empty(s));
        tc_commit(acc_full(s));
      }
      __syncwarp();
    }
  } else {  // ================================================================= epilogue warps 0..7
    const int quad = warp & 3, half = warp >> 2;
    const int r = quad * 32 + lane;
    const uint32_t lane_addr = uint32_t(quad * 32) << 16;
    const float* col = reinterpret_cast<const float*>(sm + kFtColOff);
    float2* stats = reinterpret_cast<float2*>(sm + kFtStatOff);
    uint8_t* wsm = sm + kFtWarpOff + uint32_t(warp) * kFtWarpBytes;          // position ring [2][4 KB]
    const uint32_t wsm_addr = base + kFtWarpOff + uint32_t(warp) * kFtWarpBytes;
    uint8_t* obuf = wsm + 8192;                                              // output box pair
    const uint32_t obuf_addr = wsm_addr + 8192u;
    const bool no_pos = SCATT_FT_ABLATE && (P.ablate & 1);
    // first frame of this warp's 32 in tile `it`, and whether their position rows are 32 consecutive table rows
    auto warp_row0 = [&](int it) { return int64_t(slot + it * P.ctas_per_pair) * 128 + quad * 32; };
    auto t_first = [&](int it) { return int(warp_row0(it) % P.T); };
    auto boxed = [&](int it) { return it < ntiles && t_first(it) + 31 < P.T && !no_pos; };
    uint32_t pos_par = 0u;  // bit sl: parity of the next completion of ring slot sl
    auto issue_pos = [&](int it, int c) {  // chunk c (32 columns) of tile it -> ring slot c & 1
      if (lane == 0) {
        const uint32_t sl = uint32_t(c) & 1u;
        mbar_expect_tx(pos_full(uint32_t(warp), sl), 4096u);
        tma_load_2d(wsm_addr + sl * 4096u, &P.map_pos[pi], pos_full(uint32_t(warp), sl), half * 128 + 32 * c, t_first(it) + 2);
      }
    };
    bool pre = false;  // chunks 0 and 1 of the coming tile have been requested
    uint32_t stores = 0;
    for (int it = 0; it < ntiles; ++it) {
      const uint32_t s = uint32_t(it) & 1u, ph = (uint32_t(it) >> 1) & 1u;
      const int64_t m0 = int64_t(slot + it * P.ctas_per_pair) * 128;
      const int64_t row = m0 + r;
      const bool valid = row < P.M;
      const bool use_box = boxed(it), next_box = boxed(it + 1);
      if (use_box && !pre) {
        issue_pos(it, 0);
        issue_pos(it, 1);
      }
      const float* prow = Q.pos + (int64_t(valid ? row % P.T : 0) + 2) * kFtD + half * 128;
      mbar_wait(acc_full(s), ph);
      tc_fence_after();
      const uint32_t acc = tmem + s * 256u + lane_addr + uint32_t(half * 128);
      float v[32];
      // ---- pass 1: (dot + bias) + position row; shifted sums of the 128 columns
      float shift = 0.f;
      float2 s1 = make_float2(0.f, 0.f), s2 = s1;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        float4 p[8];
        if (use_box) {
          const uint32_t sl = uint32_t(c) & 1u;
          mbar_wait(pos_full(uint32_t(warp), sl), (pos_par >> sl) & 1u);
          pos_par ^= 1u << sl;
          const uint8_t* box = wsm + sl * 4096u;
#pragma unroll
          for (int j = 0; j < 8; ++j) p[j] = *reinterpret_cast<const float4*>(box + lane * 128 + ((j ^ (lane & 7)) << 4));
          __syncwarp();  // every lane has read the slot: it may be refilled
          if (c < 2) issue_pos(it, c + 2);
          else if (next_box) issue_pos(it + 1, c - 2);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            p[j] = (valid && !no_pos) ? __ldg(reinterpret_cast<const float4*>(prow + 32 * c + 4 * j)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        tc_ld32(acc + uint32_t(32 * c), v);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float2 a = __fadd2_rn(make_float2(v[4 * j], v[4 * j + 1]), make_float2(p[j].x, p[j].y));
          const float2 b = __fadd2_rn(make_float2(v[4 * j + 2], v[4 * j + 3]), make_float2(p[j].z, p[j].w));
          v[4 * j] = a.x, v[4 * j + 1] = a.y, v[4 * j + 2] = b.x, v[4 * j + 3] = b.y;
        }
        if (c == 0) shift = v[0];
        const float2 ns = make_float2(-shift, -shift);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float2 d = __fadd2_rn(make_float2(v[2 * j], v[2 * j + 1]), ns);
          s1 = __fadd2_rn(s1, d);
          s2 = __ffma2_rn(d, d, s2);
        }
        tc_st32(acc + uint32_t(32 * c), v);
      }
      pre = use_box ? next_box : false;
      if (!use_box && next_box) {  // (a tile that took the per-thread loads leaves the ring idle)
        issue_pos(it + 1, 0);
        issue_pos(it + 1, 1);
        pre = true;
      }
      const float sum1 = s1.x + s1.y, sum2 = s2.x + s2.y;
      const float dm = sum1 * (1.0f / 128.0f);
      const float my_mean = shift + dm, my_m2 = fmaxf(sum2 - sum1 * dm, 0.f);
      float2* st = stats + (it & 1) * 256;  // alternating buffers: a warp may be a whole tile ahead of its partner's read
      st[half * 128 + r] = make_float2(my_mean, my_m2);
      ft_epi_bar();
      const float2 other = st[(half ^ 1) * 128 + r];
      const float mean = 0.5f * (my_mean + other.x);
      const float da = my_mean - mean, db = other.x - mean;
      const float m2 = my_m2 + other.y + 128.0f * (da * da + db * db);  // Chan et al.
      const float rstd = rsqrtf(m2 * (1.0f / float(kFtD)) + 1e-5f);
      const float2 nm = make_float2(-mean, -mean), rs = make_float2(rstd, rstd);
      // ---- pass 2: normalise, split, 32 columns per pair of TMA stores (hi box | lo box of 32 rows x 64 bytes), two
      // box pairs per warp: the stores of chunk c drain under the arithmetic of chunk c + 1
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        const int cl = half * 128 + 32 * c;
        tc_ld32(acc + uint32_t(32 * c), v);
        if (c == 3) {  // last TMEM read of this tile: the accumulator may be overwritten by tile it + 2
          tc_fence_before();
          mbar_arrive(acc_empty(s));
        }
        uint32_t wh[16], wl[16];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 g = *reinterpret_cast<const float4*>(col + cl + 4 * j);
          const float4 b = *reinterpret_cast<const float4*>(col + kFtD + cl + 4 * j);
          const float2 y0 = __ffma2_rn(__fmul2_rn(__fadd2_rn(make_float2(v[4 * j], v[4 * j + 1]), nm), rs), make_float2(g.x, g.y), make_float2(b.x, b.y));
          const float2 y1 = __ffma2_rn(__fmul2_rn(__fadd2_rn(make_float2(v[4 * j + 2], v[4 * j + 3]), nm), rs), make_float2(g.z, g.w), make_float2(b.z, b.w));
          ft_split2(y0, wh[2 * j], wl[2 * j]);
          ft_split2(y1, wh[2 * j + 1], wl[2 * j + 1]);
        }
        const uint32_t buf = (stores & 1u) * 4096u;
        if (stores >= 2) {  // the box pair last written into this buffer has been read out
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          __syncwarp();
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t off = uint32_t(lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4));
          *reinterpret_cast<uint4*>(obuf + buf + off) = make_uint4(wh[4 * j], wh[4 * j + 1], wh[4 * j + 2], wh[4 * j + 3]);
          *reinterpret_cast<uint4*>(obuf + buf + 2048u + off) = make_uint4(wl[4 * j], wl[4 * j + 1], wl[4 * j + 2], wl[4 * j + 3]);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0 && !(SCATT_FT_ABLATE && (P.ablate & 4))) {
          tma_store_3d(&P.map_out[pi], obuf_addr + buf, cl, int(m0) + quad * 32, 0);  // rows past M are clipped
          tma_store_3d(&P.map_out[pi], obuf_addr + buf + 2048u, cl, int(m0) + quad * 32, 1);
        }
        if (lane == 0) asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        ++stores;
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kFtMmaWarp) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

}  // namespace

// Returns SCATT_OK after launching, or a positive value when the request is outside this kernel's envelope
// (the caller then runs the CUDA-core kernel): fp32 outputs, the exact gathered copy, bf16 planes, tiny inputs.
int launch_frontend_tc(const float* kp, int B, int T, int K, const scatt_frontend_stream* streams, int n, int max_pos, int fmt, bool force,
                       cudaStream_t s) {
  const int64_t M = int64_t(B) * T;
  // below two tiles per CTA the one-off staging of the pair's weights and the serial gather -> MMA -> epilogue chain of a
  // CTA's only tile cost more than the CUDA-core kernel takes (B = 8: 30 us against 25 us)
  const int64_t min_rows = force ? 128 : int64_t(2) * 128 * (148 / (2 * n));
  if (fmt != SCATT_PLANE_F16 || M < min_rows || M > 0x7fffffff - 128) return 1;
  for (int i = 0; i < n; ++i) {
    if (streams[i].n_joints > 31) return 1;  // one A column carries the bias
    if (streams[i].gathered || streams[i].out[0] || streams[i].out[1] || !streams[i].out_planes[0] || !streams[i].out_planes[1]) return 1;
    if ((reinterpret_cast<uintptr_t>(streams[i].pos[0]) | reinterpret_cast<uintptr_t>(streams[i].pos[1])) & 15) return 1;
  }
  FtParams P{};
  P.kp = kp, P.M = M, P.T = T, P.K = K, P.npairs = 2 * n;
  P.tiles_m = int((M + 127) / 128);
  if (SCATT_FT_ABLATE) {
    const char* e = std::getenv("SCATT_FT_DBG");
    P.ablate = e ? std::atoi(e) : 0;
  }
  for (int i = 0; i < n; ++i)
    for (int br = 0; br < 2; ++br) {
      FtPair& q = P.pair[2 * i + br];
      q.joint_idx = streams[i].joint_idx, q.n_joints = streams[i].n_joints, q.coord = streams[i].coord[br];
      q.wt = streams[i].map_wt[br], q.bias = streams[i].map_b[br], q.pos = streams[i].pos[br];
      q.ln_g = streams[i].ln_g[br], q.ln_b = streams[i].ln_b[br];
      const int rc = encode_out_maps(nullptr, &P.map_out[2 * i + br], nullptr, 0, streams[i].out_planes[br], M, kFtD, fmt);
      if (rc != SCATT_OK) return rc;
      const int rc2 = encode_out_maps(&P.map_pos[2 * i + br], nullptr, const_cast<float*>(q.pos), kFtD, nullptr, int64_t(max_pos) + 2, kFtD, fmt);
      if (rc2 != SCATT_OK) return rc2;
    }
  static PerDeviceFlag attr_done;
  if (!attr_done.load()) {
    SCATT_CUDA(cudaFuncSetAttribute(frontend_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kFtSmemBytes)));
    attr_done.store(true);
  }
  const int per = 148 / P.npairs;
  P.ctas_per_pair = P.tiles_m < per ? P.tiles_m : per;
  (void)launch_kernel(frontend_tc_kernel, dim3(P.ctas_per_pair * P.npairs), dim3(kFtThreads), kFtSmemBytes, s, P);
  return after_launch("frontend_tc_kernel");
}

}  // namespace scatt
