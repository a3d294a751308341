// K3 stream attention, TMA-fed variant: q / k / v arrive as the 16-bit hi/lo split
// planes the projection GEMMs wrote ([2][rows][ld], head h at columns col + 16 h),
// so no operand is converted or re-laid-out by this kernel:
//
//   TMA (32-byte-swizzle boxes of 16 columns)  ->  Q [128 x 16], K [keys x 16], V [keys x 16] tiles
//   S = Q K^T   one tcgen05.mma per product term (M=128, N=keys<=224, K=16), K-major operands
//   softmax     two threads per query row (alternate 32-key chunks) out of TMEM: exact two-pass with
//               the reference's mask semantics; chunks with no padded / future key take a
//               branch-free fast path
//   P           written back into the S columns of TMEM (packed 16-bit hi | lo)
//   O = P V     tcgen05.mma, A from TMEM, B = V read MN-major (keys x 16 rows as stored: no transpose)
//
// One CTA = one (batch, head, 128-query tile); 256 TMEM columns (S/P 224 + O 16) -> two CTAs per SM.
// A tile of up to 224 keys computes S once (max pass and exp pass both read it from TMEM), and the PV
// MMAs of a 64-key slice are issued as soon as the softmax warps have written that slice of P, so all
// but the last slice of O = P V overlaps the exponentials.
// Longer key sequences are walked in blocks of 224 keys with an exact two-pass softmax: pass A
// recomputes S block by block for the row maxima, pass B recomputes it again, exponentiates and
// accumulates O = sum_blocks P_blk V_blk (QK^T is one K=16 MMA per term, so recomputing it is free;
// all K / V rows of the sequence stay resident in shared memory).  Up to 7 blocks (1568 keys: 214 KB of
// shared memory, one CTA per SM from the fourth block on); beyond that, and for dense additive masks,
// the fp32 CUDA-core kernel runs.
#include <atomic>
#include <cfloat>
#include <cstdlib>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace scatt {

namespace {

using namespace tc;

constexpr int HD = 16;
constexpr int QT = 128;
constexpr int KBLK = 224;  // keys per block: S/P columns [0, 224), O columns [224, 240)
constexpr int kMaxBlocks = 7;  // 7 x 224 keys x 4 planes of K / V rows = 196 KB: one CTA per SM beyond 3 blocks
constexpr int KMAX = KBLK * kMaxBlocks;
constexpr int kSoftmaxWarps = 8;                     // two per TMEM lane quadrant
constexpr int kThreadsFa = 32 * kSoftmaxWarps + 32;  // + TMA / MMA warp

__device__ long long* g_trace_fa = nullptr;
__device__ __forceinline__ void trace(int slot) {
  if (g_trace_fa != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) g_trace_fa[slot] = clock64();
}

struct FaProblem {
  const uint8_t* key_mask;
  float* out;
  uint16_t* out_planes;
  int32_t q_col, k_col, v_col;
};

struct alignas(64) FaParams {
  CUtensorMap map_q[SCATT_MAX_GROUP];
  CUtensorMap map_k[SCATT_MAX_GROUP];
  CUtensorMap map_v[SCATT_MAX_GROUP];
  FaProblem p[SCATT_MAX_GROUP];
  int32_t B, Tq, Tk, H, kind, terms, kbox, nblk, groups, debug_stage;   // kbox: rows per TMA box / keys per block; nblk: blocks covering Tk
};

// shared memory map (relative to a 1024-aligned base); every tile row is 32 bytes (16 halves).
// K / V regions hold nblk * kbox rows each, so their offsets depend on the launch.
struct FaSmem {
  uint32_t qh, ql, kh, kl, vh, vl, cls, xch, flag, bar, total;
};
__host__ __device__ inline FaSmem fa_smem_map(int nblk, int kbox) {
  FaSmem m;
  const uint32_t kv = (uint32_t(nblk) * uint32_t(kbox) * 32u + 1023u) & ~1023u;
  m.qh = 0, m.ql = QT * 32;
  m.kh = 2 * QT * 32, m.kl = m.kh + kv, m.vh = m.kl + kv, m.vl = m.vh + kv;
  m.cls = m.vl + kv;                              // float[nblk * 224 + 32] key class: 0 valid / -FLT_MAX padded / -inf absent
  m.xch = m.cls + (uint32_t(nblk) * KBLK + 32) * 4;   // float[2][128] row max, float[2][128] row sum (pair exchange)
  m.flag = m.xch + 4 * 128 * 4;                   // uint32[nblk * 7]: chunk has only valid keys
  m.bar = (m.flag + uint32_t(nblk) * 7 * 4 + 15u) & ~15u;  // 6 mbarriers + tmem pointer + 3 more P-slice mbarriers
  m.total = m.bar + 96 + 1024;
  return m;
}

// 32-byte-swizzled tile: 8-row groups of 256 bytes.  K-major use (Q, K): rows = M/N index, 16 K-elements per row.
// MN-major use (V as B operand): rows = K index (keys), 16 N-elements per row.
__device__ __forceinline__ uint64_t umma_desc_sw32(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= uint64_t((smem_addr & 0x3FFFFu) >> 4);
  d |= uint64_t(1) << 16;            // leading byte offset: unused (one swizzle atom wide)
  d |= uint64_t(256 >> 4) << 32;     // stride byte offset: next 8-row group
  d |= uint64_t(1) << 46;            // descriptor version (sm_100)
  d |= uint64_t(6) << 61;            // SWIZZLE_32B
  return d;
}

__device__ __forceinline__ void tc_mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

template <int FMT>
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  if (FMT == SCATT_PLANE_F16) {
    const __half2 h = __floats2half2_rn(a, b);
    const float2 back = __half22float2(h);
    const float2 d = __fadd2_rn(make_float2(a, b), make_float2(-back.x, -back.y));
    const __half2 l = __floats2half2_rn(d.x, d.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
  } else {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    const float2 back = __bfloat1622float2(h);
    const float2 d = __fadd2_rn(make_float2(a, b), make_float2(-back.x, -back.y));
    const __nv_bfloat162 l = __floats2bfloat162_rn(d.x, d.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
  }
}

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int FMT>
__global__ void __launch_bounds__(kThreadsFa, 2) stream_attention_fa_kernel(const __grid_constant__ FaParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  const FaSmem L = fa_smem_map(P.nblk, P.kbox);
  const uint32_t bar_qk = base + L.bar, bar_v = bar_qk + 8, bar_s = bar_qk + 16, bar_p = bar_qk + 24, bar_o = bar_qk + 32;
  const uint32_t bar_sf = bar_qk + 40, tmem_ptr_addr = bar_qk + 48;
  // P is handed to the MMA warp in slices of 64 keys (one chunk per thread of a row pair): slice 0 on bar_p,
  // slices 1..3 on the barriers behind the TMEM pointer
  auto bar_pk = [&](int it) -> uint32_t { return it == 0 ? bar_p : bar_qk + 48 + 8 * it; };
  float* cls = reinterpret_cast<float*>(sm + L.cls);

  const int warp = scatt_warp_idx(), lane = threadIdx.x & 31;
  const int g = blockIdx.z / P.B, b = blockIdx.z % P.B, h = blockIdx.y;
  const FaProblem& A = P.p[g];
  const int m0 = blockIdx.x * QT;
  const int Tq = P.Tq, Tk = P.Tk, D = P.H * HD;
  const bool causal = P.kind == SCATT_ATTN_CAUSAL;
  const int kbox = P.kbox;                               // keys per block (multiple of 16)
  const int nk = causal ? min(Tk, m0 + QT) : Tk;         // keys this tile may see
  const int nblk = (nk + kbox - 1) / kbox;               // key blocks this tile walks (<= P.nblk)
  const int nchunk = (kbox + 31) >> 5;                   // 32-key chunks per block
  const bool lo_q = P.terms >= 2, lo_k = P.terms >= 3;   // S: q_hi k_lo (3), q_lo k_hi (2), q_hi k_hi
  constexpr uint32_t kTmemCols = 256, kOCol = 224;

  constexpr int kMmaWarp = kSoftmaxWarps;
  if (threadIdx.x == 0) trace(0);
  if (threadIdx.x == 32 * kMmaWarp) {
    mbar_init(bar_qk, 1);
    mbar_init(bar_v, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_p, 32 * kSoftmaxWarps);
    mbar_init(bar_o, 1);
    mbar_init(bar_sf, 32 * kSoftmaxWarps);
    for (int it = 1; it < 4; ++it) mbar_init(bar_pk(it), 32 * kSoftmaxWarps);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&P.map_q[g]) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&P.map_k[g]) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&P.map_v[g]) : "memory");
  }
  if (warp == kMmaWarp) {
    // operands first (they only need the barriers): Q box + nblk K boxes, nblk V boxes (x2 planes) are in
    // flight while TMEM is allocated and the softmax warps classify the keys
    __syncwarp();
    pdl_launch_dependents();
    pdl_wait();  // q / k / v planes come from the preceding GEMM
    if (elect_one()) {
      const int qrow = b * Tq + m0, krow = b * Tk;
      const int nq = lo_q ? 2 : 1, nkpl = lo_k ? 2 : 1;
      mbar_expect_tx(bar_qk, uint32_t(QT * 32 * nq + nblk * kbox * 32 * nkpl));
      tma_load_3d(base + L.qh, &P.map_q[g], bar_qk, A.q_col + h * HD, qrow, 0);
      if (lo_q) tma_load_3d(base + L.ql, &P.map_q[g], bar_qk, A.q_col + h * HD, qrow, 1);
      for (int blk = 0; blk < nblk; ++blk) {
        tma_load_3d(base + L.kh + blk * kbox * 32, &P.map_k[g], bar_qk, A.k_col + h * HD, krow + blk * kbox, 0);
        if (lo_k) tma_load_3d(base + L.kl + blk * kbox * 32, &P.map_k[g], bar_qk, A.k_col + h * HD, krow + blk * kbox, 1);
      }
      mbar_expect_tx(bar_v, uint32_t(nblk * kbox * 32 * nkpl));
      for (int blk = 0; blk < nblk; ++blk) {
        tma_load_3d(base + L.vh + blk * kbox * 32, &P.map_v[g], bar_v, A.v_col + h * HD, krow + blk * kbox, 0);
        if (lo_k) tma_load_3d(base + L.vl + blk * kbox * 32, &P.map_v[g], bar_v, A.v_col + h * HD, krow + blk * kbox, 1);
      }
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_addr), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  } else {
    // key classes; a warp also publishes whether its 32-key chunk holds valid keys only
    // (keys of block blk live at class / flag index blk * nchunk * 32 + ..., blocks padded to whole chunks)
    uint32_t* flags = reinterpret_cast<uint32_t*>(sm + L.flag);
    for (int cc = warp; cc < nblk * nchunk; cc += kSoftmaxWarps) {
      const int blk = cc / nchunk, j = blk * kbox + (cc % nchunk) * 32 + lane;  // key index
      float c = -INFINITY;  // absent key (past the block, past nk): probability exactly 0
      if ((cc % nchunk) * 32 + lane < kbox && j < nk) c = (A.key_mask && A.key_mask[int64_t(b) * Tk + j] == 0) ? -FLT_MAX : 0.f;
      cls[cc * 32 + lane] = c;
      const bool all_valid = __all_sync(0xffffffffu, c == 0.f);
      if (lane == 0) flags[cc] = all_valid ? 1u : 0u;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp != kMmaWarp) {
    pdl_launch_dependents();
    pdl_wait();
  }
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(sm + L.bar + 48);  // written by tcgen05.alloc
  const uint32_t tmem_s = tmem, tmem_o = tmem + kOCol;
  if (threadIdx.x == 0) trace(1);
  const uint32_t idesc_base = (1u << 4) | (uint32_t(FMT) << 7) | (uint32_t(FMT) << 10) | (uint32_t(QT >> 4) << 24);

  if (warp == kMmaWarp) {
    // The whole warp walks this role uniformly; single-thread instructions are issued by an elected lane.
    if (lane == 0) trace(2);
    mbar_wait(bar_qk, 0);
    if (lane == 0) trace(3);
    tc_fence_after();
    const uint32_t idesc_s = idesc_base | (uint32_t(kbox >> 3) << 17);
    const uint32_t idesc_o = idesc_base | (1u << 16) | (uint32_t(HD >> 3) << 17);  // bit 16: B is MN-major
    const uint64_t qh = umma_desc_sw32(base + L.qh), ql = umma_desc_sw32(base + L.ql);
    const uint64_t kh0 = umma_desc_sw32(base + L.kh), kl0 = umma_desc_sw32(base + L.kl);
    const uint64_t vh0 = umma_desc_sw32(base + L.vh), vl0 = umma_desc_sw32(base + L.vl);
    auto issue_s = [&](int blk) {  // S = Q K_blk^T, one K = 16 MMA per product term
      const uint64_t off = uint64_t(blk * kbox * 32 >> 4);
      if (elect_one()) {
        uint32_t acc = 0;
        if (lo_k) {
          tc_mma_f16(tmem_s, qh, kl0 + off, idesc_s, acc);
          acc = 1;
        }
        if (lo_q) {
          tc_mma_f16(tmem_s, ql, kh0 + off, idesc_s, acc);
          acc = 1;
        }
        tc_mma_f16(tmem_s, qh, kh0 + off, idesc_s, acc);
        tc_commit(bar_s);
      }
      __syncwarp();
    };
    // ---- pass A: row maxima (the softmax warps hand the S buffer back through bar_sf)
    for (int blk = 0; blk < nblk; ++blk) {
      if (blk > 0) {
        mbar_wait(bar_sf, (blk - 1) & 1);
        tc_fence_after();
      }
      issue_s(blk);
    }
    mbar_wait(bar_sf, (nblk - 1) & 1);
    tc_fence_after();
    // ---- pass B: S again, P = exp(S - max) written over it by the softmax warps, O += P V_blk
    // (tcgen05.mma executes in issue order, so S of the next block cannot overtake the PV MMAs reading P)
    mbar_wait(bar_v, 0);
    uint32_t acc_o = 0;
    for (int blk = 0; blk < nblk; ++blk) {
      if (nblk > 1) issue_s(blk);  // a single block still holds the S of pass A
      const int nkeys = min(kbox, nk - blk * kbox);  // causal tiles stop at the diagonal
      const int ksteps = (nkeys + 15) >> 4, niter = (((nkeys + 31) >> 5) + 1) >> 1;
      const uint64_t vh = vh0 + uint64_t(blk * kbox * 32 >> 4), vl = vl0 + uint64_t(blk * kbox * 32 >> 4);
      for (int it = 0; it < niter; ++it) {
        mbar_wait(bar_pk(it), blk & 1);  // P of keys [64 it, 64 it + 64) is in TMEM
        tc_fence_after();
        if (elect_one()) {
          const int ks1 = min(4 * it + 4, ksteps);
          for (int ks = 4 * it; ks < ks1; ++ks) {
            // P of keys [16 ks, 16 ks + 16): hi in 8 columns of the 32-key chunk, lo 16 columns further
            const uint32_t p_hi = tmem_s + uint32_t(ks >> 1) * 32 + uint32_t(ks & 1) * 8;
            const uint64_t voff = uint64_t(ks) * 32;  // 16 keys x 32 bytes
            if (lo_k) {
              tc_mma_ts(tmem_o, p_hi, vl + voff, idesc_o, acc_o);
              acc_o = 1;
            }
            if (lo_q) {
              tc_mma_ts(tmem_o, p_hi + 16, vh + voff, idesc_o, acc_o);
              acc_o = 1;
            }
            tc_mma_ts(tmem_o, p_hi, vh + voff, idesc_o, acc_o);
            acc_o = 1;
          }
        }
        __syncwarp();
        acc_o = 1;
      }
    }
    if (elect_one()) tc_commit(bar_o);
    __syncwarp();
  } else {
    // ---------------- softmax: two threads per query row; row = (warp % 4) * 32 + lane (TMEM lane),
    // warp / 4 picks the even or the odd 32-key chunks
    const int quad = warp & 3, half = warp >> 2;
    const int r = quad * 32 + lane;
    const int i = m0 + r;
    const uint32_t lane_addr = uint32_t(quad * 32) << 16;
    const int jmax = causal ? i : 0x7fffffff;
    // a warp whose 32 rows all lie past Tq (the second query tile of T = 200 holds 72 rows: 1.75 of its 4 quadrants are
    // dead) only takes part in the hand-shakes: its issue slots and MUFU cycles go to the SM's other resident CTA
    const bool live = m0 + quad * 32 < Tq;
    const uint32_t* chunk_valid = reinterpret_cast<const uint32_t*>(sm + L.flag);
    float* xch = reinterpret_cast<float*>(sm + L.xch);
    uint32_t s_uses = 0;  // completed phases of bar_s
    const float kLog2e = 1.4426950408889634f;
    float v[32];
    float mx = -INFINITY;
#pragma unroll 1
    for (int blk = 0; blk < nblk; ++blk) {
      mbar_wait(bar_s, s_uses++ & 1);
      if (threadIdx.x == 0 && blk == 0) trace(4);
      tc_fence_after();
      int nch = live ? (min(kbox, nk - blk * kbox) + 31) >> 5 : 0;  // chunks holding a key this tile may see
      // causal: a chunk that starts past this warp's last query row holds future keys only (probability 0 for every row)
      if (causal) nch = min(nch, max(0, (m0 + quad * 32 + 31 - blk * kbox) >> 5) + 1);
#pragma unroll 1
      for (int c = half; c < nch; c += 2) {
        tc_ld32(tmem_s + lane_addr + c * 32, v);
        const int cc = blk * nchunk + c, key0 = blk * kbox + c * 32;
        // fast path: every key of the chunk is valid and visible to every row of this warp
        const bool fast = chunk_valid[cc] != 0u && (!causal || key0 + 31 <= m0 + quad * 32);
        if (fast) {
#pragma unroll
          for (int j = 0; j < 32; ++j) mx = fmaxf(mx, v[j]);
        } else {
          // 8-key groups past the last key of the tile (T = 200: 24 of the 32 keys of the last chunk) are skipped
          const int live = min(kbox, nk - blk * kbox) - c * 32;  // keys of this chunk the tile may see (warp-uniform)
#pragma unroll
          for (int g8 = 0; g8 < 4; ++g8) {
            if (8 * g8 < live) {
#pragma unroll
              for (int j = 8 * g8; j < 8 * g8 + 8; ++j) {
                const float kc = cls[cc * 32 + j];
                float s = kc == 0.f ? v[j] : kc;
                if (key0 + j > jmax) s = -INFINITY;
                mx = fmaxf(mx, s);
              }
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(bar_sf);  // this thread is done reading the block's S
    }
    if (threadIdx.x == 0) trace(5);
    xch[half * 128 + r] = mx;
    asm volatile("bar.sync 1, %0;" ::"n"(32 * kSoftmaxWarps) : "memory");
    mx = fmaxf(mx, xch[(half ^ 1) * 128 + r]);  // every row sees key 0, so mx is finite
    const float mneg = -mx * kLog2e;
    float l = 0.f;
    float2 la = make_float2(0.f, 0.f), lb = la;
#pragma unroll 1
    for (int blk = 0; blk < nblk; ++blk) {
      if (nblk > 1) {  // S of this block was recomputed; a single block is still there from pass A
        mbar_wait(bar_s, s_uses++ & 1);
        tc_fence_after();
      }
      const int nch = (min(kbox, nk - blk * kbox) + 31) >> 5;
      const int niter = (nch + 1) >> 1;
#pragma unroll 1
      for (int it = 0; it < niter; ++it) {
        const int c = 2 * it + half;
        if (c < nch && live && causal && blk * kbox + c * 32 > m0 + quad * 32 + 31) {
          float z[32];  // future keys only for this warp's rows: P = 0 without reading S
#pragma unroll
          for (int j = 0; j < 32; ++j) z[j] = 0.f;
          tc_st32(tmem_s + lane_addr + c * 32, z);
        } else if (c < nch && live) {
          tc_ld32(tmem_s + lane_addr + c * 32, v);
          const int cc = blk * nchunk + c, key0 = blk * kbox + c * 32;
          const bool fast = chunk_valid[cc] != 0u && (!causal || key0 + 31 <= m0 + quad * 32);
          if (fast) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {  // packed row sums, two independent chains
              const float p0 = ex2(fmaf(v[j], kLog2e, mneg)), p1 = ex2(fmaf(v[j + 1], kLog2e, mneg));
              const float p2 = ex2(fmaf(v[j + 2], kLog2e, mneg)), p3 = ex2(fmaf(v[j + 3], kLog2e, mneg));
              la = __fadd2_rn(la, make_float2(p0, p1));
              lb = __fadd2_rn(lb, make_float2(p2, p3));
              v[j] = p0, v[j + 1] = p1, v[j + 2] = p2, v[j + 3] = p3;
            }
          } else {
            const int live = min(kbox, nk - blk * kbox) - c * 32;  // as in the max pass: absent keys have probability 0
#pragma unroll
            for (int g8 = 0; g8 < 4; ++g8) {
              if (8 * g8 < live) {
#pragma unroll
                for (int j = 8 * g8; j < 8 * g8 + 8; ++j) {
                  const float kc = cls[cc * 32 + j];
                  float s = kc == 0.f ? v[j] : kc;
                  if (key0 + j > jmax) s = -INFINITY;
                  const float p = ex2((s - mx) * kLog2e);  // subtract first: s and mx may both be -FLT_MAX
                  l += p;
                  v[j] = p;
                }
              } else {
#pragma unroll
                for (int j = 8 * g8; j < 8 * g8 + 8; ++j) v[j] = 0.f;
              }
            }
          }
          float w[32];
          uint32_t* wp = reinterpret_cast<uint32_t*>(w);
#pragma unroll
          for (int q = 0; q < 16; ++q) split2<FMT>(v[2 * q], v[2 * q + 1], wp[q], wp[16 + q]);
          tc_st32(tmem_s + lane_addr + c * 32, w);
        }
        tc_fence_before();
        mbar_arrive(bar_pk(it));  // this 64-key slice of P may be multiplied
      }
    }
    if (threadIdx.x == 0) trace(6);
    l += (la.x + la.y) + (lb.x + lb.y);
    xch[256 + half * 128 + r] = l;
    asm volatile("bar.sync 1, %0;" ::"n"(32 * kSoftmaxWarps) : "memory");

    {  // each thread of the pair scales and stores 8 of the row's 16 columns
      l += xch[256 + (half ^ 1) * 128 + r];
      mbar_wait(bar_o, 0);
      if (threadIdx.x == 0) trace(7);
      tc_fence_after();
      float o[8];
      tc_ld8(tmem_o + lane_addr + 8 * half, o);
      const float inv = 1.0f / l;
      if (i < Tq) {
        const int64_t off = (int64_t(b) * Tq + i) * D + h * HD + 8 * half;
#pragma unroll
        for (int c = 0; c < 8; ++c) o[c] *= inv;
        if (A.out) {
          *reinterpret_cast<float4*>(A.out + off) = make_float4(o[0], o[1], o[2], o[3]);
          *reinterpret_cast<float4*>(A.out + off + 4) = make_float4(o[4], o[5], o[6], o[7]);
        }
        if (A.out_planes) {
          uint4 ph, pl;
          split2<FMT>(o[0], o[1], ph.x, pl.x);
          split2<FMT>(o[2], o[3], ph.y, pl.y);
          split2<FMT>(o[4], o[5], ph.z, pl.z);
          split2<FMT>(o[6], o[7], ph.w, pl.w);
          *reinterpret_cast<uint4*>(A.out_planes + off) = ph;
          *reinterpret_cast<uint4*>(A.out_planes + int64_t(P.B) * Tq * D + off) = pl;
        }
      }
    }
  }

  if (threadIdx.x == 0) trace(8);
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) trace(9);
  if (warp == kMmaWarp) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");
  }
}

#ifdef SCATT_FA2_DEBUG
#include <cstdio>
__device__ __forceinline__ void fa2_wait_dbg(int tag, uint32_t bar, uint32_t parity) {
  const long long t0 = clock64();
  while (clock64() - t0 < 40000000ll) {  // 20 ms
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return;
  }
  if ((threadIdx.x & 31) == 0 && blockIdx.x == 0) printf("fa2 wait %d timed out: block %d warp %d parity %u\n", tag, blockIdx.x, threadIdx.x >> 5, parity);
}
#define FA2_WAIT(tag, ...) fa2_wait_dbg(tag, __VA_ARGS__)
#else
#define FA2_WAIT(tag, ...) mbar_wait(__VA_ARGS__)
#endif
// ------------------------------------------------------------------ persistent variant
// One CTA per SM hosts TWO independent groups (8 softmax warps + 1 TMA / MMA warp each, 256 of the 512 TMEM columns
// each) that walk (stream, batch, head, query tile) items: barrier initialisation, the TMEM allocation and the
// griddepcontrol hand-over are paid once per launch instead of once per item, the operands of a group's next item
// are fetched by TMA (double-buffered Q / K / V) and its key-mask bytes are read while the current item's
// exponentials run, and S of the next item is issued right behind the last PV MMAs.  With one-item CTAs two thirds of
// the items (2.6 waves at B = 8) paid ~4 k of their ~14 k cycles for exactly those phases
// (profiles/r01_attention_fa_phase_trace.txt).  Query-tile rows past Tq (72 of 128 rows in the second tile of
// T = 200) skip the exponentials.  Item arithmetic is the same as stream_attention_fa_kernel's.
constexpr int kGroups = 2;
constexpr int kGroupWarps = kSoftmaxWarps + 1;
constexpr int kThreadsFa2 = 32 * kGroupWarps * kGroups;

struct Fa2Smem {
  uint32_t op[2], q_hi, q_lo, k_hi, k_lo, v_hi, v_lo;  // operand buffers (offsets inside a buffer)
  uint32_t cls, xch, flag, bar, group_bytes, total, nbuf;
};
__host__ __device__ inline Fa2Smem fa2_smem_map(int nblk, int kbox) {
  Fa2Smem m;
  const uint32_t kv = (uint32_t(nblk) * uint32_t(kbox) * 32u + 1023u) & ~1023u;
  m.q_hi = 0, m.q_lo = QT * 32, m.k_hi = 2 * QT * 32, m.k_lo = m.k_hi + kv, m.v_hi = m.k_lo + kv, m.v_lo = m.v_hi + kv;
  const uint32_t buf_bytes = m.v_lo + kv;
  m.nbuf = nblk == 1 ? 2u : 1u;  // long key sequences: no room for a second K / V set
  m.op[0] = 0, m.op[1] = m.nbuf == 2 ? buf_bytes : 0;
  m.cls = buf_bytes * m.nbuf;
  m.xch = m.cls + (uint32_t(nblk) * KBLK + 32) * 4;
  m.flag = m.xch + 4 * 128 * 4;
  m.bar = (m.flag + uint32_t(nblk) * 7 * 4 + 15u) & ~15u;  // 14 mbarriers
  m.group_bytes = (m.bar + 14 * 8 + 1023u) & ~1023u;
  m.total = m.group_bytes * kGroups + 16 + 1024;
  return m;
}

template <int FMT>
__global__ void __launch_bounds__(kThreadsFa2, 1) stream_attention_fa2_kernel(const __grid_constant__ FaParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base0 = (raw + 1023u) & ~1023u;
  const Fa2Smem L = fa2_smem_map(P.nblk, P.kbox);
  const int warp = scatt_warp_idx(), lane = threadIdx.x & 31;
  const int grp = warp / kGroupWarps, wl = warp % kGroupWarps;
  const uint32_t base = base0 + uint32_t(grp) * L.group_bytes;
  uint8_t* sm = smem_raw + (base - raw);
  const uint32_t tmem_ptr_addr = base0 + kGroups * L.group_bytes;
  // barriers of this group
  const uint32_t bars = base + L.bar;
  const uint32_t bar_qk0 = bars, bar_v0 = bars + 16, bar_s = bars + 32, bar_sf = bars + 40, bar_o = bars + 48, bar_pk0 = bars + 56;  // [2],[2],1,1,1,[4]
  float* cls = reinterpret_cast<float*>(sm + L.cls);
  const int Tq = P.Tq, Tk = P.Tk, D = P.H * HD;
  const bool causal = P.kind == SCATT_ATTN_CAUSAL;
  const int kbox = P.kbox, nchunk = (kbox + 31) >> 5;
  const bool lo_q = P.terms >= 2, lo_k = P.terms >= 3;
  const int nqt = (Tq + QT - 1) / QT, nbh = P.groups * P.B * P.H, total_items = nbh * nqt;
  const int item0 = 2 * int(blockIdx.x) + grp, item_step = 2 * int(gridDim.x);
  constexpr uint32_t kOCol = 224;

  if (wl == kSoftmaxWarps && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_qk0 + 8 * i, 1);
      mbar_init(bar_v0 + 8 * i, 1);
    }
    mbar_init(bar_s, 1);
    mbar_init(bar_sf, 32 * kSoftmaxWarps);
    mbar_init(bar_o, 1);
    for (int i = 0; i < 4; ++i) mbar_init(bar_pk0 + 8 * i, 32 * kSoftmaxWarps);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int g = 0; g < P.groups; ++g) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&P.map_q[g]) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&P.map_k[g]) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&P.map_v[g]) : "memory");
    }
  }
  if (warp == kSoftmaxWarps) {  // group 0's MMA warp allocates all 512 columns
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_addr), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_launch_dependents();
  pdl_wait();  // q / k / v planes come from the preceding GEMM
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_ptr_addr - raw)) + uint32_t(grp) * 256u;
  const uint32_t tmem_s = tmem, tmem_o = tmem + kOCol;
  const uint32_t idesc_base = (1u << 4) | (uint32_t(FMT) << 7) | (uint32_t(FMT) << 10) | (uint32_t(QT >> 4) << 24);
#ifdef SCATT_FA2_DEBUG
  if (threadIdx.x == 0 && blockIdx.x == 0) printf("fa2: setup done, tmem %x items %d step %d stage %d\n", tmem, total_items, item_step, P.debug_stage);
  const bool dbg_skip_roles = P.debug_stage == 1;
#else
  constexpr bool dbg_skip_roles = false;
#endif

  auto decode = [&](int item, int& g, int& b, int& h, int& m0) {
    const int qt = item / nbh, r = item % nbh;  // full query tiles first: the partial last tile of every sequence ends the launch
    g = r / (P.B * P.H);
    b = (r / P.H) % P.B;
    h = r % P.H;
    m0 = qt * QT;
  };

  if (dbg_skip_roles) {
  } else if (wl == kSoftmaxWarps) {
    // ================= TMA + MMA warp of this group
    const uint32_t idesc_s = idesc_base | (uint32_t(kbox >> 3) << 17);
    const uint32_t idesc_o = idesc_base | (1u << 16) | (uint32_t(HD >> 3) << 17);  // bit 16: B is MN-major
    auto load_item = [&](int item, int w) {  // operands of the group's w-th item into buffer w % nbuf
      int g, b, h, m0;
      decode(item, g, b, h, m0);
      const FaProblem& A = P.p[g];
      const int nk = causal ? min(Tk, m0 + QT) : Tk, nblk = (nk + kbox - 1) / kbox;
      const uint32_t ob = base + L.op[w % int(L.nbuf)];
      const uint32_t bqk = bar_qk0 + 8 * (w % int(L.nbuf)), bv = bar_v0 + 8 * (w % int(L.nbuf));
      if (elect_one()) {
        const int qrow = b * Tq + m0, krow = b * Tk;
        const int nq = lo_q ? 2 : 1, nkpl = lo_k ? 2 : 1;
        mbar_expect_tx(bqk, uint32_t(QT * 32 * nq + nblk * kbox * 32 * nkpl));
        tma_load_3d(ob + L.q_hi, &P.map_q[g], bqk, A.q_col + h * HD, qrow, 0);
        if (lo_q) tma_load_3d(ob + L.q_lo, &P.map_q[g], bqk, A.q_col + h * HD, qrow, 1);
        for (int blk = 0; blk < nblk; ++blk) {
          tma_load_3d(ob + L.k_hi + blk * kbox * 32, &P.map_k[g], bqk, A.k_col + h * HD, krow + blk * kbox, 0);
          if (lo_k) tma_load_3d(ob + L.k_lo + blk * kbox * 32, &P.map_k[g], bqk, A.k_col + h * HD, krow + blk * kbox, 1);
        }
        mbar_expect_tx(bv, uint32_t(nblk * kbox * 32 * nkpl));
        for (int blk = 0; blk < nblk; ++blk) {
          tma_load_3d(ob + L.v_hi + blk * kbox * 32, &P.map_v[g], bv, A.v_col + h * HD, krow + blk * kbox, 0);
          if (lo_k) tma_load_3d(ob + L.v_lo + blk * kbox * 32, &P.map_v[g], bv, A.v_col + h * HD, krow + blk * kbox, 1);
        }
      }
      __syncwarp();
    };
    uint32_t sf_uses = 0, pk_uses[4] = {0, 0, 0, 0};
    if (item0 < total_items) load_item(item0, 0);
    if (L.nbuf == 2 && item0 + item_step < total_items) load_item(item0 + item_step, 1);
    int w = 0;
    for (int item = item0; item < total_items; item += item_step, ++w) {
      int g, b, h, m0;
      decode(item, g, b, h, m0);
      const int nk = causal ? min(Tk, m0 + QT) : Tk, nblk = (nk + kbox - 1) / kbox;
      const int bi = w % int(L.nbuf);
      const uint32_t ob = base + L.op[bi], par_op = uint32_t(w / int(L.nbuf)) & 1u;
      FA2_WAIT(1, bar_qk0 + 8 * bi, par_op);
      tc_fence_after();
      const uint64_t qh = umma_desc_sw32(ob + L.q_hi), ql = umma_desc_sw32(ob + L.q_lo);
      const uint64_t kh0 = umma_desc_sw32(ob + L.k_hi), kl0 = umma_desc_sw32(ob + L.k_lo);
      const uint64_t vh0 = umma_desc_sw32(ob + L.v_hi), vl0 = umma_desc_sw32(ob + L.v_lo);
      auto issue_s = [&](int blk) {
        const uint64_t off = uint64_t(blk * kbox * 32 >> 4);
        if (elect_one()) {
          uint32_t acc = 0;
          if (lo_k) {
            tc_mma_f16(tmem_s, qh, kl0 + off, idesc_s, acc);
            acc = 1;
          }
          if (lo_q) {
            tc_mma_f16(tmem_s, ql, kh0 + off, idesc_s, acc);
            acc = 1;
          }
          tc_mma_f16(tmem_s, qh, kh0 + off, idesc_s, acc);
          tc_commit(bar_s);
        }
        __syncwarp();
      };
      // S of this item may overwrite P of the previous one: its PV MMAs were issued earlier and tcgen05.mma executes in order
      issue_s(0);
      // the other operand buffer held item w - 1: free once that item's MMAs have retired (bar_o), then it takes item w + 1
      if (L.nbuf == 2 && w >= 1 && item + item_step < total_items) {
        FA2_WAIT(2, bar_o, uint32_t(w - 1) & 1u);
        load_item(item + item_step, w + 1);
      }
      for (int blk = 1; blk < nblk; ++blk) {  // pass A over the remaining key blocks
        FA2_WAIT(3, bar_sf, sf_uses++ & 1u);
        tc_fence_after();
        issue_s(blk);
      }
      FA2_WAIT(4, bar_sf, sf_uses++ & 1u);
      tc_fence_after();
      FA2_WAIT(5, bar_v0 + 8 * bi, par_op);
      uint32_t acc_o = 0;
      for (int blk = 0; blk < nblk; ++blk) {
        if (nblk > 1) issue_s(blk);
        const int nkeys = min(kbox, nk - blk * kbox);
        const int ksteps = (nkeys + 15) >> 4, niter = (((nkeys + 31) >> 5) + 1) >> 1;
        const uint64_t vh = vh0 + uint64_t(blk * kbox * 32 >> 4), vl = vl0 + uint64_t(blk * kbox * 32 >> 4);
        for (int it = 0; it < niter; ++it) {
          FA2_WAIT(6, bar_pk0 + 8 * it, pk_uses[it]++ & 1u);
          tc_fence_after();
          if (elect_one()) {
            const int ks1 = min(4 * it + 4, ksteps);
            for (int ks = 4 * it; ks < ks1; ++ks) {
              const uint32_t p_hi = tmem_s + uint32_t(ks >> 1) * 32 + uint32_t(ks & 1) * 8;
              const uint64_t voff = uint64_t(ks) * 32;
              if (lo_k) {
                tc_mma_ts(tmem_o, p_hi, vl + voff, idesc_o, acc_o);
                acc_o = 1;
              }
              if (lo_q) {
                tc_mma_ts(tmem_o, p_hi + 16, vh + voff, idesc_o, acc_o);
                acc_o = 1;
              }
              tc_mma_ts(tmem_o, p_hi, vh + voff, idesc_o, acc_o);
              acc_o = 1;
            }
          }
          __syncwarp();
          acc_o = 1;
        }
      }
      if (elect_one()) tc_commit(bar_o);
      __syncwarp();
      if (L.nbuf == 1 && item + item_step < total_items) {  // single operand set: the next item's loads wait for this item's MMAs
        FA2_WAIT(7, bar_o, uint32_t(w) & 1u);
        load_item(item + item_step, w + 1);
      }
    }
  } else {
    // ================= softmax warps: two threads per query row, row = quad * 32 + lane
    const int quad = warp & 3, half = wl >> 2;  // the hardware ties a warp to TMEM lane quadrant warp % 4
    const int r = quad * 32 + lane;
    const int gtid = wl * 32 + lane;  // 0..255 inside the group
    const uint32_t lane_addr = uint32_t(quad * 32) << 16;
    const uint32_t* chunk_valid = reinterpret_cast<const uint32_t*>(sm + L.flag);
    uint32_t* flags = reinterpret_cast<uint32_t*>(sm + L.flag);
    float* xch = reinterpret_cast<float*>(sm + L.xch);
    const float kLog2e = 1.4426950408889634f;
    const int bar_id = 1 + grp;
    auto group_sync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(32 * kSoftmaxWarps) : "memory"); };
    uint32_t s_uses = 0;
    // key-mask bytes of an item: thread i of the group holds keys i, i + 256, i + 512 (1 = valid; keys past Tk read as 1)
    auto fetch_mask = [&](int item, uint32_t (&mk)[3]) {
      int g, b, h, m0;
      decode(item, g, b, h, m0);
      const uint8_t* km = P.p[g].key_mask;
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const int k = gtid + 256 * q;
        mk[q] = (km != nullptr && k < Tk) ? uint32_t(km[int64_t(b) * Tk + k]) : 1u;
      }
    };
    uint32_t mk[3] = {1u, 1u, 1u};
    if (item0 < total_items) fetch_mask(item0, mk);
    int w = 0;
    for (int item = item0; item < total_items; item += item_step, ++w) {
      int g, b, h, m0;
      decode(item, g, b, h, m0);
      const FaProblem& A = P.p[g];
      const int i = m0 + r;
      const int nk = causal ? min(Tk, m0 + QT) : Tk, nblk = (nk + kbox - 1) / kbox;
      const int jmax = causal ? i : 0x7fffffff;
#ifdef SCATT_FA2_DEBUG
      const bool live = P.debug_stage != 2 && m0 + quad * 32 < Tq;
#else
      const bool live = m0 + quad * 32 < Tq;  // some row of this warp exists: otherwise only the hand-shakes run
#endif
      // key classes of this item (flat key index k: class / flag arrays are indexed like the S chunks)
      {
        const int kend = nblk * nchunk * 32;
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          const int k = gtid + 256 * q;
          if (k < kend) {  // warp-uniform: kend is a multiple of 32
            const bool in_block = nblk > 1 || k < kbox;
            const float c = (in_block && k < nk) ? (mk[q] ? 0.f : -FLT_MAX) : -INFINITY;
            cls[k] = c;
            const bool all_valid = __all_sync(0xffffffffu, c == 0.f);
            if (lane == 0) flags[k >> 5] = all_valid ? 1u : 0u;
          }
        }
        group_sync();
      }
      float v[32];
      float mx = -INFINITY;
#pragma unroll 1
      for (int blk = 0; blk < nblk; ++blk) {
        FA2_WAIT(8, bar_s, s_uses++ & 1u);
        tc_fence_after();
        const int nch = (min(kbox, nk - blk * kbox) + 31) >> 5;
        if (live) {
#pragma unroll 1
          for (int c = half; c < nch; c += 2) {
            tc_ld32(tmem_s + lane_addr + c * 32, v);
            const int cc = blk * nchunk + c, key0 = blk * kbox + c * 32;
            const bool fast = chunk_valid[cc] != 0u && (!causal || key0 + 31 <= m0 + quad * 32);
            if (fast) {
#pragma unroll
              for (int j = 0; j < 32; ++j) mx = fmaxf(mx, v[j]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float kc = cls[cc * 32 + j];
                float s = kc == 0.f ? v[j] : kc;
                if (key0 + j > jmax) s = -INFINITY;
                mx = fmaxf(mx, s);
              }
            }
          }
        }
        tc_fence_before();
        mbar_arrive(bar_sf);
      }
      // the next item's key-mask bytes travel while the exponentials run
      if (item + item_step < total_items) fetch_mask(item + item_step, mk);
      xch[half * 128 + r] = mx;
      group_sync();
      mx = fmaxf(mx, xch[(half ^ 1) * 128 + r]);
      const float mneg = -mx * kLog2e;
      float l = 0.f;
      float2 la = make_float2(0.f, 0.f), lb = la;
#pragma unroll 1
      for (int blk = 0; blk < nblk; ++blk) {
        if (nblk > 1) {
          FA2_WAIT(9, bar_s, s_uses++ & 1u);
          tc_fence_after();
        }
        const int nch = (min(kbox, nk - blk * kbox) + 31) >> 5;
        const int niter = (nch + 1) >> 1;
#pragma unroll 1
        for (int it = 0; it < niter; ++it) {
          const int c = 2 * it + half;
          if (c < nch && live) {
            tc_ld32(tmem_s + lane_addr + c * 32, v);
            const int cc = blk * nchunk + c, key0 = blk * kbox + c * 32;
            const bool fast = chunk_valid[cc] != 0u && (!causal || key0 + 31 <= m0 + quad * 32);
            if (fast) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float p0 = ex2(fmaf(v[j], kLog2e, mneg)), p1 = ex2(fmaf(v[j + 1], kLog2e, mneg));
                const float p2 = ex2(fmaf(v[j + 2], kLog2e, mneg)), p3 = ex2(fmaf(v[j + 3], kLog2e, mneg));
                la = __fadd2_rn(la, make_float2(p0, p1));
                lb = __fadd2_rn(lb, make_float2(p2, p3));
                v[j] = p0, v[j + 1] = p1, v[j + 2] = p2, v[j + 3] = p3;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float kc = cls[cc * 32 + j];
                float s = kc == 0.f ? v[j] : kc;
                if (key0 + j > jmax) s = -INFINITY;
                const float p = ex2((s - mx) * kLog2e);
                l += p;
                v[j] = p;
              }
            }
            float wv[32];
            uint32_t* wp = reinterpret_cast<uint32_t*>(wv);
#pragma unroll
            for (int q = 0; q < 16; ++q) split2<FMT>(v[2 * q], v[2 * q + 1], wp[q], wp[16 + q]);
            tc_st32(tmem_s + lane_addr + c * 32, wv);
          }
          tc_fence_before();
          mbar_arrive(bar_pk0 + 8 * it);
        }
      }
      l += (la.x + la.y) + (lb.x + lb.y);
      xch[256 + half * 128 + r] = l;
      group_sync();
      l += xch[256 + (half ^ 1) * 128 + r];
      FA2_WAIT(10, bar_o, uint32_t(w) & 1u);
      tc_fence_after();
      float o[8];
      tc_ld8(tmem_o + lane_addr + 8 * half, o);  // .sync.aligned: every lane of the warp, whether its row exists or not
      if (i < Tq) {
        const float inv = 1.0f / l;
        const int64_t off = (int64_t(b) * Tq + i) * D + h * HD + 8 * half;
#pragma unroll
        for (int c = 0; c < 8; ++c) o[c] *= inv;
        if (A.out) {
          *reinterpret_cast<float4*>(A.out + off) = make_float4(o[0], o[1], o[2], o[3]);
          *reinterpret_cast<float4*>(A.out + off + 4) = make_float4(o[4], o[5], o[6], o[7]);
        }
        if (A.out_planes) {
          uint4 ph, pl;
          split2<FMT>(o[0], o[1], ph.x, pl.x);
          split2<FMT>(o[2], o[3], ph.y, pl.y);
          split2<FMT>(o[4], o[5], ph.z, pl.z);
          split2<FMT>(o[6], o[7], ph.w, pl.w);
          *reinterpret_cast<uint4*>(A.out_planes + off) = ph;
          *reinterpret_cast<uint4*>(A.out_planes + int64_t(P.B) * Tq * D + off) = pl;
        }
      }
      // the next item's S may start PV-accumulating into O only after this group has read it: the hand-shake is
      // the P slices of the next item (the MMA warp waits for them), written after this point in program order
      tc_fence_before();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kSoftmaxWarps) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem - uint32_t(grp) * 256u), "r"(512u) : "memory");
  }
}

using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                              const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeFn get_encode_fa() {
  static EncodeFn fn = nullptr;
  static std::atomic<bool> done{false};
  if (!done.load()) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeFn>(p);
    done.store(true);
  }
  return fn;
}

// planes [2][rows][ld] -> boxes of 16 columns x box_rows rows x 1 plane, 32-byte swizzle
int encode_operand_map(CUtensorMap* map, const scatt_attn_operand& op, int box_rows, int fmt) {
  EncodeFn enc = get_encode_fa();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return SCATT_ERR_CUDA;
  }
  const cuuint64_t dims[3] = {cuuint64_t(op.ld), cuuint64_t(op.rows), 2};
  const cuuint64_t strides[2] = {cuuint64_t(op.ld) * 2, cuuint64_t(op.rows) * cuuint64_t(op.ld) * 2};
  const cuuint32_t box[3] = {HD, cuuint32_t(box_rows), 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, fmt == SCATT_PLANE_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3,
                   const_cast<void*>(op.planes), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(attention operand) failed with CUresult %d (rows=%lld ld=%lld box_rows=%d)", int(r),
              (long long)op.rows, (long long)op.ld, box_rows);
    return SCATT_ERR_CUDA;
  }
  return SCATT_OK;
}

}  // namespace

int debug_set_trace_fa(void* dev_buf) {
  long long* p = reinterpret_cast<long long*>(dev_buf);
  SCATT_CUDA(cudaMemcpyToSymbol(g_trace_fa, &p, sizeof(p)));
  return SCATT_OK;
}

constexpr int kPersistMinItems = 4096;
static std::atomic<int> g_attn_persist{[] {
  const char* e = std::getenv("SCATT_ATTN_PERSIST");
  return e ? (e[0] == '1' ? 1 : 2) : 0;
}()};
int debug_set_attn_persist(int mode) {
  SCATT_REQUIRE(mode >= 0 && mode <= 2, "debug_set_attn_persist: 0 (by item count), 1 (persistent) or 2 (one item per CTA)");
  g_attn_persist.store(mode, std::memory_order_relaxed);
  return SCATT_OK;
}

bool attention_planes_supported(int Tq, int Tk, int hd) { return hd == HD && Tk >= 1 && Tk <= KMAX; }
int attention_planes_max_keys() { return KMAX; }

int launch_attention_planes(const scatt_attention_planes_problem* p, int group, int B, int Tq, int Tk, int H, int hd, int kind,
                            int fmt, int terms, cudaStream_t s) {
  SCATT_REQUIRE(attention_planes_supported(Tq, Tk, hd), "attention(planes): needs head_dim 16 and 1 <= Tk <= %d", KMAX);
  SCATT_REQUIRE(terms >= 1 && terms <= 3, "attention(planes): terms must be 1..3");
  SCATT_REQUIRE(kind != SCATT_ATTN_CAUSAL || Tq == Tk, "attention(planes): causal needs Tq == Tk");
  SCATT_REQUIRE(int64_t(B) * group <= 65535 && H <= 65535, "attention(planes): grid too large");
  if (B == 0 || Tq == 0) return SCATT_OK;
  FaParams P{};
  P.B = B, P.Tq = Tq, P.Tk = Tk, P.H = H, P.kind = kind, P.terms = terms;
  P.nblk = (Tk + KBLK - 1) / KBLK;
  P.kbox = P.nblk == 1 ? ((Tk + 15) & ~15) : KBLK;
  for (int i = 0; i < group; ++i) {
    const scatt_attention_planes_problem& a = p[i];
    SCATT_REQUIRE(a.q.planes && a.k.planes && a.v.planes && (a.out || a.out_planes), "attention(planes): null operand");
    SCATT_REQUIRE(a.q.ld % 8 == 0 && a.k.ld % 8 == 0 && a.v.ld % 8 == 0 && a.q.col % 8 == 0 && a.k.col % 8 == 0 && a.v.col % 8 == 0,
                  "attention(planes): leading dimensions and column offsets must be multiples of 8");
    SCATT_REQUIRE(a.q.rows >= int64_t(B) * Tq && a.k.rows >= int64_t(B) * Tk && a.v.rows >= int64_t(B) * Tk,
                  "attention(planes): operand has fewer rows than B*T");
    int rc = encode_operand_map(&P.map_q[i], a.q, QT, fmt);
    if (rc == SCATT_OK) rc = encode_operand_map(&P.map_k[i], a.k, P.kbox, fmt);
    if (rc == SCATT_OK) rc = encode_operand_map(&P.map_v[i], a.v, P.kbox, fmt);
    if (rc != SCATT_OK) return rc;
    P.p[i] = FaProblem{a.key_mask, a.out, reinterpret_cast<uint16_t*>(a.out_planes), a.q.col, a.k.col, a.v.col};
  }
  P.groups = group;
  {
    const char* e = std::getenv("SCATT_FA2_STAGE");
    P.debug_stage = e ? std::atoi(e) : 0;
  }
  // Schedule: one item per CTA, or the persistent two-group kernel.  Measured with the warp-uniform role dispatch
  // (profiles/r02_toggles.txt): B = 8 (768 items per launch) 0.886 vs 0.917 ms per step - the one-item kernel wins
  // while the grid is a few waves; B = 64 / 256 (6 k / 25 k items) +1.6 % / +2.7 % frames/s for the persistent one.
  // 0 = by item count, 1 / 2 = force persistent / one-item (SCATT_ATTN_PERSIST=1|0, scatt_debug_set_attn_persist).
  const int items = group * B * H * ((Tq + QT - 1) / QT);
  const int mode = g_attn_persist.load(std::memory_order_relaxed);
  const bool persist = fa2_smem_map(P.nblk, P.kbox).total <= 227u * 1024u && (mode == 1 || (mode == 0 && items >= kPersistMinItems));
  if (persist) {
    const Fa2Smem L2 = fa2_smem_map(P.nblk, P.kbox);
    static PerDeviceFlag attr2_done;
    if (!attr2_done.load()) {
      SCATT_CUDA(cudaFuncSetAttribute(stream_attention_fa2_kernel<SCATT_PLANE_F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      SCATT_CUDA(cudaFuncSetAttribute(stream_attention_fa2_kernel<SCATT_PLANE_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      attr2_done.store(true);
    }
    SCATT_REQUIRE(L2.total <= 227u * 1024u, "attention(planes): shared memory map of the persistent kernel exceeds 227 KB");
    dim3 grid2(unsigned(min(148, (items + 1) / 2)));
    if (fmt == SCATT_PLANE_F16)
      (void)launch_kernel(stream_attention_fa2_kernel<SCATT_PLANE_F16>, grid2, dim3(kThreadsFa2), L2.total, s, P);
    else
      (void)launch_kernel(stream_attention_fa2_kernel<SCATT_PLANE_BF16>, grid2, dim3(kThreadsFa2), L2.total, s, P);
    const int rc2 = after_launch("stream_attention_fa2_kernel");
    set_last_kernel("stream_attention_fa2_kernel<%d>", fmt);
    return rc2;
  }
  const uint32_t kFaSmem = fa_smem_map(P.nblk, P.kbox).total;
  static PerDeviceFlag attr_done;
  if (!attr_done.load()) {
    const int max_smem = int(fa_smem_map(kMaxBlocks, KBLK).total);
    SCATT_CUDA(cudaFuncSetAttribute(stream_attention_fa_kernel<SCATT_PLANE_F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    SCATT_CUDA(cudaFuncSetAttribute(stream_attention_fa_kernel<SCATT_PLANE_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    attr_done.store(true);
  }
  dim3 grid((Tq + QT - 1) / QT, H, B * group);
  if (fmt == SCATT_PLANE_F16)
    (void)launch_kernel(stream_attention_fa_kernel<SCATT_PLANE_F16>, grid, dim3(kThreadsFa), kFaSmem, s, P);
  else
    (void)launch_kernel(stream_attention_fa_kernel<SCATT_PLANE_BF16>, grid, dim3(kThreadsFa), kFaSmem, s, P);
  const int rc = after_launch("stream_attention_fa_kernel");
  set_last_kernel("stream_attention_fa_kernel<%d>", fmt);
  return rc;
}

}  // namespace scatt
