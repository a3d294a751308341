// K5 fusion attention on the tensor cores (model/fusion.py:46-49 of the reference):
//
//     out[b] = softmax(q[b] k[b]^T) v[b]        single head of width D = 1024, no mask, no scaling
//
// q / k / v are the split planes [2][B*T][D] the three squeeze GEMMs wrote (GELU already applied), so nothing
// is converted or transposed here:
//
//   S = Q K^T    a K loop over D in 64-element chunks: Q [rows x 64] and K [keys x 64] tiles are K-major,
//                128-byte-swizzled UMMA operands; one TMA operation per operand and ring stage fetches `kch`
//                chunks of both planes (4-D tensor maps: element, row, chunk, plane); 3 product terms
//                (hi*lo + lo*hi + hi*hi) keep logits of magnitude ~200 at fp32 accuracy
//   softmax      two threads per query row straight out of TMEM (exact two-pass: max, then exp2 / sum);
//                P is written back over S as packed 16-bit hi | lo and is the TMEM A operand of the second GEMM
//   O = P V      N = `ncols` (<= 256) output columns per slice, V read MN-major exactly as stored
//                (keys x 64-channel swizzle atoms, `ncols / 64` atoms per MMA through the leading byte offset);
//                a CTA walks `spc` slices one after the other against the same P
//
// One CTA = one (batch element, 128-query tile, group of `spc` column slices).  Small batches split the D
// columns over several CTAs (each recomputes S: Q and K come out of L2), large ones keep all of a row in one
// CTA.  TMEM: S / P in columns [0, 256), the O slice in [256, 256 + ncols).  T <= 256 keys; longer sequences
// stay on the fp32 CUDA-core kernel (attention.cu).
#include <cstdlib>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace scatt {

namespace {

using namespace tc;

constexpr int kFuQT = 128;                                // query rows per tile (UMMA M)
constexpr int kFuMaxKeys = 256;                           // S columns
constexpr int kFuSoftmaxWarps = 8;                        // two per TMEM lane quadrant
constexpr int kFuThreads = 32 * kFuSoftmaxWarps + 96;     // + two TMA producer warps + the MMA warp
constexpr int kFuMaxStages = 8;
constexpr uint32_t kFuCtrlBytes = 4096;                   // barriers, TMEM pointer, row exchange
constexpr uint32_t kFuSlackBytes = 16384;                 // an M = 128 MMA reads 128 rows of a Q tile that may hold fewer
constexpr uint32_t kFuSmemBudget = 227u * 1024u;

struct alignas(64) FuParams {
  CUtensorMap map_q, map_k, map_v;
  float* out;
  uint16_t* out_planes;
  int32_t B, T, D, terms;
  int32_t qr;           // rows of a Q box (multiple of 8)
  int32_t tkp;          // keys rounded up to 16 (UMMA N of S)
  int32_t kch;          // 64-element K chunks per ring stage
  int32_t stages, stage_bytes;
  int32_t ncols, spc;   // O columns per slice, slices per CTA
};

// MN-major, 128-byte-swizzled B operand: rows = K index (keys) of 128 bytes (64 N-elements), 8-row groups 1024 B
// apart (stride byte offset), the next 64 N-elements `lbo` bytes further (leading byte offset).
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo) {
  uint64_t d = 0;
  d |= uint64_t((smem_addr & 0x3FFFFu) >> 4);
  d |= uint64_t(lbo >> 4) << 16;
  d |= uint64_t(1024 >> 4) << 32;
  d |= uint64_t(1) << 46;
  d |= uint64_t(2) << 61;
  return d;
}

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

__device__ __forceinline__ float fu_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int FMT>
__device__ __forceinline__ void fu_split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  if (FMT == SCATT_PLANE_F16) {
    const __half2 h = __floats2half2_rn(a, b);
    const float2 back = __half22float2(h);
    const __half2 l = __floats2half2_rn(a - back.x, b - back.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
  } else {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    const float2 back = __bfloat1622float2(h);
    const __nv_bfloat162 l = __floats2bfloat162_rn(a - back.x, b - back.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
  }
}

template <int FMT>
__global__ void __launch_bounds__(kFuThreads, 1) fusion_attention_tc_kernel(const __grid_constant__ FuParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  // control block: full[8] | empty[8] | bar_s | bar_p[4] | bar_o | bar_ofree | tmem pointer ; row exchange at +1024
  const uint32_t bar_full = base, bar_empty = base + 64, bar_s = base + 128, bar_p = base + 136, bar_o = base + 168;
  const uint32_t bar_ofree = base + 176, tmem_ptr_addr = base + 184;
  float* xch = reinterpret_cast<float*>(sm + 1024);  // [2][128] row max, [2][128] row sum
  const uint32_t ring = base + kFuCtrlBytes;

  const int warp = scatt_warp_idx(), lane = threadIdx.x & 31;
  const int b = blockIdx.z, m0 = blockIdx.y * kFuQT, slice0 = blockIdx.x * P.spc;
  const int T = P.T, D = P.D, tkp = P.tkp, kch = P.kch, qr = P.qr, ncols = P.ncols, stages = P.stages;
  const bool lo_q = P.terms >= 2, lo_k = P.terms >= 3;  // product terms: q_hi k_lo (3), q_lo k_hi (2), q_hi k_hi
  const int npq = lo_q ? 2 : 1, npk = lo_k ? 2 : 1;
  const int nqk = (D >> 6) / kch;                 // ring items of the first GEMM
  const int nvc = (tkp + 63) >> 6;                // 64-key chunks of the second GEMM
  const int nitems = nqk + P.spc * nvc;
  const uint32_t q_tile = uint32_t(qr) * 128u, k_tile = uint32_t(tkp) * 128u;  // bytes of one (plane, chunk) tile
  const uint32_t q_bytes = uint32_t(npq * kch) * q_tile, k_bytes = uint32_t(npk * kch) * k_tile;
  const uint32_t v_plane = uint32_t(ncols >> 6) * 8192u;  // one plane of a V item: ncols / 64 atoms of 64 keys x 128 B
  constexpr uint32_t kTmemCols = 512, kOCol = 256;
  constexpr int kQWarp = kFuSoftmaxWarps, kKWarp = kFuSoftmaxWarps + 1, kMmaWarp = kFuSoftmaxWarps + 2;

  if (threadIdx.x == 32 * kMmaWarp) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(bar_full + 8 * s, 2);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_s, 1);
    for (int it = 0; it < 4; ++it) mbar_init(bar_p + 8 * it, 32 * kFuSoftmaxWarps);
    mbar_init(bar_o, 1);
    mbar_init(bar_ofree, 32 * kFuSoftmaxWarps);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&P.map_q) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&P.map_k) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&P.map_v) : "memory");
  }
  if (warp == kMmaWarp) {
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_addr), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_launch_dependents();
  pdl_wait();  // the planes come from the squeeze GEMM in front of this launch
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(sm + 184);
  const uint32_t tmem_s = tmem, tmem_o = tmem + kOCol;

  if (warp == kQWarp || warp == kKWarp) {
    // ---------------- producers: warp kQWarp fetches Q tiles and the hi plane of V, warp kKWarp K tiles and V's lo plane
    const bool is_q = warp == kQWarp;
    const int qrow = b * T + m0, krow = b * T;
    for (int it = 0; it < nitems; ++it) {
      const int s = it % stages;
      mbar_wait(bar_empty + 8 * s, ((it / stages) & 1) ^ 1);
      if (elect_one()) {
        const uint32_t st = ring + uint32_t(s) * uint32_t(P.stage_bytes), full = bar_full + 8 * s;
        if (it < nqk) {
          if (is_q) {
            mbar_expect_tx(full, q_bytes);
            tma_load_4d(st, &P.map_q, full, 0, qrow, it * kch, 0);
          } else {
            mbar_expect_tx(full, k_bytes);
            tma_load_4d(st + q_bytes, &P.map_k, full, 0, krow, it * kch, 0);
          }
        } else {
          const int v = it - nqk, sl = v / nvc, vc = v % nvc;
          const int atom0 = (slice0 + sl) * (ncols >> 6);
          if (is_q || lo_k) {
            mbar_expect_tx(full, v_plane);
            tma_load_4d(st + (is_q ? 0u : v_plane), &P.map_v, full, 0, krow + vc * 64, atom0, is_q ? 0 : 1);
          } else {
            mbar_arrive(full);
          }
        }
      }
      __syncwarp();
    }
  } else if (warp == kMmaWarp) {
    // ---------------- MMA issuer (the whole warp walks the role uniformly, one elected lane issues)
    const uint32_t idesc_base = (1u << 4) | (uint32_t(FMT) << 7) | (uint32_t(FMT) << 10) | (uint32_t(kFuQT >> 4) << 24);
    const uint32_t idesc_s = idesc_base | (uint32_t(tkp >> 3) << 17);
    const uint32_t idesc_o = idesc_base | (1u << 16) | (uint32_t(ncols >> 3) << 17);  // bit 16: B is MN-major
    uint32_t acc = 0;
    for (int it = 0; it < nqk; ++it) {
      const int s = it % stages;
      mbar_wait(bar_full + 8 * s, (it / stages) & 1);
      tc_fence_after();
      const uint32_t st = ring + uint32_t(s) * uint32_t(P.stage_bytes);
      if (elect_one()) {
        for (int c = 0; c < kch; ++c) {
          const uint64_t a_hi = umma_desc_sw128(st + uint32_t(c) * q_tile), a_lo = umma_desc_sw128(st + uint32_t(kch + c) * q_tile);
          const uint64_t b_hi = umma_desc_sw128(st + q_bytes + uint32_t(c) * k_tile);
          const uint64_t b_lo = umma_desc_sw128(st + q_bytes + uint32_t(kch + c) * k_tile);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint64_t adv = uint64_t(kk * 32 >> 4);
            if (lo_k) {
              tc_mma_f16(tmem_s, a_hi + adv, b_lo + adv, idesc_s, acc);
              acc = 1;
            }
            if (lo_q) {
              tc_mma_f16(tmem_s, a_lo + adv, b_hi + adv, idesc_s, acc);
              acc = 1;
            }
            tc_mma_f16(tmem_s, a_hi + adv, b_hi + adv, idesc_s, acc);
            acc = 1;
          }
        }
        tc_commit(bar_empty + 8 * s);
        if (it == nqk - 1) tc_commit(bar_s);
      }
      __syncwarp();
      acc = 1;
    }
    for (int sl = 0; sl < P.spc; ++sl) {
      uint32_t acc_o = 0;
      if (sl > 0) {  // the epilogue warps have drained the previous slice out of the O columns
        mbar_wait(bar_ofree, (sl - 1) & 1);
        tc_fence_after();
      }
      for (int vc = 0; vc < nvc; ++vc) {
        const int it = nqk + sl * nvc + vc, s = it % stages;
        mbar_wait(bar_full + 8 * s, (it / stages) & 1);
        if (sl == 0) mbar_wait(bar_p + 8 * vc, 0);  // P of keys [64 vc, 64 vc + 64) is in TMEM
        tc_fence_after();
        const uint32_t st = ring + uint32_t(s) * uint32_t(P.stage_bytes);
        const int ksteps = min(4, (tkp - 64 * vc) >> 4);
        if (elect_one()) {
          for (int ks = 0; ks < ksteps; ++ks) {
            const int g = 4 * vc + ks;  // 16-key step: P hi in 8 columns of its 32-key chunk, lo 16 columns further
            const uint32_t p_hi = tmem_s + uint32_t(g >> 1) * 32 + uint32_t(g & 1) * 8;
            const uint64_t v_hi = umma_desc_mn_sw128(st + uint32_t(ks) * 2048u, 8192u);
            const uint64_t v_lo = umma_desc_mn_sw128(st + v_plane + uint32_t(ks) * 2048u, 8192u);
            if (lo_k) {
              tc_mma_f16_ts(tmem_o, p_hi, v_lo, idesc_o, acc_o);
              acc_o = 1;
            }
            if (lo_q) {
              tc_mma_f16_ts(tmem_o, p_hi + 16, v_hi, idesc_o, acc_o);
              acc_o = 1;
            }
            tc_mma_f16_ts(tmem_o, p_hi, v_hi, idesc_o, acc_o);
            acc_o = 1;
          }
          tc_commit(bar_empty + 8 * s);
          if (vc == nvc - 1) tc_commit(bar_o);
        }
        __syncwarp();
        acc_o = 1;
      }
    }
  } else {
    // ---------------- softmax + epilogue: row = (warp % 4) * 32 + lane (TMEM lane), warp / 4 picks the even or odd 32-key chunks
    const int quad = warp & 3, half = warp >> 2;
    const int r = quad * 32 + lane, i = m0 + r;
    const uint32_t lane_addr = uint32_t(quad * 32) << 16;
    const int nch = (tkp + 31) >> 5, niter = (nch + 1) >> 1;
    const float kLog2e = 1.4426950408889634f;
    float v[32];
    mbar_wait(bar_s, 0);
    tc_fence_after();
    float mx = -INFINITY;
#pragma unroll 1
    for (int c = half; c < nch; c += 2) {
      tc_ld32(tmem_s + lane_addr + c * 32, v);
      const int key0 = c * 32;
      if (key0 + 32 <= T) {
#pragma unroll
        for (int j = 0; j < 32; ++j) mx = fmaxf(mx, v[j]);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) mx = fmaxf(mx, key0 + j < T ? v[j] : -INFINITY);
      }
    }
    xch[half * 128 + r] = mx;
    asm volatile("bar.sync 1, %0;" ::"n"(32 * kFuSoftmaxWarps) : "memory");
    mx = fmaxf(mx, xch[(half ^ 1) * 128 + r]);
    float l = 0.f;
#pragma unroll 1
    for (int it = 0; it < niter; ++it) {
      const int c = 2 * it + half;
      if (c < nch) {
        tc_ld32(tmem_s + lane_addr + c * 32, v);
        const int key0 = c * 32;
        // subtract first: logits reach ~200 and (s - mx) is exact for the keys that matter
        if (key0 + 32 <= T) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            v[j] = fu_ex2((v[j] - mx) * kLog2e);
            l += v[j];
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            v[j] = key0 + j < T ? fu_ex2((v[j] - mx) * kLog2e) : 0.f;
            l += v[j];
          }
        }
        float w[32];
        uint32_t* wp = reinterpret_cast<uint32_t*>(w);
#pragma unroll
        for (int q = 0; q < 16; ++q) fu_split2<FMT>(v[2 * q], v[2 * q + 1], wp[q], wp[16 + q]);
        tc_st32(tmem_s + lane_addr + c * 32, w);
      }
      tc_fence_before();
      mbar_arrive(bar_p + 8 * it);
    }
    xch[256 + half * 128 + r] = l;
    asm volatile("bar.sync 1, %0;" ::"n"(32 * kFuSoftmaxWarps) : "memory");
    l += xch[256 + (half ^ 1) * 128 + r];
    const float inv = 1.0f / l;
    const int64_t plane_stride = int64_t(P.B) * T * D;
    const int hc = ncols >> 1;  // columns of a slice this thread stores
    for (int sl = 0; sl < P.spc; ++sl) {
      mbar_wait(bar_o, sl & 1);
      tc_fence_after();
      const int col_base = (slice0 + sl) * ncols + half * hc;
#pragma unroll 1
      for (int c0 = 0; c0 < hc; c0 += 32) {
        tc_ld32(tmem_o + lane_addr + uint32_t(half * hc + c0), v);
        if (i < T) {
          const int64_t off = (int64_t(b) * T + i) * D + col_base + c0;
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] *= inv;
          if (P.out) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(P.out + off + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          }
          if (P.out_planes) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              uint4 ph, pl;
              fu_split2<FMT>(v[j], v[j + 1], ph.x, pl.x);
              fu_split2<FMT>(v[j + 2], v[j + 3], ph.y, pl.y);
              fu_split2<FMT>(v[j + 4], v[j + 5], ph.z, pl.z);
              fu_split2<FMT>(v[j + 6], v[j + 7], ph.w, pl.w);
              *reinterpret_cast<uint4*>(P.out_planes + off + j) = ph;
              *reinterpret_cast<uint4*>(P.out_planes + plane_stride + off + j) = pl;
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(bar_ofree);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");
  }
}

using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                              const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeFn get_encode_fu() {
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

// planes [2][rows][D] seen as (64 elements, row, 64-element chunk, plane): boxes of 64 x box_rows x box_chunks x box_planes,
// 128-byte swizzle -> shared memory holds [plane][chunk][row][128 B], every (plane, chunk) tile a canonical UMMA tile
int encode_chunk_map(CUtensorMap* map, const void* planes, int64_t rows, int D, int box_rows, int box_chunks, int box_planes, int fmt) {
  EncodeFn enc = get_encode_fu();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return SCATT_ERR_CUDA;
  }
  const cuuint64_t dims[4] = {64, cuuint64_t(rows), cuuint64_t(D >> 6), 2};
  const cuuint64_t strides[3] = {cuuint64_t(D) * 2, 128, cuuint64_t(rows) * cuuint64_t(D) * 2};
  const cuuint32_t box[4] = {64, cuuint32_t(box_rows), cuuint32_t(box_chunks), cuuint32_t(box_planes)};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(map, fmt == SCATT_PLANE_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4,
                   const_cast<void*>(planes), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(fusion operand) failed with CUresult %d (rows=%lld D=%d box=%d x %d x %d)", int(r),
              (long long)rows, D, box_rows, box_chunks, box_planes);
    return SCATT_ERR_CUDA;
  }
  return SCATT_OK;
}

}  // namespace

bool fusion_attention_tc_supported(int T, int D) { return T >= 1 && T <= kFuMaxKeys && D >= 256 && D % 256 == 0 && D <= 4096; }

int launch_fusion_attention_tc(const void* q_planes, const void* k_planes, const void* v_planes, int B, int T, int D, float* out,
                               void* out_planes, int fmt, int terms, cudaStream_t s) {
  SCATT_REQUIRE(fusion_attention_tc_supported(T, D), "fusion_attention(planes): needs 1 <= T <= %d and D a multiple of 256 (got T=%d D=%d)",
                kFuMaxKeys, T, D);
  SCATT_REQUIRE(terms >= 1 && terms <= 3, "fusion_attention(planes): terms must be 1..3");
  SCATT_REQUIRE(B <= 65535, "fusion_attention(planes): batch too large for one launch");
  if (B == 0) return SCATT_OK;
  FuParams P{};
  P.out = out, P.out_planes = reinterpret_cast<uint16_t*>(out_planes);
  P.B = B, P.T = T, P.D = D, P.terms = terms;
  P.qr = (min(T, kFuQT) + 7) & ~7;
  P.tkp = (T + 15) & ~15;
  const int npq = terms >= 2 ? 2 : 1, npk = terms >= 3 ? 2 : 1;
  const int qtiles = (T + kFuQT - 1) / kFuQT;
  // Schedule.  A CTA computes S once and then `spc` column slices of O against it; CTAs of the same (batch, tile) pair
  // recompute S (Q and K come out of L2).  Cost model from profiles/r02_sweep_fusion.md: waves x (1 + 0.3 spc) with
  // waves = ceil(pairs x slices / spc / 148) - few pairs: one slice per CTA (narrow slices for the smallest grids),
  // many pairs: all of a row's columns in one CTA.
  const int env_ncols = [] { const char* e = std::getenv("SCATT_FUSION_NCOLS"); return e ? std::atoi(e) : 0; }();  // read per launch: tests / sweeps override the schedule
  const int env_spc = [] { const char* e = std::getenv("SCATT_FUSION_SPC"); return e ? std::atoi(e) : 0; }();
  const int pairs = B * qtiles;
  P.ncols = 256;
  if (pairs * (D / 256) < 96) P.ncols = 128;
  if (env_ncols == 64 || env_ncols == 128 || env_ncols == 256) P.ncols = env_ncols;
  const int nslices = D / P.ncols;
  P.spc = 1;
  {
    float best = 1e30f;
    for (int spc = 1; spc <= nslices; spc *= 2) {
      if (nslices % spc) continue;
      const int ctas = pairs * (nslices / spc);
      const float cost = float((ctas + 147) / 148) * (1.0f + 0.3f * float(spc) * float(P.ncols) / 256.0f);
      if (cost < best) best = cost, P.spc = spc;
    }
  }
  if (env_spc >= 1 && nslices % env_spc == 0) P.spc = env_spc;
  const uint32_t ring_budget = kFuSmemBudget - 1024u - kFuCtrlBytes - kFuSlackBytes;
  const uint32_t v_bytes = uint32_t(npk) * uint32_t(P.ncols >> 6) * 8192u;
  int kch = 4;
  uint32_t stage = 0;
  for (;; kch >>= 1) {
    stage = max(uint32_t(kch) * 128u * uint32_t(npq * P.qr + npk * P.tkp), v_bytes);
    if (kch == 1 || (ring_budget / stage >= 3 && (D >> 6) % kch == 0)) break;
  }
  P.kch = kch, P.stage_bytes = int(stage);
  P.stages = int(min(uint32_t(kFuMaxStages), ring_budget / stage));
  SCATT_REQUIRE(P.stages >= 2, "fusion_attention(planes): a ring stage of %u bytes does not fit twice", stage);
  const int64_t rows = int64_t(B) * T;
  int rc = encode_chunk_map(&P.map_q, q_planes, rows, D, P.qr, kch, npq, fmt);
  if (rc == SCATT_OK) rc = encode_chunk_map(&P.map_k, k_planes, rows, D, P.tkp, kch, npk, fmt);
  if (rc == SCATT_OK) rc = encode_chunk_map(&P.map_v, v_planes, rows, D, 64, P.ncols >> 6, 1, fmt);
  if (rc != SCATT_OK) return rc;
  static PerDeviceFlag attr_done;
  if (!attr_done.load()) {
    SCATT_CUDA(cudaFuncSetAttribute(fusion_attention_tc_kernel<SCATT_PLANE_F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kFuSmemBudget)));
    SCATT_CUDA(cudaFuncSetAttribute(fusion_attention_tc_kernel<SCATT_PLANE_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kFuSmemBudget)));
    attr_done.store(true);
  }
  const size_t smem = 1024u + kFuCtrlBytes + size_t(P.stages) * stage + kFuSlackBytes;
  dim3 grid(unsigned(nslices / P.spc), unsigned(qtiles), unsigned(B));
  if (fmt == SCATT_PLANE_F16)
    (void)launch_kernel(fusion_attention_tc_kernel<SCATT_PLANE_F16>, grid, dim3(kFuThreads), smem, s, P);
  else
    (void)launch_kernel(fusion_attention_tc_kernel<SCATT_PLANE_BF16>, grid, dim3(kFuThreads), smem, s, P);
  const int rc2 = after_launch("fusion_attention_tc_kernel");
  set_last_kernel("fusion_attention_tc_kernel<%d>", fmt);
  return rc2;
}

}  // namespace scatt
