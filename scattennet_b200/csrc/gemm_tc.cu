// tcgen05 engine of scatt_linear (sm_100a):
//   y = epilogue(x W^T + bias),  x and W given as 16-bit hi/lo split planes.
//
// One CTA computes a 128 x BN output tile.  Warp roles (192 threads):
//   warp 0   TMA producer   cp.async.bulk.tensor (3-D maps: K x rows x plane, 128B swizzle)
//   warp 1   TMEM allocator + single-thread tcgen05.mma issuer (accumulators in TMEM)
//   warps 2-5 epilogue      tcgen05.ld -> bias / scale / act / residual / LayerNorm -> global
// A `stages`-deep smem ring is handed between producer and issuer with
// full/empty mbarriers; the issuer signals the epilogue through a TMEM-full
// mbarrier (tcgen05.commit).  With `terms` = 3 every K step issues
// hi*hi + lo*hi + hi*lo so products are fp32-grade while running on the bf16 /
// fp16 tensor pipe; `terms` = 1 is the plain 16-bit product.
//
// LayerNorm is fused when the CTA owns the whole row (N == BN <= 256): the
// pre-norm value is written back to TMEM (tcgen05.st) during the statistics
// pass, so the residual is read once.
#include <cuda.h>

#include <mutex>

#include "common.cuh"

namespace scatt {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 x 2 B = one 128-byte swizzle row
constexpr int kThreads = 192;
constexpr uint32_t kWaitLimit = 1u << 22;  // bounded mbarrier spin: trap instead of hanging the GPU

struct TcProblem {
  const float* bias;
  const float* residual;
  const float* ln_g;
  const float* ln_b;
  float* y;
  uint16_t* y_planes;
};

struct alignas(64) TcParams {
  CUtensorMap map_a[SCATT_MAX_GROUP];
  CUtensorMap map_b[SCATT_MAX_GROUP];
  TcProblem prob[SCATT_MAX_GROUP];
  scatt_epilogue ep;
  int64_t M, ldres, ldy;
  int32_t N, K, stages, terms, fmt, fused_ln;
};

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  for (uint32_t it = 0; it < kWaitLimit; ++it) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return;
  }
  __trap();  // pipeline dead-lock: surface as a launch failure, never a hang
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_st32(uint32_t taddr, const float* v) {
  const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// K-major, 128-byte-swizzled operand tile: rows of 128 B, 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= uint64_t((smem_addr & 0x3FFFFu) >> 4);  // start address
  d |= uint64_t(1) << 16;                      // leading byte offset (unused for swizzled K-major)
  d |= uint64_t(1024 >> 4) << 32;              // stride byte offset: next 8-row group
  d |= uint64_t(1) << 46;                      // descriptor version (sm_100)
  d |= uint64_t(2) << 61;                      // SWIZZLE_128B
  return d;
}

// ------------------------------------------------------------------ epilogue
// One thread owns one output row (TMEM lane); columns arrive 32 at a time.
// All per-column parameters are fetched as float4 (warp-uniform addresses ->
// L1 broadcasts), every option is tested once per chunk, never per element,
// and the chunk loops are not unrolled: the body stays a few KB of SASS so it
// lives in the instruction cache (the first version of this epilogue was
// 136 KB of straight-line code and spent ~55 us per launch fetching it).

__device__ __forceinline__ void add_vec32(float* v, const float* __restrict__ p) {
#pragma unroll
  for (int j = 0; j < 32; j += 4) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p + j));
    v[j] += t.x, v[j + 1] += t.y, v[j + 2] += t.z, v[j + 3] += t.w;
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

// acc -> (acc + bias) * colscale -> act_pre -> (+ residual)
__device__ __forceinline__ void chunk_pre(const TcParams& P, const TcProblem& Q, float* v, int c0, int64_t row, bool row_ok,
                                          bool add_res) {
  if (Q.bias) add_vec32(v, Q.bias + c0);
  if (c0 < P.ep.scale_cols) {  // scale_cols is a multiple of 32 (checked on the host)
    const float s = P.ep.scale;
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] *= s;
  }
  act_vec32(v, P.ep.act_pre);
  if (add_res && row_ok) {
    const float* r = Q.residual + row * P.ldres + c0;
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 t = *reinterpret_cast<const float4*>(r + j);
      v[j] += t.x, v[j + 1] += t.y, v[j + 2] += t.z, v[j + 3] += t.w;
    }
  }
}

// act_post -> clamp -> y (fp32) and / or split planes
__device__ __forceinline__ void chunk_store(const TcParams& P, const TcProblem& Q, float* v, int c0, int64_t row, bool row_ok) {
  act_vec32(v, P.ep.act_post);
  if (P.ep.clamp > 0.f) {
    const float c = P.ep.clamp;
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fminf(fmaxf(v[j], -c), c);
  }
  if (!row_ok) return;
  if (Q.y) {
    float* y = Q.y + row * P.ldy + c0;
#pragma unroll
    for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(y + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
  }
  if (Q.y_planes) {
    uint16_t* hi = Q.y_planes + row * P.N + c0;
    uint16_t* lo = hi + P.M * int64_t(P.N);
    if (P.fmt == SCATT_PLANE_F16) {
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        uint32_t h[4], l[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const __half2 hh = __floats2half2_rn(v[j + 2 * e], v[j + 2 * e + 1]);
          const float2 back = __half22float2(hh);
          const __half2 ll = __floats2half2_rn(v[j + 2 * e] - back.x, v[j + 2 * e + 1] - back.y);
          h[e] = *reinterpret_cast<const uint32_t*>(&hh);
          l[e] = *reinterpret_cast<const uint32_t*>(&ll);
        }
        *reinterpret_cast<uint4*>(hi + j) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4*>(lo + j) = make_uint4(l[0], l[1], l[2], l[3]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        uint32_t h[4], l[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const __nv_bfloat162 hh = __floats2bfloat162_rn(v[j + 2 * e], v[j + 2 * e + 1]);
          const float2 back = __bfloat1622float2(hh);
          const __nv_bfloat162 ll = __floats2bfloat162_rn(v[j + 2 * e] - back.x, v[j + 2 * e + 1] - back.y);
          h[e] = *reinterpret_cast<const uint32_t*>(&hh);
          l[e] = *reinterpret_cast<const uint32_t*>(&ll);
        }
        *reinterpret_cast<uint4*>(hi + j) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4*>(lo + j) = make_uint4(l[0], l[1], l[2], l[3]);
      }
    }
  }
}

template <int BN, bool FUSED_LN>
__device__ __forceinline__ void epilogue_rows(const TcParams& P, const TcProblem& Q, uint32_t tmem_acc, int64_t row,
                                              int n0, bool row_ok) {
  const scatt_epilogue& ep = P.ep;
  constexpr int kChunks = BN / 32;
  float v[32];

  if constexpr (!FUSED_LN) {
    const bool add_res = ep.residual_mode != SCATT_RES_NONE;  // no LayerNorm here: before == after
#pragma unroll 1
    for (int c = 0; c < kChunks; ++c) {
      const int c0 = n0 + c * 32;
      if (c0 >= P.N) break;  // N is a multiple of 32; warp-uniform
      tc_ld32(tmem_acc + c * 32, v);
      chunk_pre(P, Q, v, c0, row, row_ok, add_res);
      chunk_store(P, Q, v, c0, row, row_ok);
    }
  } else {
    // LayerNorm over the BN == N columns of this row (n0 == 0).  Pass 1 builds the
    // pre-norm value, accumulates shifted sums and parks the value back in TMEM.
    float shift = 0.f, s1 = 0.f, s2 = 0.f;
    const bool res_before = ep.residual_mode == SCATT_RES_BEFORE_LN;
#pragma unroll 1
    for (int c = 0; c < kChunks; ++c) {
      tc_ld32(tmem_acc + c * 32, v);
      chunk_pre(P, Q, v, c * 32, row, row_ok, res_before);
      if (c == 0) shift = v[0];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float d = v[j] - shift;
        s1 += d;
        s2 = fmaf(d, d, s2);
      }
      tc_st32(tmem_acc + c * 32, v);
    }
    const float inv_n = 1.0f / float(BN);
    const float dm = s1 * inv_n;
    const float mean = shift + dm;
    const float rstd = rsqrtf(fmaxf(s2 * inv_n - dm * dm, 0.f) + ep.ln_eps);
    const bool res_after = ep.residual_mode == SCATT_RES_AFTER_LN;
#pragma unroll 1
    for (int c = 0; c < kChunks; ++c) {
      const int c0 = c * 32;
      tc_ld32(tmem_acc + c0, v);
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(Q.ln_g + c0 + j));
        const float4 b = __ldg(reinterpret_cast<const float4*>(Q.ln_b + c0 + j));
        v[j] = (v[j] - mean) * rstd * g.x + b.x;
        v[j + 1] = (v[j + 1] - mean) * rstd * g.y + b.y;
        v[j + 2] = (v[j + 2] - mean) * rstd * g.z + b.z;
        v[j + 3] = (v[j + 3] - mean) * rstd * g.w + b.w;
      }
      if (res_after && row_ok) {
        const float* r = Q.residual + row * P.ldres + c0;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 t = *reinterpret_cast<const float4*>(r + j);
          v[j] += t.x, v[j + 1] += t.y, v[j + 2] += t.z, v[j + 3] += t.w;
        }
      }
      chunk_store(P, Q, v, c0, row, row_ok);
    }
  }
}


template <int BN, bool FUSED_LN>
__global__ void __launch_bounds__(kThreads, 1) linear_tc_kernel(const __grid_constant__ TcParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [stages][A_hi | A_lo | B_hi | B_lo] tiles, then barriers
  constexpr uint32_t kABytes = BM * 128, kBBytes = BN * 128;
  const bool need_a_lo = P.terms >= 2, need_b_lo = P.terms >= 3;
  const uint32_t kBOff = kABytes * (need_a_lo ? 2 : 1);  // B tiles follow the A plane(s) of a stage
  const uint32_t kStageBytes = kBOff + kBBytes * (need_b_lo ? 2 : 1);
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int stages = P.stages;
  const uint32_t bar_base = base + stages * kStageBytes;  // full[stages], empty[stages], tmem_full, tmem_ptr
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (stages + s); };
  const uint32_t tmem_full_bar = bar_base + 16u * stages;
  const uint32_t tmem_ptr_addr = tmem_full_bar + 8u;
  volatile uint32_t* tmem_ptr_gen =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_ptr_addr - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = blockIdx.z;
  const int n0 = blockIdx.x * BN;
  const int64_t m0 = int64_t(blockIdx.y) * BM;
  const int num_kb = (P.K + BK - 1) / BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&P.map_a[g]) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&P.map_b[g]) : "memory");
  }
  if (warp == 1) {  // TMEM allocation (whole warp, .sync.aligned)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_addr), "r"(uint32_t(BN))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = *tmem_ptr_gen;

  if (warp == 0) {
    if (lane == 0) {  // ---------------- TMA producer
      const uint32_t tx = kStageBytes;
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % stages;
        mbar_wait(empty_bar(s), ((kb / stages) & 1) ^ 1);
        const uint32_t st = base + s * kStageBytes;
        mbar_expect_tx(full_bar(s), tx);
        tma_load_3d(st, &P.map_a[g], full_bar(s), kb * BK, int(m0), 0);
        if (need_a_lo) tma_load_3d(st + kABytes, &P.map_a[g], full_bar(s), kb * BK, int(m0), 1);
        tma_load_3d(st + kBOff, &P.map_b[g], full_bar(s), kb * BK, n0, 0);
        if (need_b_lo) tma_load_3d(st + kBOff + kBBytes, &P.map_b[g], full_bar(s), kb * BK, n0, 1);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ---------------- MMA issuer
      // instruction descriptor: D=f32, A/B = f16|bf16, both K-major, N, M=128
      const uint32_t idesc = (1u << 4) | (uint32_t(P.fmt) << 7) | (uint32_t(P.fmt) << 10) | (uint32_t(BN >> 3) << 17) |
                             (uint32_t(BM >> 4) << 24);
      uint32_t accumulate = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % stages;
        mbar_wait(full_bar(s), (kb / stages) & 1);
        tc_fence_after();
        const uint32_t st = base + s * kStageBytes;
        const uint64_t a_hi = umma_desc_sw128(st), a_lo = umma_desc_sw128(st + kABytes);
        const uint64_t b_hi = umma_desc_sw128(st + kBOff), b_lo = umma_desc_sw128(st + kBOff + kBBytes);
#pragma unroll
        for (int kk = 0; kk < BK / 16; ++kk) {
          const uint64_t adv = uint64_t(kk * 32 >> 4);  // 16 elements x 2 B along K inside the swizzle row
          if (need_b_lo) {
            tc_mma_f16(tmem_acc, a_hi + adv, b_lo + adv, idesc, accumulate);
            accumulate = 1;
          }
          if (need_a_lo) {
            tc_mma_f16(tmem_acc, a_lo + adv, b_hi + adv, idesc, accumulate);
            accumulate = 1;
          }
          tc_mma_f16(tmem_acc, a_hi + adv, b_hi + adv, idesc, accumulate);
          accumulate = 1;
        }
        tc_commit(empty_bar(s));  // smem slot reusable once these MMAs retire
      }
      tc_commit(tmem_full_bar);  // accumulator complete
    }
  } else {  // ---------------- epilogue warps 2..5
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access
    const int64_t row = m0 + quad * 32 + lane;
    epilogue_rows<BN, FUSED_LN>(P, P.prob[g], tmem_acc + (uint32_t(quad * 32) << 16), row, n0, row < P.M);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"(uint32_t(BN)) : "memory");
  }
}

// ------------------------------------------------------------------ host side
using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                              const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeFn get_encode() {
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

int encode_planes_map(CUtensorMap* map, const void* planes, int64_t rows, int K, int box_rows, int fmt) {
  EncodeFn enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return SCATT_ERR_CUDA;
  }
  const cuuint64_t dims[3] = {cuuint64_t(K), cuuint64_t(rows), 2};
  const cuuint64_t strides[2] = {cuuint64_t(K) * 2, cuuint64_t(rows) * cuuint64_t(K) * 2};
  const cuuint32_t box[3] = {BK, cuuint32_t(box_rows), 1};
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

template <int BN, bool FUSED_LN>
int launch_bn(TcParams& P, int group, cudaStream_t s) {
  const uint32_t kStageBytes = BM * 128 * (P.terms >= 2 ? 2 : 1) + BN * 128 * (P.terms >= 3 ? 2 : 1);
  const int num_kb = (P.K + BK - 1) / BK;
  int stages = int((200u * 1024u) / kStageBytes);
  if (stages > num_kb) stages = num_kb;
  if (stages > 8) stages = 8;
  if (stages < 1) stages = 1;
  P.stages = stages;
  const size_t smem = size_t(stages) * kStageBytes + 1024 /*align slack*/ + 16 * stages + 16;
  static std::atomic<bool> attr_done{false};
  if (!attr_done.load()) {
    SCATT_CUDA(cudaFuncSetAttribute(linear_tc_kernel<BN, FUSED_LN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_done.store(true);
  }
  dim3 grid((P.N + BN - 1) / BN, unsigned((P.M + BM - 1) / BM), group);
  linear_tc_kernel<BN, FUSED_LN><<<grid, kThreads, smem, s>>>(P);
  return after_launch("linear_tc_kernel");
}

}  // namespace

int launch_linear_tc(const scatt_linear_problem* p, int group, int64_t M, int N, int K, int64_t ldres, int64_t ldy,
                     const scatt_epilogue& ep, int fmt, int terms, cudaStream_t s) {
  SCATT_REQUIRE(terms >= 1 && terms <= 3, "linear(tcgen05): terms must be 1, 2 or 3");
  SCATT_REQUIRE(K % 8 == 0 && N % 32 == 0, "linear(tcgen05): K=%d must be a multiple of 8 and N=%d of 32", K, N);
  SCATT_REQUIRE(ep.scale_cols % 32 == 0, "linear(tcgen05): scale_cols must be a multiple of 32");
  SCATT_REQUIRE(ldres % 4 == 0 && ldy % 4 == 0, "linear(tcgen05): row strides must be multiples of 4");
  SCATT_REQUIRE(M < (int64_t(1) << 31), "linear(tcgen05): M too large");
  if (M == 0) return SCATT_OK;
  const bool fused_ln = ep.layer_norm && N == 256;
  // LayerNorm wider than one tile: GEMM with the pre-norm part of the chain, then the row-wise tail in place.
  const bool split_ln = ep.layer_norm && !fused_ln;
  const int BN = (fused_ln || N % 256 == 0) ? 256 : 128;

  TcParams P{};
  P.ep = ep;
  if (split_ln) {
    P.ep.layer_norm = 0;
    P.ep.act_post = SCATT_ACT_NONE;
    P.ep.clamp = 0.f;
    if (ep.residual_mode == SCATT_RES_AFTER_LN) P.ep.residual_mode = SCATT_RES_NONE;
  }
  P.M = M, P.N = N, P.K = K, P.ldres = ldres, P.ldy = ldy, P.terms = terms, P.fmt = fmt, P.fused_ln = fused_ln ? 1 : 0;
  for (int i = 0; i < group; ++i) {
    SCATT_REQUIRE(p[i].x_planes && p[i].w_planes, "linear(tcgen05): problem %d lacks split planes", i);
    SCATT_REQUIRE(ep.residual_mode == SCATT_RES_NONE || p[i].residual, "linear(tcgen05): residual missing");
    SCATT_REQUIRE(!ep.layer_norm || (p[i].ln_g && p[i].ln_b), "linear(tcgen05): LayerNorm needs gamma and beta");
    SCATT_REQUIRE(!split_ln || p[i].y, "linear(tcgen05): LayerNorm with N != 256 needs y as scratch");
    SCATT_REQUIRE(p[i].y || p[i].y_planes, "linear(tcgen05): no output");
    int rc = encode_planes_map(&P.map_a[i], p[i].x_planes, M, K, BM, fmt);
    if (rc != SCATT_OK) return rc;
    rc = encode_planes_map(&P.map_b[i], p[i].w_planes, N, K, BN, fmt);
    if (rc != SCATT_OK) return rc;
    P.prob[i] = TcProblem{p[i].bias, p[i].residual, p[i].ln_g, p[i].ln_b, p[i].y,
                          split_ln ? nullptr : reinterpret_cast<uint16_t*>(p[i].y_planes)};
  }
  int rc = fused_ln ? launch_bn<256, true>(P, group, s)
                    : (BN == 256 ? launch_bn<256, false>(P, group, s) : launch_bn<128, false>(P, group, s));
  if (rc != SCATT_OK || !split_ln) return rc;
  for (int i = 0; i < group; ++i) {
    rc = launch_rowwise(p[i].y, M, N, ldy, p[i].residual, ldres, p[i].ln_g, p[i].ln_b, ep, p[i].y, ldy, p[i].y_planes,
                        fmt, s);
    if (rc != SCATT_OK) return rc;
  }
  return SCATT_OK;
}

}  // namespace scatt
