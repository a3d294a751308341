// K3 stream attention on the 5th-generation tensor cores (tcgen05 + TMEM).
//
// One CTA = one (batch, head, 128-query tile).  head_dim is 16, so
//   S = Q K^T   is ONE tcgen05.mma per product term (M=128, N=keys<=256, K=16),
//   O = P V     is keys/16 MMAs of shape M=128, N=16, K=16,
// and the kernel is bound by the fp32 softmax between them, not by the MMAs.
//
// Operands are 16-bit hi/lo split planes built in shared memory by the CTA
// itself from the fp32 q / k / v rows (K-major, 128-byte swizzle - the layout
// the linear kernel gets from TMA): `terms` = 3 issues hi*lo + lo*hi + hi*hi for
// both contractions (fp32-grade products), `terms` = 1 the plain 16-bit product.
//
//   warps 0-3  load + split operands; then one thread per query row: read the
//              S row from TMEM, exact two-pass softmax (max, then exp / sum) with
//              the reference's mask semantics, write P (hi/lo, packed 16-bit) back
//              into the TMEM columns S occupied - the second contraction takes its
//              A operand straight from TMEM (no shared-memory round trip for P);
//              finally scale and store the O row
//   warp 4     TMEM allocation and the single MMA-issuing thread
//
// Mask semantics (model/utils.py:3-28, model/attention.py:63-72,165-171): a
// padded key's logit is exactly finfo(float32).min (so a row whose permitted
// keys are all padded is uniform over them), causal rows see keys j <= i only.
#include <cfloat>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace scatt {

namespace {

using namespace tc;

constexpr int HD = 16;
constexpr int QT = 128;      // queries per CTA
constexpr int KMAX = 256;    // keys per CTA (one S tile)
constexpr int kThreadsAtt = 160;

__device__ long long* g_trace_att = nullptr;
__device__ __forceinline__ void trace(int slot) {
  if (g_trace_att != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) g_trace_att[slot] = clock64();
}

struct AttnTcParams {
  scatt_attention_problem p[SCATT_MAX_GROUP];
  int32_t B, Tq, Tk, H, kind, fmt, terms;
  int64_t ldq, ldk, ldv;
};

// shared memory map (bytes, relative to a 1024-aligned base)
constexpr uint32_t kQOff = 0;                       // [128 rows][128 B]: hi at k 0..15, lo at k 16..31
constexpr uint32_t kKOff = kQOff + QT * 128;        // [256 rows][128 B]: same packing
constexpr uint32_t kVhOff = kKOff + KMAX * 128;     // V^T hi: 4 key blocks x [16 rows][128 B]
constexpr uint32_t kVlOff = kVhOff + 4 * 2048;      // V^T lo
constexpr uint32_t kPadOff = kVlOff + 4 * 2048;     // float[256] key class: 0 valid / -FLT_MAX padded / -inf absent
constexpr uint32_t kBarOff = kPadOff + KMAX * 4;    // 3 mbarriers + tmem pointer
constexpr uint32_t kSmemBytes = kBarOff + 64 + 1024;

// byte offset of element (row, k) inside a K-major 128B-swizzled tile whose rows are 128 B
__device__ __forceinline__ uint32_t sw128(uint32_t row, uint32_t byte_in_row) {
  const uint32_t chunk = (byte_in_row >> 4) ^ (row & 7);
  return (row >> 3) * 1024 + (row & 7) * 128 + chunk * 16 + (byte_in_row & 15);
}

template <int FMT>
__device__ __forceinline__ void split4(const float4& x, uint2& hi, uint2& lo) {
  if (FMT == SCATT_PLANE_F16) {
    const __half2 h0 = __floats2half2_rn(x.x, x.y), h1 = __floats2half2_rn(x.z, x.w);
    const float2 b0 = __half22float2(h0), b1 = __half22float2(h1);
    const __half2 l0 = __floats2half2_rn(x.x - b0.x, x.y - b0.y), l1 = __floats2half2_rn(x.z - b1.x, x.w - b1.y);
    hi = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
    lo = make_uint2(*reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&l1));
  } else {
    const __nv_bfloat162 h0 = __floats2bfloat162_rn(x.x, x.y), h1 = __floats2bfloat162_rn(x.z, x.w);
    const float2 b0 = __bfloat1622float2(h0), b1 = __bfloat1622float2(h1);
    const __nv_bfloat162 l0 = __floats2bfloat162_rn(x.x - b0.x, x.y - b0.y), l1 = __floats2bfloat162_rn(x.z - b1.x, x.w - b1.y);
    hi = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
    lo = make_uint2(*reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&l1));
  }
}

template <int FMT>
__device__ __forceinline__ void split1(float x, uint16_t& hi, uint16_t& lo) {
  split16<FMT>(x, hi, lo);
}

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// D[tmem] (+)= A[tmem] * B[smem]: the A operand (P) is read from tensor memory
__device__ __forceinline__ void tc_mma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
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
__global__ void __launch_bounds__(kThreadsAtt, 2) stream_attention_tc_kernel(const __grid_constant__ AttnTcParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  const uint32_t bar_s = base + kBarOff, bar_p = bar_s + 8, bar_o = bar_s + 16, tmem_ptr_addr = bar_s + 24;
  float* pad = reinterpret_cast<float*>(sm + kPadOff);

  const int warp = scatt_warp_idx(), lane = threadIdx.x & 31;
  const int g = blockIdx.z / P.B, b = blockIdx.z % P.B, h = blockIdx.y;
  const scatt_attention_problem& A = P.p[g];
  const int m0 = blockIdx.x * QT;
  const int Tq = P.Tq, Tk = P.Tk, D = P.H * HD;
  const bool causal = P.kind == SCATT_ATTN_CAUSAL;
  const int nk = causal ? min(Tk, m0 + QT) : Tk;       // keys this tile can see
  const int nkp = (nk + 15) & ~15;                      // MMA N / K granularity
  // S / P take nkp columns, O the next 16: 256 columns (two CTAs per SM) up to 224 keys, else all 512
  const uint32_t tmem_cols = nkp <= 224 ? 256u : 512u;
  const uint32_t o_col = nkp <= 224 ? 224u : 256u;

  if (threadIdx.x == 0) trace(0);
  if (threadIdx.x == 0) {
    mbar_init(bar_s, 1);
    mbar_init(bar_p, 128);
    mbar_init(bar_o, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_addr), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }

  pdl_launch_dependents();
  pdl_wait();
  // ---------------- operand staging: fp32 global rows -> split 16-bit swizzled tiles
  // Q and K: row r holds hi(k 0..15) in bytes 0..31 and lo(k 0..15) in bytes 32..63
  {
    const int total = (QT + nkp) * 4;
    constexpr int kBatch = 4;  // independent 16-byte loads in flight per thread
    for (int i0 = threadIdx.x; i0 < total; i0 += kThreadsAtt * kBatch) {
      float4 x[kBatch];
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        const int i = i0 + u * kThreadsAtt;
        const int r = i >> 2, c = i & 3;  // c: which float4 of the 16-wide head row
        x[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < total) {
          if (r < QT) {
            if (m0 + r < Tq) x[u] = *reinterpret_cast<const float4*>(A.q + (int64_t(b) * Tq + m0 + r) * P.ldq + h * HD + 4 * c);
          } else if (r - QT < nk) {
            x[u] = *reinterpret_cast<const float4*>(A.k + (int64_t(b) * Tk + r - QT) * P.ldk + h * HD + 4 * c);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        const int i = i0 + u * kThreadsAtt;
        if (i < total) {
          const int r = i >> 2, c = i & 3;
          const bool is_q = r < QT;
          const int row = is_q ? r : r - QT;
          uint2 hi, lo;
          split4<FMT>(x[u], hi, lo);
          uint8_t* tile = sm + (is_q ? kQOff : kKOff);
          *reinterpret_cast<uint2*>(tile + sw128(row, 8 * c)) = hi;
          *reinterpret_cast<uint2*>(tile + sw128(row, 32 + 8 * c)) = lo;
        }
      }
    }
  }
  if (threadIdx.x == 0) trace(1);
  // V^T: B operand of O = P V, [16 rows (head dim)][keys], 64 keys per 2 KB block
  {
    const int total = nkp * 4;
    constexpr int kBatch = 3;
    for (int i0 = threadIdx.x; i0 < total; i0 += kThreadsAtt * kBatch) {
      float4 x[kBatch];
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        const int i = i0 + u * kThreadsAtt;
        const int j = i >> 2, c = i & 3;
        x[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < total && j < nk) x[u] = *reinterpret_cast<const float4*>(A.v + (int64_t(b) * Tk + j) * P.ldv + h * HD + 4 * c);
      }
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        const int i = i0 + u * kThreadsAtt;
        if (i < total) {
          const int j = i >> 2, c = i & 3;
          const float xs[4] = {x[u].x, x[u].y, x[u].z, x[u].w};
          const uint32_t blk = (j >> 6) * 2048, kb = (j & 63) * 2;
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            uint16_t hi, lo;
            split1<FMT>(xs[e], hi, lo);
            const uint32_t off = blk + sw128(4 * c + e, kb);
            *reinterpret_cast<uint16_t*>(sm + kVhOff + off) = hi;
            *reinterpret_cast<uint16_t*>(sm + kVlOff + off) = lo;
          }
        }
      }
    }
  }
  for (int j = threadIdx.x; j < KMAX; j += kThreadsAtt) {
    float cls = -INFINITY;  // absent key (beyond nk): probability exactly 0
    if (j < nk) cls = (A.key_mask && A.key_mask[int64_t(b) * Tk + j] == 0) ? -FLT_MAX : 0.f;
    pad[j] = cls;
  }
  if (threadIdx.x == 0) trace(2);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) trace(3);
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(sm + kBarOff + 24);
  const uint32_t tmem_s = tmem, tmem_o = tmem + o_col;

  // instruction descriptors: D = f32, A/B = f16|bf16 K-major, M = 128
  const uint32_t idesc_base = (1u << 4) | (uint32_t(FMT) << 7) | (uint32_t(FMT) << 10) | (uint32_t(QT >> 4) << 24);

  if (warp == 4) {
    if (lane == 0) {
      // ---- S = Q K^T (K = 16: one MMA per term)
      const uint32_t idesc_s = idesc_base | (uint32_t(nkp >> 3) << 17);
      const uint64_t qd = umma_desc_sw128(base + kQOff), kd = umma_desc_sw128(base + kKOff);
      uint32_t acc = 0;
      if (P.terms >= 3) {
        tc_mma_f16(tmem_s, qd, kd + 2, idesc_s, acc);  // q_hi * k_lo
        acc = 1;
      }
      if (P.terms >= 2) {
        tc_mma_f16(tmem_s, qd + 2, kd, idesc_s, acc);  // q_lo * k_hi
        acc = 1;
      }
      tc_mma_f16(tmem_s, qd, kd, idesc_s, acc);        // q_hi * k_hi
      tc_commit(bar_s);
      // ---- O = P V once the softmax warps have written P
      mbar_wait(bar_p, 0);
      tc_fence_after();
      const uint32_t idesc_o = idesc_base | (uint32_t(HD >> 3) << 17);
      acc = 0;
      for (int ks = 0; ks < nkp / 16; ++ks) {
        // P of keys [16 ks, 16 ks + 16): hi in 8 TMEM columns, lo 16 columns further (see the softmax warps)
        const uint32_t p_hi = tmem_s + 32 * (ks >> 1) + 8 * (ks & 1), p_lo = p_hi + 16;
        const uint32_t blk = ks >> 2, adv = (ks & 3) * 2;
        const uint64_t vh = umma_desc_sw128(base + kVhOff + blk * 2048) + adv;
        const uint64_t vl = umma_desc_sw128(base + kVlOff + blk * 2048) + adv;
        if (P.terms >= 3) {
          tc_mma_f16_ts(tmem_o, p_hi, vl, idesc_o, acc);
          acc = 1;
        }
        if (P.terms >= 2) {
          tc_mma_f16_ts(tmem_o, p_lo, vh, idesc_o, acc);
          acc = 1;
        }
        tc_mma_f16_ts(tmem_o, p_hi, vh, idesc_o, acc);
        acc = 1;
      }
      tc_commit(bar_o);
    }
  } else {
    // ---------------- softmax: thread = query row (TMEM lane = warp * 32 + lane)
    const int r = warp * 32 + lane;
    const int i = m0 + r;  // query index
    const uint32_t lane_addr = uint32_t(warp * 32) << 16;
    const int nchunk = (nkp + 31) >> 5;
    const int jmax = causal ? i : 0x7fffffff;  // last key this row may attend to
    float v[32];
    mbar_wait(bar_s, 0);
    tc_fence_after();
    if (threadIdx.x == 0) trace(4);
    float mx = -INFINITY;
#pragma unroll 1
    for (int c = 0; c < nchunk; ++c) {
      tc_ld32(tmem_s + lane_addr + c * 32, v);
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int key = c * 32 + j;
        const float cls = pad[key];
        float s = cls == 0.f ? v[j] : cls;
        if (key > jmax) s = -INFINITY;
        mx = fmaxf(mx, s);
      }
    }
    if (threadIdx.x == 0) trace(5);
    // rows past Tq (tile tail) have q = 0 and still see >= 1 key, so mx is finite for every row
    const float kLog2e = 1.4426950408889634f;
    float l = 0.f;
#pragma unroll 1
    for (int c = 0; c < nchunk; ++c) {
      tc_ld32(tmem_s + lane_addr + c * 32, v);
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int key = c * 32 + j;
        const float cls = pad[key];
        float s = cls == 0.f ? v[j] : cls;
        if (key > jmax) s = -INFINITY;
        const float p = fast_exp2((s - mx) * kLog2e);  // subtract first: s and mx may both be -FLT_MAX
        l += p;
        v[j] = p;
      }
      // P row back into the columns this S chunk came from, as the TMEM A operand of O = P V:
      // 16-bit pairs (key 2w, 2w+1) per 32-bit column; hi plane in columns [32c, 32c+16), lo in [32c+16, 32c+32)
      float w[32];
      uint32_t* wp = reinterpret_cast<uint32_t*>(w);
#pragma unroll
      for (int q = 0; q < 16; ++q) split2<FMT>(v[2 * q], v[2 * q + 1], wp[q], wp[16 + q]);
      tc_st32(tmem_s + lane_addr + c * 32, w);
    }
    if (threadIdx.x == 0) trace(6);
    tc_fence_before();
    mbar_arrive(bar_p);

    mbar_wait(bar_o, 0);
    tc_fence_after();
    if (threadIdx.x == 0) trace(7);
    float o[16];
    tc_ld16(tmem_o + lane_addr, o);
    const float inv = 1.0f / l;
    if (i < Tq) {
      const int64_t row = int64_t(b) * Tq + i;
#pragma unroll
      for (int c = 0; c < HD; c += 4) {
        const float4 ov = make_float4(o[c] * inv, o[c + 1] * inv, o[c + 2] * inv, o[c + 3] * inv);
        if (A.out) *reinterpret_cast<float4*>(A.out + row * D + h * HD + c) = ov;
        if (A.out_planes)
          store_planes4(reinterpret_cast<uint16_t*>(A.out_planes), int64_t(P.B) * Tq * D, row * D + h * HD + c, ov, FMT);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) trace(8);
  if (warp == 4) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(tmem_cols) : "memory");
  }
}

}  // namespace

int debug_set_trace_attention(void* dev_buf) {
  long long* p = reinterpret_cast<long long*>(dev_buf);
  SCATT_CUDA(cudaMemcpyToSymbol(g_trace_att, &p, sizeof(p)));
  return SCATT_OK;
}

bool attention_tc_supported(int Tq, int Tk, int hd, const scatt_attention_problem* p, int group) {
  if (hd != HD || Tk > KMAX || Tk < 1) return false;
  for (int i = 0; i < group; ++i)
    if (p[i].additive) return false;  // dense additive masks stay on the fp32 kernel
  return true;
}

int launch_attention_tc(const scatt_attention_problem* p, int group, int B, int Tq, int Tk, int H, int hd, int64_t ldq,
                        int64_t ldk, int64_t ldv, int kind, int fmt, int terms, cudaStream_t s) {
  SCATT_REQUIRE(attention_tc_supported(Tq, Tk, hd, p, group), "attention(tcgen05): unsupported shape / mask");
  SCATT_REQUIRE(terms >= 1 && terms <= 3, "attention(tcgen05): terms must be 1..3");
  SCATT_REQUIRE(ldq % 4 == 0 && ldk % 4 == 0 && ldv % 4 == 0, "attention(tcgen05): row strides must be multiples of 4");
  SCATT_REQUIRE(kind != SCATT_ATTN_CAUSAL || Tq == Tk, "attention(tcgen05): causal needs Tq == Tk");
  SCATT_REQUIRE(int64_t(B) * group <= 65535 && H <= 65535, "attention(tcgen05): grid too large");
  if (B == 0 || Tq == 0) return SCATT_OK;
  AttnTcParams P{};
  for (int i = 0; i < group; ++i) {
    P.p[i] = p[i];
    SCATT_REQUIRE(p[i].q && p[i].k && p[i].v && (p[i].out || p[i].out_planes), "attention(tcgen05): null operand");
  }
  P.B = B, P.Tq = Tq, P.Tk = Tk, P.H = H, P.kind = kind, P.fmt = fmt, P.terms = terms;
  P.ldq = ldq, P.ldk = ldk, P.ldv = ldv;
  static PerDeviceFlag attr_done;
  if (!attr_done.load()) {
    SCATT_CUDA(cudaFuncSetAttribute(stream_attention_tc_kernel<SCATT_PLANE_F16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    int(kSmemBytes)));
    SCATT_CUDA(cudaFuncSetAttribute(stream_attention_tc_kernel<SCATT_PLANE_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    int(kSmemBytes)));
    attr_done.store(true);
  }
  dim3 grid((Tq + QT - 1) / QT, H, B * group);
  if (fmt == SCATT_PLANE_F16)
    (void)launch_kernel(stream_attention_tc_kernel<SCATT_PLANE_F16>, grid, dim3(kThreadsAtt), kSmemBytes, s, P);
  else
    (void)launch_kernel(stream_attention_tc_kernel<SCATT_PLANE_BF16>, grid, dim3(kThreadsAtt), kSmemBytes, s, P);
  return after_launch("stream_attention_tc_kernel");
}

}  // namespace scatt
