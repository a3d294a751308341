// Memory-bound row kernels: K1 front end (region gather + coordinate mapping +
// position embedding + first LayerNorm), position-embed + LayerNorm, the
// row-wise epilogue tail (LayerNorm / residual / activation / clamp), K4
// temporal max-pool and the split-plane packer.
//
// All of them are "one warp owns one row of <= 1024 fp32" kernels: the row
// lives in registers as float4 chunks, reductions are warp shuffles, every
// global access is a coalesced 16-byte vector.
#include <cstdlib>

#include "common.cuh"

namespace scatt {

namespace {

constexpr int kMaxVec = 8;  // float4 chunks per lane -> rows up to 8*32*4 = 1024 columns

struct RowStats {
  float mean, rstd;
};

// LayerNorm statistics of a row held as v[0..nv) float4 per lane (two-pass, fp32).
__device__ __forceinline__ RowStats row_stats(const float4* v, int nvec_row, int lane, int n, float eps) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i)
    if (lane + 32 * i < nvec_row) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  const float mean = warp_sum(s) / float(n);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i)
    if (lane + 32 * i < nvec_row) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      q += (a * a + b * b) + (c * c + d * d);
    }
  const float var = warp_sum(q) / float(n);
  return {mean, rsqrtf(var + eps)};
}

__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

// ---------------------------------------------------------------- row-wise tail
struct RowwiseGroup {  // one entry per problem of a grouped launch (blockIdx.y)
  const float* z[SCATT_MAX_GROUP];
  const float* residual[SCATT_MAX_GROUP];
  const float* g[SCATT_MAX_GROUP];
  const float* b[SCATT_MAX_GROUP];
  float* y[SCATT_MAX_GROUP];
  uint16_t* planes[SCATT_MAX_GROUP];
};

__global__ void __launch_bounds__(256) rowwise_kernel(RowwiseGroup grp, int64_t M, int N, int64_t ldz, int64_t ldres,
                                                      scatt_epilogue ep, int64_t ldy, int fmt) {
  const float* __restrict__ z = grp.z[blockIdx.y];
  const float* __restrict__ residual = grp.residual[blockIdx.y];
  const float* __restrict__ g = grp.g[blockIdx.y];
  const float* __restrict__ b = grp.b[blockIdx.y];
  float* __restrict__ y = grp.y[blockIdx.y];
  uint16_t* __restrict__ planes = grp.planes[blockIdx.y];
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  const int nvec_row = N >> 2;
  for (int64_t row = int64_t(blockIdx.x) * (blockDim.x >> 5) + scatt_warp_idx(); row < M; row += warps) {  // warp-uniform
    float4 v[kMaxVec];
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i)
      if (lane + 32 * i < nvec_row) v[i] = ld4(z + row * ldz + 4 * (lane + 32 * i));
    RowStats st{0.f, 1.f};
    if (ep.layer_norm) st = row_stats(v, nvec_row, lane, N, ep.ln_eps);
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i) {
      const int c = 4 * (lane + 32 * i);
      if (lane + 32 * i < nvec_row) {
        float4 o = v[i];
        if (ep.layer_norm) {
          const float4 gg = ld4(g + c), bb = ld4(b + c);
          o.x = (o.x - st.mean) * st.rstd * gg.x + bb.x;
          o.y = (o.y - st.mean) * st.rstd * gg.y + bb.y;
          o.z = (o.z - st.mean) * st.rstd * gg.z + bb.z;
          o.w = (o.w - st.mean) * st.rstd * gg.w + bb.w;
        }
        if (ep.residual_mode == SCATT_RES_AFTER_LN) {
          const float4 r = ld4(residual + row * ldres + c);
          o.x += r.x, o.y += r.y, o.z += r.z, o.w += r.w;
        }
        o.x = apply_act(o.x, ep.act_post), o.y = apply_act(o.y, ep.act_post);
        o.z = apply_act(o.z, ep.act_post), o.w = apply_act(o.w, ep.act_post);
        if (ep.clamp > 0.f) {
          o.x = fminf(fmaxf(o.x, -ep.clamp), ep.clamp), o.y = fminf(fmaxf(o.y, -ep.clamp), ep.clamp);
          o.z = fminf(fmaxf(o.z, -ep.clamp), ep.clamp), o.w = fminf(fmaxf(o.w, -ep.clamp), ep.clamp);
        }
        if (y) st4(y + row * ldy + c, o);
        if (planes) store_planes4(planes, M * int64_t(N), row * N + c, o, fmt);
      }
    }
  }
}

// ---------------------------------------------------------------- split-K reduce + epilogue
// z = sum of `nsplit` fp32 partial products [nsplit][M][N] (slot order: deterministic) - then exactly the chain of the
// GEMM epilogue (gemm_tc.cu chunk_pre / LayerNorm pass): + bias, column scaling, act_pre, residual before LayerNorm,
// LayerNorm, residual after it, act_post, clamp; fp32 and / or split-plane outputs.  One warp per row, N <= 1024.
struct SplitKArgs {
  const float* partials;
  const float* bias;
  const float* residual;
  const uint16_t* res_planes;
  const float* g;
  const float* b;
  float* y;
  uint16_t* planes;
};

// One CTA per row, one float4 of the row per thread (N / 4 <= 256 threads): every load of a thread - the nsplit partials,
// bias, residual - is independent and in flight at once; with a warp per row the 4 x 8 dependent L2 round trips took
// 17 us per launch.  LayerNorm statistics: two-pass (mean, then centred squares) through warp shuffles + shared memory.
__device__ __forceinline__ float block_sum(float v, float* red, int nwarps) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = scatt_warp_idx();
  __syncthreads();  // red may still be read from the previous reduction
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = 0.f;
  for (int w = 0; w < nwarps; ++w) t += red[w];  // same order in every thread
  return t;
}

__global__ void __launch_bounds__(256) rowwise_splitk_kernel(SplitKArgs a, int nsplit, int64_t M, int N, int64_t ldres, scatt_epilogue ep,
                                                             int64_t ldy, int fmt) {
  __shared__ float red[8];
  pdl_launch_dependents();
  pdl_wait();
  const int64_t mn = M * int64_t(N);
  const int64_t row = blockIdx.x;
  const int c = 4 * threadIdx.x;
  const bool in = c < N;
  const int nwarps = (blockDim.x + 31) >> 5;
  auto residual_at = [&]() -> float4 {
    if (a.residual) return ld4(a.residual + row * ldres + c);
    const uint2 h = *reinterpret_cast<const uint2*>(a.res_planes + row * N + c);
    const uint2 l = *reinterpret_cast<const uint2*>(a.res_planes + mn + row * N + c);
    float2 h0, h1, l0, l1;
    if (fmt == SCATT_PLANE_F16) {
      h0 = __half22float2(*reinterpret_cast<const __half2*>(&h.x)), h1 = __half22float2(*reinterpret_cast<const __half2*>(&h.y));
      l0 = __half22float2(*reinterpret_cast<const __half2*>(&l.x)), l1 = __half22float2(*reinterpret_cast<const __half2*>(&l.y));
    } else {
      h0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&h.x)), h1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&h.y));
      l0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&l.x)), l1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&l.y));
    }
    return make_float4(h0.x + l0.x, h0.y + l0.y, h1.x + l1.x, h1.y + l1.y);
  };
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f), res = acc, gg = acc, bb = acc;
  if (in) {
    float4 part[SCATT_MAX_GROUP];
#pragma unroll
    for (int sp = 0; sp < SCATT_MAX_GROUP; ++sp)
      part[sp] = sp < nsplit ? ld4(a.partials + sp * mn + row * N + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 bias = a.bias ? ld4(a.bias + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    if (ep.residual_mode != SCATT_RES_NONE) res = residual_at();
    if (ep.layer_norm) gg = ld4(a.g + c), bb = ld4(a.b + c);
    acc = part[0];
#pragma unroll
    for (int sp = 1; sp < SCATT_MAX_GROUP; ++sp)  // slot order; absent slots add +0
      acc.x += part[sp].x, acc.y += part[sp].y, acc.z += part[sp].z, acc.w += part[sp].w;
    acc.x += bias.x, acc.y += bias.y, acc.z += bias.z, acc.w += bias.w;
    if (c < ep.scale_cols) acc.x *= ep.scale, acc.y *= ep.scale, acc.z *= ep.scale, acc.w *= ep.scale;
    if (ep.act_pre != SCATT_ACT_NONE) {
      acc.x = apply_act(acc.x, ep.act_pre), acc.y = apply_act(acc.y, ep.act_pre);
      acc.z = apply_act(acc.z, ep.act_pre), acc.w = apply_act(acc.w, ep.act_pre);
    }
    if (ep.residual_mode == SCATT_RES_BEFORE_LN || (ep.residual_mode == SCATT_RES_AFTER_LN && !ep.layer_norm))
      acc.x += res.x, acc.y += res.y, acc.z += res.z, acc.w += res.w;
  }
  float4 o = acc;
  if (ep.layer_norm) {
    const float mean = block_sum(in ? (acc.x + acc.y) + (acc.z + acc.w) : 0.f, red, nwarps) / float(N);
    const float dx = acc.x - mean, dy = acc.y - mean, dz = acc.z - mean, dw = acc.w - mean;
    const float var = block_sum(in ? (dx * dx + dy * dy) + (dz * dz + dw * dw) : 0.f, red, nwarps) / float(N);
    const float rstd = rsqrtf(var + ep.ln_eps);
    o.x = dx * rstd * gg.x + bb.x, o.y = dy * rstd * gg.y + bb.y;
    o.z = dz * rstd * gg.z + bb.z, o.w = dw * rstd * gg.w + bb.w;
    if (ep.residual_mode == SCATT_RES_AFTER_LN) o.x += res.x, o.y += res.y, o.z += res.z, o.w += res.w;
  }
  if (!in) return;
  o.x = apply_act(o.x, ep.act_post), o.y = apply_act(o.y, ep.act_post);
  o.z = apply_act(o.z, ep.act_post), o.w = apply_act(o.w, ep.act_post);
  if (ep.clamp > 0.f) {
    o.x = fminf(fmaxf(o.x, -ep.clamp), ep.clamp), o.y = fminf(fmaxf(o.y, -ep.clamp), ep.clamp);
    o.z = fminf(fmaxf(o.z, -ep.clamp), ep.clamp), o.w = fminf(fmaxf(o.w, -ep.clamp), ep.clamp);
  }
  if (a.y) st4(a.y + row * ldy + c, o);
  if (a.planes) store_planes4(a.planes, mn, row * N + c, o, fmt);
}

// ---------------------------------------------------------------- pos-embed + LayerNorm
__global__ void __launch_bounds__(256) posembed_ln_kernel(const float* __restrict__ x, const float* __restrict__ table,
                                                          const float* __restrict__ g, const float* __restrict__ b,
                                                          float* __restrict__ out, uint16_t* __restrict__ planes, int B,
                                                          int T, int D, int fmt) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int64_t M = int64_t(B) * T;
  const int64_t warps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  const int nvec_row = D >> 2;
  for (int64_t row = int64_t(blockIdx.x) * (blockDim.x >> 5) + scatt_warp_idx(); row < M; row += warps) {  // warp-uniform
    const int t = int(row % T);
    float4 v[kMaxVec];
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i)
      if (lane + 32 * i < nvec_row) {
        const int c = 4 * (lane + 32 * i);
        const float4 a = ld4(x + row * D + c), p = ld4(table + int64_t(t + 2) * D + c);
        v[i] = make_float4(a.x + p.x, a.y + p.y, a.z + p.z, a.w + p.w);
      }
    const RowStats st = row_stats(v, nvec_row, lane, D, 1e-5f);
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i)
      if (lane + 32 * i < nvec_row) {
        const int c = 4 * (lane + 32 * i);
        const float4 gg = ld4(g + c), bb = ld4(b + c);
        float4 o;
        o.x = (v[i].x - st.mean) * st.rstd * gg.x + bb.x;
        o.y = (v[i].y - st.mean) * st.rstd * gg.y + bb.y;
        o.z = (v[i].z - st.mean) * st.rstd * gg.z + bb.z;
        o.w = (v[i].w - st.mean) * st.rstd * gg.w + bb.w;
        if (out) st4(out + row * D + c, o);
        if (planes) store_planes4(planes, M * int64_t(D), row * D + c, o, fmt);
      }
  }
}

// ---------------------------------------------------------------- K1 front end
// One CTA = one anatomical stream; W^T of both coordinate mappings staged in
// shared memory once, then each warp walks groups of R frames: lane j loads joint j's
// (x, y) pair straight from keypoints[b,t,idx[j],:] (the region gather),
// broadcasts it by shuffle, and every lane accumulates its 8 output channels
// of R frames for both branches (packed fp32 FMAs; a weight vector read from
// shared memory serves R frames).  D is fixed at 256 (8 channels per lane).
//
// Outputs leave through the bulk-copy engine: a warp writes two finished rows
// (fp32 + hi plane + lo plane = 4 KB) into its shared-memory staging buffer and
// one lane hands them over with cp.async.bulk (shared -> global, contiguous
// because consecutive rows are).  With st.global the kernel sat at the SM's
// LSU store path (~11 B/clk/SM of the ~16 measured, 50 % of the HBM rate); the
// bulk engine moves ~27 B/clk/SM.
struct FrontendParams {
  scatt_frontend_stream s[SCATT_MAX_GROUP];
  int32_t cta_begin[SCATT_MAX_GROUP + 1];  // CTAs [cta_begin[g], cta_begin[g + 1]) work on stream g
};

constexpr int kFeWarps = 8;
constexpr int kFeStageRows = 2;                              // rows per bulk hand-over
constexpr int kFeBufBytes = kFeStageRows * (1024 + 2 * 512); // fp32 rows | hi rows | lo rows
constexpr int kFeStageBytes = kFeWarps * 2 * kFeBufBytes;    // two buffers per warp

__device__ __forceinline__ void bulk_store(void* gdst, uint32_t ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(ssrc), "r"(bytes) : "memory");
}

template <int R>  // frames per warp pass: every staged weight vector is used for R frames
__global__ void __launch_bounds__(32 * kFeWarps, 2) frontend_kernel(const float* __restrict__ kp, int B, int T, int K,
                                                                    const __grid_constant__ FrontendParams prm, int fmt, int wt_stride) {
  constexpr int D = 256;
  static_assert(R == 1 || R % kFeStageRows == 0, "R is one row or whole staging buffers");
  extern __shared__ __align__(128) float smem[];
  int g = 0;
  while (g + 1 < SCATT_MAX_GROUP && int(blockIdx.x) >= prm.cta_begin[g + 1]) ++g;
  const scatt_frontend_stream& S = prm.s[g];
  const int cta = blockIdx.x - prm.cta_begin[g], nctas = prm.cta_begin[g + 1] - prm.cta_begin[g];
  const int nj = S.n_joints;
  float* wt[2] = {smem, smem + wt_stride};  // [nj][D] transposed mapping weights per branch
  for (int br = 0; br < 2; ++br)  // the host passes W^T [nj][D]: a straight, coalesced copy
    for (int i = threadIdx.x * 4; i < nj * D; i += blockDim.x * 4)
      *reinterpret_cast<float4*>(wt[br] + i) = *reinterpret_cast<const float4*>(S.map_wt[br] + i);
  const int lane = threadIdx.x & 31, warp = scatt_warp_idx();
  uint8_t* stage = reinterpret_cast<uint8_t*>(smem + 2 * wt_stride) + warp * 2 * kFeBufBytes;
  const uint32_t stage_addr = uint32_t(__cvta_generic_to_shared(stage));
  const int my_joint = lane < nj ? S.joint_idx[lane] : 0;
  const int c0 = 4 * lane, c1 = 128 + 4 * lane;
  // per-lane column parameters (static weights, like the staging above: before the dependency wait)
  float4 bias[2][2], gam[2][2], bet[2][2];
#pragma unroll
  for (int br = 0; br < 2; ++br) {
    bias[br][0] = ld4(S.map_b[br] + c0), bias[br][1] = ld4(S.map_b[br] + c1);
    gam[br][0] = ld4(S.ln_g[br] + c0), gam[br][1] = ld4(S.ln_g[br] + c1);
    bet[br][0] = ld4(S.ln_b[br] + c0), bet[br][1] = ld4(S.ln_b[br] + c1);
  }
  pdl_launch_dependents();
  pdl_wait();  // everything above is static; keypoints / outputs below follow stream order
  __syncthreads();

  const int64_t M = int64_t(B) * T;
  const int64_t ngroups = (M + R - 1) / R;
  uint32_t handed = 0;  // bulk groups committed by this warp (selects the staging buffer)
  for (int64_t grp = int64_t(cta) * kFeWarps + warp; grp < ngroups; grp += int64_t(nctas) * kFeWarps) {
    const int64_t row0 = grp * R;
    const int t0 = int(row0 % T);
    float2 xy[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      xy[r] = make_float2(0.f, 0.f);
      if (lane < nj && row0 + r < M) {
        xy[r] = *reinterpret_cast<const float2*>(kp + ((row0 + r) * K + my_joint) * 2);
        if (S.gathered) *reinterpret_cast<float2*>(S.gathered + ((row0 + r) * nj + lane) * 2) = xy[r];
      }
    }
#pragma unroll
    for (int br = 0; br < 2; ++br) {
      float mine[R];
#pragma unroll
      for (int r = 0; r < R; ++r) mine[r] = S.coord[br] == 0 ? xy[r].x : xy[r].y;
      // nn.Linear accumulates bias + sum_j; keep the dot product in j order.  Packed fp32 FMAs (FFMA2:
      // two fused multiply-adds per instruction, each rounded like fmaf).
      float2 acc[R][4];
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[r][q] = make_float2(0.f, 0.f);
      for (int j = 0; j < nj; ++j) {
        const float4 w0 = ld4(wt[br] + j * D + c0), w1 = ld4(wt[br] + j * D + c1);
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const float c = __shfl_sync(0xffffffffu, mine[r], j);
          const float2 cc = make_float2(c, c);
          acc[r][0] = __ffma2_rn(cc, make_float2(w0.x, w0.y), acc[r][0]);
          acc[r][1] = __ffma2_rn(cc, make_float2(w0.z, w0.w), acc[r][1]);
          acc[r][2] = __ffma2_rn(cc, make_float2(w1.x, w1.y), acc[r][2]);
          acc[r][3] = __ffma2_rn(cc, make_float2(w1.z, w1.w), acc[r][3]);
        }
      }
      float* out = S.out[br];
      uint16_t* planes = reinterpret_cast<uint16_t*>(S.out_planes[br]);
#pragma unroll
      for (int r = 0; r < R; ++r) {  // rows past M are computed on zeros and never handed over (warp-uniform flow)
        const int rs = r % kFeStageRows;  // row slot inside the staging buffer
        uint8_t* buf = stage + (handed & 1u) * kFeBufBytes;
        if (rs == 0 && handed >= 2) {  // the buffer's previous contents must have been read out
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          __syncwarp();
        }
        int t = t0 + r;
        while (t >= T) t -= T;
        const float4 a0 = bias[br][0], a1 = bias[br][1];
        const float4 p0 = ld4(S.pos[br] + int64_t(t + 2) * D + c0), p1 = ld4(S.pos[br] + int64_t(t + 2) * D + c1);
        // (dot + bias) + position row, as the reference adds them
        const float2 e0 = __fadd2_rn(__fadd2_rn(acc[r][0], make_float2(a0.x, a0.y)), make_float2(p0.x, p0.y));
        const float2 e1 = __fadd2_rn(__fadd2_rn(acc[r][1], make_float2(a0.z, a0.w)), make_float2(p0.z, p0.w));
        const float2 e2 = __fadd2_rn(__fadd2_rn(acc[r][2], make_float2(a1.x, a1.y)), make_float2(p1.x, p1.y));
        const float2 e3 = __fadd2_rn(__fadd2_rn(acc[r][3], make_float2(a1.z, a1.w)), make_float2(p1.z, p1.w));
        float4 v[kMaxVec];
        v[0] = make_float4(e0.x, e0.y, e1.x, e1.y);
        v[1] = make_float4(e2.x, e2.y, e3.x, e3.y);
        const RowStats st = row_stats(v, D / 4, lane, D, 1e-5f);
        const float2 nm = make_float2(-st.mean, -st.mean), rsd = make_float2(st.rstd, st.rstd);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int c = i == 0 ? c0 : c1;
          const float4 gg = gam[br][i], bb = bet[br][i];
          const float2 lo2 = __ffma2_rn(__fmul2_rn(__fadd2_rn(make_float2(v[i].x, v[i].y), nm), rsd), make_float2(gg.x, gg.y),
                                        make_float2(bb.x, bb.y));
          const float2 hi2 = __ffma2_rn(__fmul2_rn(__fadd2_rn(make_float2(v[i].z, v[i].w), nm), rsd), make_float2(gg.z, gg.w),
                                        make_float2(bb.z, bb.w));
          const float4 o = make_float4(lo2.x, lo2.y, hi2.x, hi2.y);
          if (out) *reinterpret_cast<float4*>(buf + rs * 1024 + c * 4) = o;
          if (planes) {
            uint2 ph, pl;
            split_pair_rt(o.x, o.y, fmt, ph.x, pl.x);
            split_pair_rt(o.z, o.w, fmt, ph.y, pl.y);
            *reinterpret_cast<uint2*>(buf + kFeStageRows * 1024 + rs * 512 + c * 2) = ph;
            *reinterpret_cast<uint2*>(buf + kFeStageRows * 1536 + rs * 512 + c * 2) = pl;
          }
        }
        if (rs == kFeStageRows - 1 || r == R - 1) {  // hand the buffer's rows to the bulk-copy engine
          const int64_t first = row0 + r - rs;
          const int64_t left = M - first;
          const uint32_t rows = uint32_t(left < int64_t(rs + 1) ? (left < 0 ? 0 : left) : rs + 1);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0 && rows > 0) {
            const uint32_t src = stage_addr + (handed & 1u) * kFeBufBytes;
            if (out) bulk_store(out + first * D, src, rows * 1024u);
            if (planes) {
              bulk_store(planes + first * D, src + kFeStageRows * 1024, rows * 512u);
              bulk_store(planes + M * int64_t(D) + first * D, src + kFeStageRows * 1536, rows * 512u);
            }
          }
          if (lane == 0) asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          ++handed;
        }
      }
    }
  }
  // the staging buffers must have been read out before the CTA exits; the global writes drain behind it
  if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  __syncwarp();
}

// ---------------------------------------------------------------- K4 temporal max-pool
struct PoolGroup {  // one entry per problem of a grouped launch (blockIdx.y)
  const float* x[SCATT_MAX_GROUP];
  float* y[SCATT_MAX_GROUP];
  uint16_t* planes[SCATT_MAX_GROUP];
};

__global__ void __launch_bounds__(256) pool_pairs_kernel(PoolGroup grp, int B, int T, int C, int fmt) {
  const float* __restrict__ x = grp.x[blockIdx.y];
  float* __restrict__ y = grp.y[blockIdx.y];
  uint16_t* __restrict__ planes = grp.planes[blockIdx.y];
  pdl_launch_dependents();
  pdl_wait();
  const int To = T >> 1, cv = C >> 2;
  const int64_t total = int64_t(B) * To * cv;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const int c = int(i % cv) * 4;
    const int64_t r = i / cv;  // output row
    const int64_t b = r / To, to = r % To;
    const float* src = x + ((b * T + 2 * to) * C) + c;
    const float4 u = ld4(src), w = ld4(src + C);
    const float4 o = make_float4(fmaxf(u.x, w.x), fmaxf(u.y, w.y), fmaxf(u.z, w.z), fmaxf(u.w, w.w));
    if (y) st4(y + r * C + c, o);
    if (planes) store_planes4(planes, int64_t(B) * To * C, r * C + c, o, fmt);
  }
}

// ---------------------------------------------------------------- split-plane packer
__global__ void __launch_bounds__(256) split_planes_kernel(const float* __restrict__ x, int64_t rows, int64_t cols,
                                                           int64_t ldx, float scale, uint16_t* __restrict__ planes,
                                                           int fmt) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t cv = cols >> 2, total = rows * cv;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t r = i / cv, c = (i % cv) * 4;
    float4 v = ld4(x + r * ldx + c);
    v.x *= scale, v.y *= scale, v.z *= scale, v.w *= scale;
    store_planes4(planes, rows * cols, r * cols + c, v, fmt);
  }
}

inline int grid_for(int64_t work_items, int per_block, int cap = 148 * 8) {
  int64_t g = (work_items + per_block - 1) / per_block;
  if (g < 1) g = 1;
  if (g > cap) g = cap;
  return int(g);
}

}  // namespace

int launch_rowwise(const float* z, int64_t M, int N, int64_t ldz, const float* residual, int64_t ldres, const float* g,
                   const float* b, const scatt_epilogue& ep, float* y, int64_t ldy, void* planes, int fmt,
                   cudaStream_t s) {
  SCATT_REQUIRE(N % 4 == 0 && N <= 1024 && N > 0, "rowwise: N=%d must be a multiple of 4 and <= 1024", N);
  SCATT_REQUIRE(ldz % 4 == 0 && (!y || ldy % 4 == 0), "rowwise: row strides must be multiples of 4");
  SCATT_REQUIRE(!ep.layer_norm || (g && b), "rowwise: LayerNorm needs gamma and beta");
  SCATT_REQUIRE(ep.residual_mode != SCATT_RES_AFTER_LN || (residual && ldres % 4 == 0), "rowwise: residual missing");
  if (M == 0) return SCATT_OK;
  RowwiseGroup grp{};
  grp.z[0] = z, grp.residual[0] = residual, grp.g[0] = g, grp.b[0] = b, grp.y[0] = y;
  grp.planes[0] = reinterpret_cast<uint16_t*>(planes);
  (void)launch_kernel(rowwise_kernel, dim3(grid_for(M, 8), 1), dim3(256), 0, s, grp, M, N, ldz, ldres, ep, ldy, fmt);
  return after_launch("rowwise_kernel");
}

// LayerNorm tail of a grouped linear launch, in place on y: one launch for all problems.
int launch_rowwise_linear_tail(const scatt_linear_problem* p, int group, int64_t M, int N, int64_t ldres, int64_t ldy,
                               const scatt_epilogue& ep, int fmt, cudaStream_t s) {
  SCATT_REQUIRE(N % 4 == 0 && N <= 1024 && N > 0, "rowwise: N=%d must be a multiple of 4 and <= 1024", N);
  SCATT_REQUIRE(ldy % 4 == 0 && ldres % 4 == 0, "rowwise: row strides must be multiples of 4");
  if (M == 0) return SCATT_OK;
  RowwiseGroup grp{};
  for (int i = 0; i < group; ++i) {
    SCATT_REQUIRE(p[i].y && (!ep.layer_norm || (p[i].ln_g && p[i].ln_b)), "rowwise: LayerNorm needs y, gamma and beta");
    SCATT_REQUIRE(ep.residual_mode != SCATT_RES_AFTER_LN || p[i].residual, "rowwise: residual missing");
    grp.z[i] = p[i].y, grp.residual[i] = p[i].residual, grp.g[i] = p[i].ln_g, grp.b[i] = p[i].ln_b, grp.y[i] = p[i].y;
    grp.planes[i] = reinterpret_cast<uint16_t*>(p[i].y_planes);
  }
  (void)launch_kernel(rowwise_kernel, dim3(grid_for(M, 8, 148 * 8 / group), group), dim3(256), 0, s, grp, M, N, ldy, ldres, ep, ldy, fmt);
  return after_launch("rowwise_kernel");
}

int launch_rowwise_splitk(const float* partials, int nsplit, const scatt_linear_problem& p, int64_t M, int N, int64_t ldres,
                          int64_t ldy, const scatt_epilogue& ep, int fmt, cudaStream_t s) {
  SCATT_REQUIRE(N % 4 == 0 && N <= 1024 && N > 0 && nsplit >= 1, "rowwise(split-K): N=%d must be a multiple of 4 and <= 1024", N);
  SCATT_REQUIRE(ldy % 4 == 0 && ldres % 4 == 0 && ep.scale_cols % 4 == 0, "rowwise(split-K): row strides must be multiples of 4");
  if (M == 0) return SCATT_OK;
  SplitKArgs a{partials, p.bias, p.residual, reinterpret_cast<const uint16_t*>(p.residual_planes), p.ln_g, p.ln_b, p.y,
               reinterpret_cast<uint16_t*>(p.y_planes)};
  SCATT_REQUIRE(nsplit <= SCATT_MAX_GROUP && M <= 0x7fffffff, "rowwise(split-K): at most %d slices", SCATT_MAX_GROUP);
  const int threads = ((N / 4) + 31) & ~31;  // one float4 per thread
  (void)launch_kernel(rowwise_splitk_kernel, dim3(unsigned(M)), dim3(threads), 0, s, a, nsplit, M, N, ldres, ep, ldy, fmt);
  return after_launch("rowwise_splitk_kernel");
}

int launch_posembed_ln(const float* x, const float* table, const float* g, const float* b, float* out, void* planes,
                       int B, int T, int D, int fmt, cudaStream_t s) {
  SCATT_REQUIRE(D % 4 == 0 && D <= 1024, "posembed_layernorm: D=%d unsupported", D);
  if (int64_t(B) * T == 0) return SCATT_OK;
  (void)launch_kernel(posembed_ln_kernel, dim3(grid_for(int64_t(B) * T, 8)), dim3(256), 0, s, x, table, g, b, out, reinterpret_cast<uint16_t*>(planes),
                                                                 B, T, D, fmt);
  return after_launch("posembed_ln_kernel");
}

int launch_frontend(const float* kp, int B, int T, int K, int D, const scatt_frontend_stream* streams, int n, int max_pos,
                    int fmt, cudaStream_t s) {
  SCATT_REQUIRE(D == 256, "frontend: d_model must be 256 (got %d)", D);
  SCATT_REQUIRE(n >= 1 && n <= SCATT_MAX_GROUP, "frontend: 1..%d streams", SCATT_MAX_GROUP);
  SCATT_REQUIRE(T <= max_pos, "frontend: T=%d exceeds max_position_embeddings=%d", T, max_pos);
  FrontendParams prm{};
  int max_nj = 0;
  for (int i = 0; i < n; ++i) {
    SCATT_REQUIRE(streams[i].n_joints >= 1 && streams[i].n_joints <= 32, "frontend: 1..32 joints per stream");
    prm.s[i] = streams[i];
    if (streams[i].n_joints > max_nj) max_nj = streams[i].n_joints;
  }
  if (int64_t(B) * T == 0) return SCATT_OK;
  {  // planes-only output of large batches: the mapping runs on tcgen05 (frontend_tc.cu).  SCATT_FRONTEND_TC=0 keeps the
     // CUDA-core kernel, =2 takes the tensor-core kernel from 128 frames up (tests, sweeps); read per launch
    const char* e = std::getenv("SCATT_FRONTEND_TC");
    const int mode = e ? std::atoi(e) : 1;
    if (mode != 0) {
      const int rc = launch_frontend_tc(kp, B, T, K, streams, n, max_pos, fmt, mode == 2, s);
      if (rc <= 0) return rc;  // launched (or failed); > 0: outside that kernel's envelope
    }
  }
  const int wt_stride = max_nj * D;
  const size_t smem = size_t(2) * wt_stride * sizeof(float) + kFeStageBytes;
  static PerDeviceFlag attr_done;
  if (!attr_done.load()) {
    const int max_smem = 2 * 32 * 256 * 4 + kFeStageBytes;
    SCATT_CUDA(cudaFuncSetAttribute(frontend_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    SCATT_CUDA(cudaFuncSetAttribute(frontend_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    attr_done.store(true);
  }
  // Tiny batches: one frame per warp pass; otherwise four (every weight vector read from shared memory serves
  // four frames).  CTAs per stream: one group per warp while everything fits the 2 x 148 resident CTAs, else
  // the resident CTAs are shared out by the streams' cost (joints) and walk their groups grid-stride.
  const int64_t M = int64_t(B) * T;
  const int R = M * n <= int64_t(148) * kFeWarps * 2 ? 1 : 4;
  const int64_t ngroups = (M + R - 1) / R;
  const int64_t want = (ngroups + kFeWarps - 1) / kFeWarps;  // CTAs per stream for one group per warp
  const int slots = 148 * 2;
  prm.cta_begin[0] = 0;
  if (want * n <= slots) {
    for (int i = 0; i < n; ++i) prm.cta_begin[i + 1] = prm.cta_begin[i] + int(want);
  } else {
    int64_t cost[SCATT_MAX_GROUP], total = 0;
    for (int i = 0; i < n; ++i) total += cost[i] = 11 * int64_t(streams[i].n_joints) + 230;  // instructions per frame, roughly
    for (int i = 0; i < n; ++i) {
      int share = int((slots * cost[i]) / total);
      if (share < 1) share = 1;
      prm.cta_begin[i + 1] = prm.cta_begin[i] + share;
    }
  }
  for (int i = n; i < SCATT_MAX_GROUP; ++i) prm.cta_begin[i + 1] = prm.cta_begin[n];
  dim3 grid(prm.cta_begin[n]);
  if (R == 1) (void)launch_kernel(frontend_kernel<1>, grid, dim3(32 * kFeWarps), smem, s, kp, B, T, K, prm, fmt, wt_stride);
  else (void)launch_kernel(frontend_kernel<4>, grid, dim3(32 * kFeWarps), smem, s, kp, B, T, K, prm, fmt, wt_stride);
  const int rc = after_launch("frontend_kernel");
  set_last_kernel("frontend_kernel<%d>", R);
  return rc;
}

int launch_pool_pairs(const float* x, int B, int T, int C, float* y, void* planes, int fmt, cudaStream_t s) {
  SCATT_REQUIRE(C % 4 == 0, "pool_pairs: C must be a multiple of 4");
  SCATT_REQUIRE(T >= 2, "pool_pairs: T=%d gives an empty output (the reference raises too)", T);
  const float* xs[1] = {x};
  float* ys[1] = {y};
  void* ps[1] = {planes};
  return launch_pool_pairs_group(xs, ys, ps, 1, B, T, C, fmt, s);
}

int launch_pool_pairs_group(const float* const* xs, float* const* ys, void* const* planes, int group, int B, int T, int C,
                            int fmt, cudaStream_t s) {
  SCATT_REQUIRE(group >= 1 && group <= SCATT_MAX_GROUP, "pool_pairs: group 1..%d", SCATT_MAX_GROUP);
  SCATT_REQUIRE(C % 4 == 0, "pool_pairs: C must be a multiple of 4");
  SCATT_REQUIRE(T >= 2, "pool_pairs: T=%d gives an empty output (the reference raises too)", T);
  const int64_t total = int64_t(B) * (T / 2) * (C / 4);
  if (total == 0) return SCATT_OK;
  PoolGroup grp{};
  for (int i = 0; i < group; ++i) {
    SCATT_REQUIRE(xs[i] && (ys[i] || (planes && planes[i])), "pool_pairs: null pointer in problem %d", i);
    grp.x[i] = xs[i], grp.y[i] = ys[i], grp.planes[i] = planes ? reinterpret_cast<uint16_t*>(planes[i]) : nullptr;
  }
  (void)launch_kernel(pool_pairs_kernel, dim3(grid_for(total, 256, 148 * 8 / group), group), dim3(256), 0, s, grp, B, T, C, fmt);
  return after_launch("pool_pairs_kernel");
}

int launch_split_planes(const float* x, int64_t rows, int64_t cols, int64_t ldx, float scale, void* planes, int fmt,
                        cudaStream_t s) {
  SCATT_REQUIRE(cols % 4 == 0 && ldx % 4 == 0, "split_planes: cols and ldx must be multiples of 4");
  if (rows * cols == 0) return SCATT_OK;
  (void)launch_kernel(split_planes_kernel, dim3(grid_for(rows * (cols / 4), 256)), dim3(256), 0, s, x, rows, cols, ldx, scale,
                                                                       reinterpret_cast<uint16_t*>(planes), fmt);
  return after_launch("split_planes_kernel");
}

}  // namespace scatt
