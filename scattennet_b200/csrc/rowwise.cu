// Memory-bound row kernels: K1 front end (region gather + coordinate mapping +
// position embedding + first LayerNorm), position-embed + LayerNorm, the
// row-wise epilogue tail (LayerNorm / residual / activation / clamp), K4
// temporal max-pool and the split-plane packer.
//
// All of them are "one warp owns one row of <= 1024 fp32" kernels: the row
// lives in registers as float4 chunks, reductions are warp shuffles, every
// global access is a coalesced 16-byte vector.
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
  for (int64_t row = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; row < M; row += warps) {
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
  for (int64_t row = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; row < M; row += warps) {
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
// shared memory once, then each warp walks frames: lane j loads joint j's
// (x, y) pair straight from keypoints[b,t,idx[j],:] (the region gather),
// broadcasts it by shuffle, and every lane accumulates its 8 output channels
// for both branches.  D is fixed at 256 (8 channels per lane as 2 float4).
struct FrontendParams {
  scatt_frontend_stream s[SCATT_MAX_GROUP];
};

__global__ void __launch_bounds__(256) frontend_kernel(const float* __restrict__ kp, int B, int T, int K,
                                                       FrontendParams prm, int fmt, int wt_stride) {
  constexpr int D = 256;
  extern __shared__ float smem[];
  const scatt_frontend_stream& S = prm.s[blockIdx.y];
  const int nj = S.n_joints;
  float* wt[2] = {smem, smem + wt_stride};  // [nj][D] transposed mapping weights per branch
  for (int br = 0; br < 2; ++br)  // the host passes W^T [nj][D]: a straight, coalesced copy
    for (int i = threadIdx.x * 4; i < nj * D; i += blockDim.x * 4)
      *reinterpret_cast<float4*>(wt[br] + i) = *reinterpret_cast<const float4*>(S.map_wt[br] + i);
  pdl_launch_dependents();
  pdl_wait();  // weights above are static; keypoints / outputs below follow stream order
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const int64_t M = int64_t(B) * T;
  const int my_joint = lane < nj ? S.joint_idx[lane] : 0;
  const int c0 = 4 * lane, c1 = 128 + 4 * lane;

  for (int64_t row = int64_t(blockIdx.x) * nwarp + warp; row < M; row += int64_t(gridDim.x) * nwarp) {
    const int t = int(row % T);
    float2 xy = make_float2(0.f, 0.f);
    if (lane < nj) {
      xy = *reinterpret_cast<const float2*>(kp + (row * K + my_joint) * 2);
      if (S.gathered) *reinterpret_cast<float2*>(S.gathered + (row * nj + lane) * 2) = xy;
    }
#pragma unroll
    for (int br = 0; br < 2; ++br) {
      const float mine = S.coord[br] == 0 ? xy.x : xy.y;
      float4 a0 = ld4(S.map_b[br] + c0), a1 = ld4(S.map_b[br] + c1);
      // nn.Linear accumulates bias + sum_j; keep the dot product in j order.
      float4 d0 = make_float4(0.f, 0.f, 0.f, 0.f), d1 = d0;
      for (int j = 0; j < nj; ++j) {
        const float c = __shfl_sync(0xffffffffu, mine, j);
        const float4 w0 = ld4(wt[br] + j * D + c0), w1 = ld4(wt[br] + j * D + c1);
        d0.x = fmaf(c, w0.x, d0.x), d0.y = fmaf(c, w0.y, d0.y), d0.z = fmaf(c, w0.z, d0.z), d0.w = fmaf(c, w0.w, d0.w);
        d1.x = fmaf(c, w1.x, d1.x), d1.y = fmaf(c, w1.y, d1.y), d1.z = fmaf(c, w1.z, d1.z), d1.w = fmaf(c, w1.w, d1.w);
      }
      const float4 p0 = ld4(S.pos[br] + int64_t(t + 2) * D + c0), p1 = ld4(S.pos[br] + int64_t(t + 2) * D + c1);
      float4 v[kMaxVec];
      v[0] = make_float4((d0.x + a0.x) + p0.x, (d0.y + a0.y) + p0.y, (d0.z + a0.z) + p0.z, (d0.w + a0.w) + p0.w);
      v[1] = make_float4((d1.x + a1.x) + p1.x, (d1.y + a1.y) + p1.y, (d1.z + a1.z) + p1.z, (d1.w + a1.w) + p1.w);
      const RowStats st = row_stats(v, D / 4, lane, D, 1e-5f);
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int c = i == 0 ? c0 : c1;
        const float4 gg = ld4(S.ln_g[br] + c), bb = ld4(S.ln_b[br] + c);
        float4 o;
        o.x = (v[i].x - st.mean) * st.rstd * gg.x + bb.x;
        o.y = (v[i].y - st.mean) * st.rstd * gg.y + bb.y;
        o.z = (v[i].z - st.mean) * st.rstd * gg.z + bb.z;
        o.w = (v[i].w - st.mean) * st.rstd * gg.w + bb.w;
        if (S.out[br]) st4(S.out[br] + row * D + c, o);
        if (S.out_planes[br])
          store_planes4(reinterpret_cast<uint16_t*>(S.out_planes[br]), M * int64_t(D), row * D + c, o, fmt);
      }
    }
  }
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
  const int wt_stride = max_nj * D;
  const size_t smem = size_t(2) * wt_stride * sizeof(float);
  static std::atomic<bool> attr_done{false};
  if (!attr_done.load()) {
    SCATT_CUDA(cudaFuncSetAttribute(frontend_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 32 * 256 * 4));
    attr_done.store(true);
  }
  dim3 grid(grid_for(int64_t(B) * T, 8 * 4, 148 * 2), n);
  (void)launch_kernel(frontend_kernel, grid, dim3(256), smem, s, kp, B, T, K, prm, fmt, wt_stride);
  return after_launch("frontend_kernel");
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
