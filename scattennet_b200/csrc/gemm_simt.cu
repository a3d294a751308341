// SIMT engine of scatt_linear: y = epilogue(x W^T + bias) with fp32 FMA on the
// CUDA cores.  This is the exact-arithmetic engine (fp32 tier, and the on-GPU
// cross-check of the tcgen05 engine); it is not the throughput path.
//
// 64x64 output tile per CTA, 256 threads, 4x4 outputs per thread, K step 16,
// operands staged transposed in shared memory so the inner product reads are
// conflict-free float4s.
#include "common.cuh"

namespace scatt {

namespace {

struct LinearGroup {
  scatt_linear_problem p[SCATT_MAX_GROUP];
};

constexpr int BM = 64, BN = 64, BK = 16;

__global__ void __launch_bounds__(256) linear_simt_kernel(LinearGroup grp, int64_t M, int N, int K, int64_t ldx,
                                                          int64_t ldres, int64_t ldy, scatt_epilogue ep, int fmt,
                                                          int fuse_tail) {
  const scatt_linear_problem& P = grp.p[blockIdx.z];
  pdl_launch_dependents();
  pdl_wait();
  __shared__ __align__(16) float xs[BK][BM + 4];
  __shared__ __align__(16) float ws[BK][BN + 4];
  const int tid = threadIdx.x;
  const int64_t m0 = int64_t(blockIdx.y) * BM;
  const int n0 = blockIdx.x * BN;
  const int tx = tid & 15, ty = tid >> 4;  // thread computes rows ty*4.., cols tx*4..
  float acc[4][4] = {};

  // loader mapping: 256 threads x one float4 along K = 64 rows x 16 k
  const int lr = tid >> 2, lk = (tid & 3) * 4;
  for (int k0 = 0; k0 < K; k0 += BK) {
    float4 xv = make_float4(0.f, 0.f, 0.f, 0.f), wv = xv;
    if (m0 + lr < M) xv = *reinterpret_cast<const float4*>(P.x + (m0 + lr) * ldx + k0 + lk);
    if (n0 + lr < N) wv = *reinterpret_cast<const float4*>(P.w + int64_t(n0 + lr) * K + k0 + lk);
    xs[lk + 0][lr] = xv.x, xs[lk + 1][lr] = xv.y, xs[lk + 2][lr] = xv.z, xs[lk + 3][lr] = xv.w;
    ws[lk + 0][lr] = wv.x, ws[lk + 1][lr] = wv.y, ws[lk + 2][lr] = wv.z, ws[lk + 3][lr] = wv.w;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&xs[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&ws[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

  const int n = n0 + tx * 4;
  if (n >= N) return;
  float bias[4] = {0.f, 0.f, 0.f, 0.f};
  if (P.bias) {
    const float4 b4 = *reinterpret_cast<const float4*>(P.bias + n);
    bias[0] = b4.x, bias[1] = b4.y, bias[2] = b4.z, bias[3] = b4.w;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + ty * 4 + i;
    if (m >= M) continue;
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float v = acc[i][j] + bias[j];
      if (n + j < ep.scale_cols) v *= ep.scale;
      o[j] = apply_act(v, ep.act_pre);
    }
    if (ep.residual_mode == SCATT_RES_BEFORE_LN || (fuse_tail && ep.residual_mode == SCATT_RES_AFTER_LN)) {
      const float4 r = *reinterpret_cast<const float4*>(P.residual + m * ldres + n);
      o[0] += r.x, o[1] += r.y, o[2] += r.z, o[3] += r.w;
    }
    if (fuse_tail) {  // no LayerNorm requested: finish the chain here
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        o[j] = apply_act(o[j], ep.act_post);
        if (ep.clamp > 0.f) o[j] = fminf(fmaxf(o[j], -ep.clamp), ep.clamp);
      }
    }
    const float4 ov = make_float4(o[0], o[1], o[2], o[3]);
    if (P.y) *reinterpret_cast<float4*>(P.y + m * ldy + n) = ov;
    if (fuse_tail && P.y_planes)
      store_planes4(reinterpret_cast<uint16_t*>(P.y_planes), M * int64_t(N), m * N + n, ov, fmt);
  }
}

}  // namespace

int launch_linear_simt(const scatt_linear_problem* p, int group, int64_t M, int N, int K, int64_t ldx, int64_t ldres,
                       int64_t ldy, const scatt_epilogue& ep, int fmt, cudaStream_t s) {
  SCATT_REQUIRE(K % BK == 0, "linear(simt): K=%d must be a multiple of %d", K, BK);
  SCATT_REQUIRE(N % 4 == 0, "linear(simt): N=%d must be a multiple of 4", N);
  SCATT_REQUIRE(ldx % 4 == 0 && ldy % 4 == 0 && ldres % 4 == 0, "linear(simt): row strides must be multiples of 4");
  LinearGroup grp{};
  for (int i = 0; i < group; ++i) {
    grp.p[i] = p[i];
    SCATT_REQUIRE(p[i].x && p[i].w, "linear(simt): problem %d lacks fp32 x / w", i);
    SCATT_REQUIRE(ep.residual_mode == SCATT_RES_NONE || p[i].residual, "linear(simt): residual missing");
    SCATT_REQUIRE(!ep.layer_norm || (p[i].y && p[i].ln_g && p[i].ln_b),
                  "linear(simt): LayerNorm needs y (scratch), gamma and beta");
  }
  if (M == 0) return SCATT_OK;
  const int fuse_tail = ep.layer_norm ? 0 : 1;
  dim3 grid((N + BN - 1) / BN, unsigned((M + BM - 1) / BM), group);
  (void)launch_kernel(linear_simt_kernel, grid, dim3(256), 0, s, grp, M, N, K, ldx, ldres, ldy, ep, fmt, fuse_tail);
  int rc = after_launch("linear_simt_kernel");
  if (rc != SCATT_OK || fuse_tail) return rc;
  return launch_rowwise_linear_tail(p, group, M, N, ldres, ldy, ep, fmt, s);
}

}  // namespace scatt
