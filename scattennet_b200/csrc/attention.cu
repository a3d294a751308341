// K3 stream attention (self / causal / cross, head_dim 16) and K5 fusion
// attention (single head, width D) - fp32 flash-style kernels on the CUDA
// cores.  Scores and masks are never materialised: the key-padding /
// additive / causal semantics of the reference (model/utils.py:3-28,
// model/attention.py:63-72,165-171) are applied to each logit in registers.
#include <cfloat>

#include "common.cuh"

namespace scatt {

namespace {

struct AttnGroup {
  scatt_attention_problem p[SCATT_MAX_GROUP];
};

constexpr int HD = 16;

// One thread = one query row of one head; K and V of the (batch, head) pair
// live in shared memory and are read as warp-wide broadcasts.
// grid = (ceil(Tq / 128), H, B * group), block = 128.
__global__ void __launch_bounds__(128) stream_attention_kernel(AttnGroup grp, int B, int Tq, int Tk, int H,
                                                               int64_t ldq, int64_t ldk, int64_t ldv, int kind,
                                                               int fmt) {
  extern __shared__ __align__(16) float smem[];
  float* ks = smem;                          // [Tk][16]
  float* vs = smem + size_t(Tk) * HD;        // [Tk][16]
  float* pad = vs + size_t(Tk) * HD;         // [Tk] 0 / -FLT_MAX
  const int g = blockIdx.z / B, b = blockIdx.z % B, h = blockIdx.y;
  const scatt_attention_problem& P = grp.p[g];
  const int D = H * HD;
  pdl_launch_dependents();
  pdl_wait();

  // causal CTAs only need keys up to their last query
  const int q_hi = min(Tq, int(blockIdx.x + 1) * 128);
  const int nk = (kind == SCATT_ATTN_CAUSAL) ? min(Tk, q_hi) : Tk;
  for (int i = threadIdx.x; i < nk * 4; i += blockDim.x) {
    const int j = i >> 2, c = (i & 3) * 4;
    *reinterpret_cast<float4*>(ks + j * HD + c) =
        *reinterpret_cast<const float4*>(P.k + (int64_t(b) * Tk + j) * ldk + h * HD + c);
    *reinterpret_cast<float4*>(vs + j * HD + c) =
        *reinterpret_cast<const float4*>(P.v + (int64_t(b) * Tk + j) * ldv + h * HD + c);
  }
  for (int j = threadIdx.x; j < nk; j += blockDim.x)
    pad[j] = (P.key_mask && P.key_mask[int64_t(b) * Tk + j] == 0) ? -FLT_MAX : 0.f;
  __syncthreads();

  const int i = blockIdx.x * 128 + threadIdx.x;
  if (i >= Tq) return;
  const int64_t row = int64_t(b) * Tq + i;
  float q[HD];
#pragma unroll
  for (int c = 0; c < HD; c += 4) {
    const float4 t = *reinterpret_cast<const float4*>(P.q + row * ldq + h * HD + c);
    q[c] = t.x, q[c + 1] = t.y, q[c + 2] = t.z, q[c + 3] = t.w;
  }
  const float* add = P.additive ? P.additive + (int64_t(b) * Tq + i) * Tk : nullptr;
  const int jend = (kind == SCATT_ATTN_CAUSAL) ? min(i + 1, Tk) : Tk;

  float m = -INFINITY, l = 0.f, acc[HD] = {};
  for (int j = 0; j < jend; ++j) {
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < HD; c += 4) {
      const float4 kk = *reinterpret_cast<const float4*>(ks + j * HD + c);
      s = fmaf(q[c], kk.x, s), s = fmaf(q[c + 1], kk.y, s), s = fmaf(q[c + 2], kk.z, s), s = fmaf(q[c + 3], kk.w, s);
    }
    // additive masks exactly as the reference adds them (fp32 '+'): a padded
    // key collapses to finfo.min, so an all-padded row ends up uniform.
    if (add) s += add[j];
    s += pad[j];
    const float mn = fmaxf(m, s);
    const float corr = __expf(m - mn);  // exp(-inf) = 0 on the first key
    const float p = __expf(s - mn);
    l = l * corr + p;
#pragma unroll
    for (int c = 0; c < HD; c += 4) {
      const float4 vv = *reinterpret_cast<const float4*>(vs + j * HD + c);
      acc[c] = fmaf(p, vv.x, acc[c] * corr), acc[c + 1] = fmaf(p, vv.y, acc[c + 1] * corr);
      acc[c + 2] = fmaf(p, vv.z, acc[c + 2] * corr), acc[c + 3] = fmaf(p, vv.w, acc[c + 3] * corr);
    }
    m = mn;
  }
  const float inv = 1.0f / l;
#pragma unroll
  for (int c = 0; c < HD; c += 4) {
    const float4 o = make_float4(acc[c] * inv, acc[c + 1] * inv, acc[c + 2] * inv, acc[c + 3] * inv);
    if (P.out) *reinterpret_cast<float4*>(P.out + row * D + h * HD + c) = o;
    if (P.out_planes)
      store_planes4(reinterpret_cast<uint16_t*>(P.out_planes), int64_t(B) * Tq * D, row * D + h * HD + c, o, fmt);
  }
}

// Fusion attention (single head of width D, no mask, no scaling): one CTA = FQ query rows of one batch
// element, 256 threads.
//   phase 1  warp w takes keys j = w, w+8, ...: it reads the key row once (coalesced float4s straight from
//            L2) and dots it with all FQ query rows held in shared memory -> logits[FQ][T]
//   phase 2  one warp per query row: max, exp, sum over the T logits
//   phase 3  thread t owns output columns 4t..4t+3 of all FQ rows and walks the value rows (coalesced)
// Requires D % 128 == 0 and D <= 1024.
constexpr int FQ = 4, FMAXV = 8;

__global__ void __launch_bounds__(256) fusion_attention_kernel(const float* __restrict__ q, const float* __restrict__ k,
                                                               const float* __restrict__ v, int B, int T, int D,
                                                               float* __restrict__ out, uint16_t* __restrict__ planes,
                                                               int fmt) {
  extern __shared__ __align__(16) float smem[];
  float* qs = smem;                    // [FQ][D]
  float* sc = smem + size_t(FQ) * D;   // [FQ][T] logits, then probabilities
  float* inv = sc + size_t(FQ) * T;    // [FQ] 1 / row sum
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31, warp = scatt_warp_idx();
  const int b = blockIdx.y, i0 = blockIdx.x * FQ;
  const int nq = min(FQ, T - i0);
  const int nv = D >> 7;  // float4 chunks per lane
  const float* qb = q + (int64_t(b) * T + i0) * D;
  const float* kb = k + int64_t(b) * T * D;
  const float* vb = v + int64_t(b) * T * D;
  for (int e = threadIdx.x; e < FQ * (D >> 2); e += blockDim.x) {
    const int r = e / (D >> 2), c = (e % (D >> 2)) * 4;
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < nq) t = *reinterpret_cast<const float4*>(qb + int64_t(r) * D + c);
    *reinterpret_cast<float4*>(qs + r * D + c) = t;
  }
  __syncthreads();

  // phase 1: logits
  for (int j = warp; j < T; j += 8) {
    float4 kr[FMAXV];
#pragma unroll
    for (int c = 0; c < FMAXV; ++c)
      if (c < nv) kr[c] = *reinterpret_cast<const float4*>(kb + int64_t(j) * D + 4 * (lane + 32 * c));
    float s[FQ];
#pragma unroll
    for (int r = 0; r < FQ; ++r) {
      float a0 = 0.f, a1 = 0.f;  // two chains per row
#pragma unroll
      for (int c = 0; c < FMAXV; ++c)
        if (c < nv) {
          const float4 qq = *reinterpret_cast<const float4*>(qs + r * D + 4 * (lane + 32 * c));
          a0 = fmaf(qq.x, kr[c].x, a0), a1 = fmaf(qq.y, kr[c].y, a1);
          a0 = fmaf(qq.z, kr[c].z, a0), a1 = fmaf(qq.w, kr[c].w, a1);
        }
      s[r] = a0 + a1;
    }
#pragma unroll
    for (int r = 0; r < FQ; ++r) s[r] = warp_sum(s[r]);
    if (lane == 0) {
#pragma unroll
      for (int r = 0; r < FQ; ++r) sc[r * T + j] = s[r];
    }
  }
  __syncthreads();
  // phase 2: softmax of each row (warps 0 .. FQ-1)
  if (warp < FQ) {
    float mx = -INFINITY;
    for (int j = lane; j < T; j += 32) mx = fmaxf(mx, sc[warp * T + j]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < T; j += 32) {
      const float p = expf(sc[warp * T + j] - mx);
      sc[warp * T + j] = p;
      sum += p;
    }
    sum = warp_sum(sum);
    if (lane == 0) inv[warp] = 1.0f / sum;
  }
  __syncthreads();
  // phase 3: out = P V
  const int col = 4 * threadIdx.x;
  if (col < D) {
    float4 acc[FQ];
#pragma unroll
    for (int r = 0; r < FQ; ++r) acc[r] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 5
    for (int j = 0; j < T; ++j) {
      const float4 vv = *reinterpret_cast<const float4*>(vb + int64_t(j) * D + col);
#pragma unroll
      for (int r = 0; r < FQ; ++r) {
        const float p = sc[r * T + j];
        acc[r].x = fmaf(p, vv.x, acc[r].x), acc[r].y = fmaf(p, vv.y, acc[r].y);
        acc[r].z = fmaf(p, vv.z, acc[r].z), acc[r].w = fmaf(p, vv.w, acc[r].w);
      }
    }
#pragma unroll
    for (int r = 0; r < FQ; ++r)
      if (r < nq) {
        const float f = inv[r];
        const float4 o = make_float4(acc[r].x * f, acc[r].y * f, acc[r].z * f, acc[r].w * f);
        const int64_t row = int64_t(b) * T + i0 + r;
        if (out) *reinterpret_cast<float4*>(out + row * D + col) = o;
        if (planes) store_planes4(planes, int64_t(B) * T * D, row * D + col, o, fmt);
      }
  }
}

}  // namespace

int launch_attention(const scatt_attention_problem* p, int group, int B, int Tq, int Tk, int H, int hd, int64_t ldq,
                     int64_t ldk, int64_t ldv, int kind, int fmt, cudaStream_t s) {
  SCATT_REQUIRE(hd == HD, "attention: head_dim must be 16 (got %d)", hd);
  SCATT_REQUIRE(group >= 1 && group <= SCATT_MAX_GROUP, "attention: group 1..%d", SCATT_MAX_GROUP);
  SCATT_REQUIRE(ldq % 4 == 0 && ldk % 4 == 0 && ldv % 4 == 0, "attention: row strides must be multiples of 4");
  SCATT_REQUIRE(kind != SCATT_ATTN_CAUSAL || Tq == Tk, "attention: causal needs Tq == Tk");
  SCATT_REQUIRE(int64_t(B) * group <= 65535 && H <= 65535, "attention: grid too large");
  if (B == 0 || Tq == 0) return SCATT_OK;
  SCATT_REQUIRE(Tk >= 1, "attention: no keys");
  AttnGroup grp{};
  for (int i = 0; i < group; ++i) {
    grp.p[i] = p[i];
    SCATT_REQUIRE(p[i].q && p[i].k && p[i].v && (p[i].out || p[i].out_planes), "attention: null operand");
  }
  const size_t smem = (size_t(Tk) * HD * 2 + Tk) * sizeof(float);
  SCATT_REQUIRE(smem <= 200 * 1024, "attention: Tk=%d too long for the shared-memory K/V stage", Tk);
  static PerDeviceFlag attr_done;
  if (!attr_done.load()) {
    SCATT_CUDA(cudaFuncSetAttribute(stream_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_done.store(true);
  }
  dim3 grid((Tq + 127) / 128, H, B * group);
  (void)launch_kernel(stream_attention_kernel, grid, dim3(128), smem, s, grp, B, Tq, Tk, H, ldq, ldk, ldv, kind, fmt);
  return after_launch("stream_attention_kernel");
}

int launch_fusion_attention(const float* q, const float* k, const float* v, int B, int T, int D, float* out, void* planes,
                            int fmt, cudaStream_t s) {
  SCATT_REQUIRE(D % 128 == 0 && D <= 1024, "fusion_attention: D=%d must be a multiple of 128, <= 1024", D);
  SCATT_REQUIRE(B <= 65535, "fusion_attention: batch too large for one launch");
  if (B == 0 || T == 0) return SCATT_OK;
  const size_t smem = (size_t(FQ) * D + size_t(FQ) * T + FQ) * sizeof(float);
  SCATT_REQUIRE(smem <= 200 * 1024, "fusion_attention: T=%d too long", T);
  static PerDeviceFlag attr_done;
  if (!attr_done.load()) {
    SCATT_CUDA(cudaFuncSetAttribute(fusion_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_done.store(true);
  }
  dim3 grid((T + FQ - 1) / FQ, B);
  (void)launch_kernel(fusion_attention_kernel, grid, dim3(256), smem, s, q, k, v, B, T, D, out, reinterpret_cast<uint16_t*>(planes), fmt);
  return after_launch("fusion_attention_kernel");
}

}  // namespace scatt
