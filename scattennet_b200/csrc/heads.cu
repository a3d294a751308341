// Consumers of the per-frame logits (scope rows f-1 and f-3):
//   * log_softmax over the vocabulary + clamp, written batch- or time-major - the
//     CTC front of MSCA_Net.compute_loss (reference model/__init__.py:243-250);
//   * one fused finiteness flag over a set of tensors, replacing the 16 host-synchronising
//     isnan/isinf checks of MSCA_Net.forward (reference model/__init__.py:130-167).
#include "common.cuh"

namespace scatt {

namespace {

constexpr int kLsThreads = 256;

__device__ __forceinline__ float block_reduce(float v, float* red, bool is_max) {
  const int lane = threadIdx.x & 31, warp = scatt_warp_idx();
  v = is_max ? warp_max(v) : warp_sum(v);
  __syncthreads();  // red may still be read from the previous reduction
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = red[0];
#pragma unroll
  for (int w = 1; w < kLsThreads / 32; ++w) r = is_max ? fmaxf(r, red[w]) : r + red[w];
  return r;
}

// One CTA per (b, t) row.  The row (V fp32, 4.5 KB at V = 1120) is read twice from L1/L2.
__global__ void __launch_bounds__(kLsThreads) log_softmax_kernel(const float* __restrict__ x, int64_t ldx, int V, int B, int T,
                                                                int time_major, float lo, float hi,
                                                                float* __restrict__ out) {
  __shared__ float red[kLsThreads / 32];
  pdl_launch_dependents();
  pdl_wait();
  const int64_t row = blockIdx.x;
  const int b = int(row / T), t = int(row % T);
  const float* xr = x + row * ldx;
  float* orow = out + (time_major ? (int64_t(t) * B + b) : row) * V;
  float mx = -INFINITY;
  for (int i = threadIdx.x; i < V; i += kLsThreads) mx = fmaxf(mx, xr[i]);
  mx = block_reduce(mx, red, true);
  float sum = 0.0f;
  for (int i = threadIdx.x; i < V; i += kLsThreads) sum += expf(xr[i] - mx);
  sum = block_reduce(sum, red, false);
  const float lse = mx + logf(sum);
  for (int i = threadIdx.x; i < V; i += kLsThreads) orow[i] = fminf(fmaxf(xr[i] - lse, lo), hi);
}

// V <= 1536 with 16-byte aligned rows: one warp owns a row, which stays in registers as float4 chunks between
// the max, the exp-sum and the store pass (one global read, warp-shuffle reductions, no block barrier).
constexpr int kLsVec = 12;

__global__ void __launch_bounds__(256) log_softmax_warp_kernel(const float* __restrict__ x, int64_t ldx, int V, int B, int T,
                                                               int time_major, float lo, float hi, float* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int64_t M = int64_t(B) * T, warps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  const int nvec = V >> 2;
  for (int64_t row = int64_t(blockIdx.x) * (blockDim.x >> 5) + scatt_warp_idx(); row < M; row += warps) {  // warp-uniform
    const float* xr = x + row * ldx;
    float4 v[kLsVec];
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < kLsVec; ++i)
      if (lane + 32 * i < nvec) {
        v[i] = *reinterpret_cast<const float4*>(xr + 4 * (lane + 32 * i));
        mx = fmaxf(fmaxf(mx, fmaxf(v[i].x, v[i].y)), fmaxf(v[i].z, v[i].w));
      }
    mx = warp_max(mx);
    float sum = 0.0f;
#pragma unroll
    for (int i = 0; i < kLsVec; ++i)
      if (lane + 32 * i < nvec) sum += (expf(v[i].x - mx) + expf(v[i].y - mx)) + (expf(v[i].z - mx) + expf(v[i].w - mx));
    sum = warp_sum(sum);
    const float lse = mx + logf(sum);
    const int b = int(row / T), t = int(row % T);
    float* orow = out + (time_major ? (int64_t(t) * B + b) : row) * V;
#pragma unroll
    for (int i = 0; i < kLsVec; ++i)
      if (lane + 32 * i < nvec) {
        float4 o;
        o.x = fminf(fmaxf(v[i].x - lse, lo), hi), o.y = fminf(fmaxf(v[i].y - lse, lo), hi);
        o.z = fminf(fmaxf(v[i].z - lse, lo), hi), o.w = fminf(fmaxf(v[i].w - lse, lo), hi);
        *reinterpret_cast<float4*>(orow + 4 * (lane + 32 * i)) = o;
      }
  }
}

struct FiniteGroup {
  const float* p[SCATT_MAX_FINITE];
  int64_t n[SCATT_MAX_FINITE];
};

// flags |= 1 << i  when tensor i holds a NaN or an infinity (all-ones exponent)
__global__ void __launch_bounds__(256) finite_check_kernel(FiniteGroup g, int* __restrict__ flags) {
  pdl_launch_dependents();
  pdl_wait();
  const float* p = g.p[blockIdx.y];
  const int64_t n = g.n[blockIdx.y];
  bool bad = false;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x)
    bad |= (__float_as_uint(p[i]) & 0x7f800000u) == 0x7f800000u;
  if (__syncthreads_or(bad) && threadIdx.x == 0) atomicOr(flags, 1 << blockIdx.y);
}

}  // namespace

int launch_log_softmax(const float* logits, int64_t ld, int V, int B, int T, int time_major, float lo, float hi, float* out,
                       cudaStream_t s) {
  SCATT_REQUIRE(logits && out && V >= 1 && ld >= V, "log_softmax: bad argument");
  SCATT_REQUIRE(lo <= hi, "log_softmax: clamp range [%g, %g] is empty", double(lo), double(hi));
  if (int64_t(B) * T == 0) return SCATT_OK;
  const bool vec = V % 4 == 0 && V <= 4 * 32 * kLsVec && ld % 4 == 0 && (reinterpret_cast<uintptr_t>(logits) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(out) & 15) == 0;
  if (vec) {
    int64_t blocks = (int64_t(B) * T + 7) / 8;
    if (blocks > 148 * 8) blocks = 148 * 8;
    (void)launch_kernel(log_softmax_warp_kernel, dim3(unsigned(blocks)), dim3(256), 0, s, logits, ld, V, B, T, time_major, lo, hi, out);
    return after_launch("log_softmax_warp_kernel");
  }
  (void)launch_kernel(log_softmax_kernel, dim3(unsigned(int64_t(B) * T)), dim3(kLsThreads), 0, s, logits, ld, V, B, T,
                      time_major, lo, hi, out);
  return after_launch("log_softmax_kernel");
}

int launch_finite_check(const float* const* tensors, const int64_t* sizes, int count, int* flags, cudaStream_t s) {
  SCATT_REQUIRE(tensors && sizes && flags && count >= 1 && count <= SCATT_MAX_FINITE, "finite_check: 1..%d tensors",
                SCATT_MAX_FINITE);
  FiniteGroup g{};
  int64_t longest = 0;
  for (int i = 0; i < count; ++i) {
    SCATT_REQUIRE(tensors[i] || sizes[i] == 0, "finite_check: tensor %d is null", i);
    g.p[i] = tensors[i], g.n[i] = sizes[i];
    if (sizes[i] > longest) longest = sizes[i];
  }
  SCATT_CUDA(cudaMemsetAsync(flags, 0, sizeof(int), s));
  if (longest == 0) return SCATT_OK;
  int64_t blocks = (longest + 256 * 8 - 1) / (256 * 8);
  if (blocks > 148 * 4) blocks = 148 * 4;
  (void)launch_kernel(finite_check_kernel, dim3(unsigned(blocks), count), dim3(256), 0, s, g, flags);
  return after_launch("finite_check_kernel");
}

}  // namespace scatt
