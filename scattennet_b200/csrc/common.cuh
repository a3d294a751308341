// Shared host/device helpers for libscatt.so (sm_100a).
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "../../include/scatt.h"

namespace scatt {

// ---------------------------------------------------------------- host-side error plumbing
void set_error(const char* fmt, ...);
extern std::atomic<uint64_t> g_launches;

inline int cuda_fail(cudaError_t e, const char* what) {
  set_error("%s: %s", what, cudaGetErrorString(e));
  return SCATT_ERR_CUDA;
}

#define SCATT_CUDA(expr)                                     \
  do {                                                       \
    cudaError_t e__ = (expr);                                \
    if (e__ != cudaSuccess) return ::scatt::cuda_fail(e__, #expr); \
  } while (0)

#define SCATT_REQUIRE(cond, ...)            \
  do {                                      \
    if (!(cond)) {                          \
      ::scatt::set_error(__VA_ARGS__);      \
      return SCATT_ERR_INVALID;             \
    }                                       \
  } while (0)

// Per-device "done" flag with the interface of std::atomic<bool>: kernel attributes (cudaFuncSetAttribute) belong to
// the device's context, so a process that drives several GPUs must set them once per device, not once per process.
class PerDeviceFlag {
 public:
  bool load() const { return flags_[index()].load(std::memory_order_acquire); }
  void store(bool v) { flags_[index()].store(v, std::memory_order_release); }

 private:
  static int index() {
    int d = 0;
    cudaGetDevice(&d);
    return d & 63;
  }
  std::atomic<bool> flags_[64] = {};
};

// Name (with template arguments, as a demangler prints them) of the kernel this host thread launched last:
// bench.py matches it with the symbols of a profiler trace to attach algorithmic flops / bytes to each kernel.
void set_last_kernel(const char* fmt, ...);

// Call after every kernel launch: counts it and surfaces launch-config errors.
inline int after_launch(const char* name) {
  set_last_kernel("%s", name);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("launch of %s failed: %s", name, cudaGetErrorString(e));
    return SCATT_ERR_CUDA;
  }
  return SCATT_OK;
}

// ---------------------------------------------------------------- programmatic dependent launch
// Every kernel of the library is launched with programmatic stream serialisation: the next kernel's CTAs
// may become resident (barrier init, TMEM allocation, descriptor prefetch, weight staging) while the
// previous kernel drains, and block in pdl_wait() until it has completed and its writes are visible.
extern std::atomic<int> g_pdl;  // 1 = attribute on (default), 0 = plain stream order (SCATT_PDL=0)

#ifdef __CUDACC__
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid, cfg.blockDim = block, cfg.dynamicSmemBytes = smem, cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_pdl.load(std::memory_order_relaxed) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
// Same, as thread-block clusters of `cluster_x` CTAs along grid.x.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                                         unsigned cluster_x, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid, cfg.blockDim = block, cfg.dynamicSmemBytes = smem, cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster_x, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_pdl.load(std::memory_order_relaxed) ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

// ---------------------------------------------------------------- device helpers
#ifdef __CUDACC__

// Blocks until the preceding kernel of the stream has completed (no-op without the launch attribute).
// Warp index for role dispatch, taken through a shuffle so that the compiler treats it - and every branch on it - as
// warp-uniform (CUTLASS's canonical_warp_idx_sync idiom).  With threadIdx.x >> 5 the role bodies count as divergent code:
// every shuffle / vote / elect gets a divergence guard, uniform registers go unused and the attention kernel needs 95
// instead of 76 registers; measured -4.4 % on the B=8 step in an A/B pair (SCATT_UNIFORM_WARP=0 is the old form).
#ifndef SCATT_UNIFORM_WARP
#define SCATT_UNIFORM_WARP 1
#endif
__device__ __forceinline__ int scatt_warp_idx() {
#if SCATT_UNIFORM_WARP
  return __shfl_sync(0xffffffffu, int(threadIdx.x >> 5), 0);
#else
  return int(threadIdx.x >> 5);
#endif
}

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// Lets the next kernel of the stream start its prologue.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

// GELU with erf from Abramowitz-Stegun 7.1.26 (|erf error| <= 1.5e-7, i.e. at the fp32 rounding level of the exact
// form); used in the tensor-core epilogues, where the instruction count per element is the critical path of the
// fused layer tail (profiles/r02_ncu_attn_block_cluster2.txt: a quarter of the stall samples were instruction
// fetches inside the GELU code).  Written as
//   gelu(x) = max(x, 0) - |x| * (p(t) t / 2) * exp(-x^2 / 2),   t = 1 / (1 + 0.3275911 |x| / sqrt 2)
// (1 - erf is what 7.1.26 produces, so neither the "1 +" nor the sign select is needed): 11 fma-pipe instructions
// and two MUFU (rcp.approx.ftz / ex2.approx.ftz - no denormal fix-up code) per element.
__device__ __forceinline__ float gelu_fast(float x) {
  const float u = fabsf(x);
  float t, e;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f * 0.70710678118654752440f, u, 1.0f)));
  float p = fmaf(0.5f * 1.061405429f, t, 0.5f * -1.453152027f);
  p = fmaf(p, t, 0.5f * 1.421413741f);
  p = fmaf(p, t, 0.5f * -0.284496736f);
  p = fmaf(p, t, 0.5f * 0.254829592f);
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(u * u * (-0.5f * 1.4426950408889634f)));
  return fmaf(-(p * t * u), e, fmaxf(x, 0.0f));
}

__device__ __forceinline__ float apply_act(float x, int act) {
  if (act == SCATT_ACT_GELU) return gelu_erf(x);
  if (act == SCATT_ACT_RELU) return fmaxf(x, 0.0f);
  return x;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// 16-bit hi/lo split of an fp32 value. Returned as raw 16-bit patterns.
template <int FMT>
__device__ __forceinline__ void split16(float x, uint16_t& hi, uint16_t& lo) {
  if (FMT == SCATT_PLANE_F16) {
    __half h = __float2half_rn(x);
    __half l = __float2half_rn(x - __half2float(h));
    hi = __half_as_ushort(h);
    lo = __half_as_ushort(l);
  } else {
    __nv_bfloat16 h = __float2bfloat16_rn(x);
    __nv_bfloat16 l = __float2bfloat16_rn(x - __bfloat162float(h));
    hi = __bfloat16_as_ushort(h);
    lo = __bfloat16_as_ushort(l);
  }
}

__device__ __forceinline__ void split16_rt(float x, int fmt, uint16_t& hi, uint16_t& lo) {
  if (fmt == SCATT_PLANE_F16) split16<SCATT_PLANE_F16>(x, hi, lo);
  else split16<SCATT_PLANE_BF16>(x, hi, lo);
}

// hi/lo split of two fp32 values with the packed conversions (same roundings as split16, half the instructions).
__device__ __forceinline__ void split_pair_rt(float a, float b, int fmt, uint32_t& hi, uint32_t& lo) {
  if (fmt == SCATT_PLANE_F16) {
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

// Store 4 consecutive fp32 values as split planes (8-byte stores).
__device__ __forceinline__ void store_planes4(uint16_t* planes, int64_t plane_stride, int64_t off, float4 v, int fmt) {
  uint2 ph, pl;
  split_pair_rt(v.x, v.y, fmt, ph.x, pl.x);
  split_pair_rt(v.z, v.w, fmt, ph.y, pl.y);
  *reinterpret_cast<uint2*>(planes + off) = ph;
  *reinterpret_cast<uint2*>(planes + plane_stride + off) = pl;
}

#endif  // __CUDACC__

// ---------------------------------------------------------------- launchers implemented in the .cu files
int launch_split_planes(const float* x, int64_t rows, int64_t cols, int64_t ldx, float scale, void* planes, int fmt,
                        cudaStream_t s);
int launch_attn_out_q(const scatt_outq_problem* p, int group, int64_t M, int D, int N, float eps, float q_scale, int fmt, int terms,
                      cudaStream_t s);
bool attn_out_q_supported(int64_t M, int D, int N);
int launch_l2_prefetch(const void* const* ptrs, const int64_t* nbytes, int n, cudaStream_t s);
int launch_frontend_tc(const float* kp, int B, int T, int K, const scatt_frontend_stream* streams, int n, int max_pos, int fmt, bool force,
                       cudaStream_t s);
int launch_frontend(const float* kp, int B, int T, int K, int D, const scatt_frontend_stream* streams, int n, int max_pos,
                    int fmt, cudaStream_t s);
int launch_posembed_ln(const float* x, const float* table, const float* g, const float* b, float* out, void* planes,
                       int B, int T, int D, int fmt, cudaStream_t s);
int launch_rowwise(const float* z, int64_t M, int N, int64_t ldz, const float* residual, int64_t ldres, const float* g,
                   const float* b, const scatt_epilogue& ep, float* y, int64_t ldy, void* planes, int fmt,
                   cudaStream_t s);
int launch_rowwise_linear_tail(const scatt_linear_problem* p, int group, int64_t M, int N, int64_t ldres, int64_t ldy,
                               const scatt_epilogue& ep, int fmt, cudaStream_t s);
int launch_pool_pairs(const float* x, int B, int T, int C, float* y, void* planes, int fmt, cudaStream_t s);
int launch_pool_pairs_group(const float* const* xs, float* const* ys, void* const* planes, int group, int B, int T, int C,
                            int fmt, cudaStream_t s);
int launch_linear_simt(const scatt_linear_problem* p, int group, int64_t M, int N, int K, int64_t ldx, int64_t ldres,
                       int64_t ldy, const scatt_epilogue& ep, int fmt, cudaStream_t s);
int launch_linear_tc(const scatt_linear_problem* p, int group, int64_t M, int N, int K, int64_t ldres, int64_t ldy,
                     const scatt_epilogue& ep, int fmt, int terms, void* workspace, size_t workspace_bytes, cudaStream_t s);
int linear_tc_splitk(int64_t M, int N, int K, int group);
size_t linear_tc_workspace_bytes(int64_t M, int N, int K, int group);
int launch_rowwise_splitk(const float* partials, int nsplit, const scatt_linear_problem& p, int64_t M, int N, int64_t ldres,
                          int64_t ldy, const scatt_epilogue& ep, int fmt, cudaStream_t s);
int linear_tc_ln_cluster(int64_t M, int N, int group, int layer_norm);
bool attn_block_supported(int64_t M, int D, int F);
int launch_attn_block(const scatt_block_problem* p, int group, int64_t M, int D, int F, float eps, int fmt, int terms,
                      cudaStream_t s);
int debug_set_trace_block(void* dev_buf);
int debug_set_block_cluster(int cl);
int debug_set_attn_persist(int mode);
int debug_set_trace(void* dev_buf);
int debug_set_trace_attention(void* dev_buf);
int debug_set_trace_fa(void* dev_buf);
int debug_set_trace_ctc(void* dev_buf);
int launch_attention(const scatt_attention_problem* p, int group, int B, int Tq, int Tk, int H, int hd, int64_t ldq,
                     int64_t ldk, int64_t ldv, int kind, int fmt, cudaStream_t s);
bool attention_tc_supported(int Tq, int Tk, int hd, const scatt_attention_problem* p, int group);
int launch_attention_tc(const scatt_attention_problem* p, int group, int B, int Tq, int Tk, int H, int hd, int64_t ldq,
                        int64_t ldk, int64_t ldv, int kind, int fmt, int terms, cudaStream_t s);
bool attention_planes_supported(int Tq, int Tk, int hd);
int launch_attention_planes(const scatt_attention_planes_problem* p, int group, int B, int Tq, int Tk, int H, int hd, int kind,
                            int fmt, int terms, cudaStream_t s);
size_t lstm_workspace_bytes(int64_t B, int H);
int launch_lstm_bidir(const float* gates_x, int64_t ldg, const float* w_hh, float* y, void* y_planes, void* workspace,
                      int64_t B, int T, int H, int fmt, cudaStream_t s);
int launch_log_softmax(const float* logits, int64_t ld, int V, int B, int T, int time_major, float lo, float hi, float* out,
                       cudaStream_t s);
int launch_ctc_beam(const float* logits, int B, int T, int V, const int* lengths, int beam, int* out_ids, int* out_len,
                    float* out_score, cudaStream_t s);
int launch_finite_check(const float* const* tensors, const int64_t* sizes, int count, int* flags, cudaStream_t s);
int launch_peer_allgather(const void* src, int64_t bytes, void* const* peer_bufs, void* const* peer_flags, int world, int rank,
                          void* counter, uint64_t seq, cudaStream_t s);
bool fusion_attention_tc_supported(int T, int D);
int launch_fusion_attention_tc(const void* q_planes, const void* k_planes, const void* v_planes, int B, int T, int D, float* out,
                               void* out_planes, int fmt, int terms, cudaStream_t s);
int launch_fusion_attention(const float* q, const float* k, const float* v, int B, int T, int D, float* out, void* planes,
                            int fmt, cudaStream_t s);

}  // namespace scatt
