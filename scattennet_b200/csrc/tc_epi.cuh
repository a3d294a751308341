// Epilogue helpers shared by the tcgen05 kernels (gemm_tc.cu, block_tc.cu): split-plane packing,
// TMA bulk tensor stores, per-column parameter adds, thread-block-cluster primitives.
#pragma once

#include "common.cuh"
#include "tc_ptx.cuh"

namespace scatt {
namespace tc {

__device__ __forceinline__ float2 unpack_pair(float hi_bits, float lo_bits, int fmt) {
  const uint32_t h = __float_as_uint(hi_bits), l = __float_as_uint(lo_bits);
  float2 a, b;
  if (fmt == SCATT_PLANE_F16) {
    a = __half22float2(*reinterpret_cast<const __half2*>(&h));
    b = __half22float2(*reinterpret_cast<const __half2*>(&l));
  } else {
    a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&h));
    b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&l));
  }
  return make_float2(a.x + b.x, a.y + b.y);
}
__device__ __forceinline__ void add_cols(float* v, const float* __restrict__ p) {  // p: shared memory, warp-uniform
#pragma unroll
  for (int j = 0; j < 32; j += 4) {
    const float4 t = *reinterpret_cast<const float4*>(p + j);
    const float2 a = __fadd2_rn(make_float2(v[j], v[j + 1]), make_float2(t.x, t.y));
    const float2 b = __fadd2_rn(make_float2(v[j + 2], v[j + 3]), make_float2(t.z, t.w));
    v[j] = a.x, v[j + 1] = a.y, v[j + 2] = b.x, v[j + 3] = b.y;
  }
}

template <int FMT>
__device__ __forceinline__ void split8(const float4& a, const float4& b, uint4& hi, uint4& lo) {
  const float x[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  uint32_t h[4], l[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    if (FMT == SCATT_PLANE_F16) {
      const __half2 hh = __floats2half2_rn(x[2 * e], x[2 * e + 1]);
      const float2 back = __half22float2(hh);
      const float2 d = __fadd2_rn(make_float2(x[2 * e], x[2 * e + 1]), make_float2(-back.x, -back.y));  // packed: one FADD2
      const __half2 ll = __floats2half2_rn(d.x, d.y);
      h[e] = *reinterpret_cast<const uint32_t*>(&hh);
      l[e] = *reinterpret_cast<const uint32_t*>(&ll);
    } else {
      const __nv_bfloat162 hh = __floats2bfloat162_rn(x[2 * e], x[2 * e + 1]);
      const float2 back = __bfloat1622float2(hh);
      const float2 d = __fadd2_rn(make_float2(x[2 * e], x[2 * e + 1]), make_float2(-back.x, -back.y));
      const __nv_bfloat162 ll = __floats2bfloat162_rn(d.x, d.y);
      h[e] = *reinterpret_cast<const uint32_t*>(&hh);
      l[e] = *reinterpret_cast<const uint32_t*>(&ll);
    }
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map), "r"(src), "r"(c0),
               "r"(c1), "r"(c2)
               : "memory");
}

// cluster helpers (LN >= 2: the row's columns live in the LN CTAs of a cluster)
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_sync_all() {
  cluster_arrive();
  cluster_wait();
}
__device__ __forceinline__ void st_peer_f32x2(uint32_t local_addr, uint32_t peer_rank, float a, float b) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_addr), "r"(peer_rank));
  asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(remote), "f"(a), "f"(b) : "memory");
}


}  // namespace tc
}  // namespace scatt
