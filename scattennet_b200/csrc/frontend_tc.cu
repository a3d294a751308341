// scatt_frontend on the tensor cores (sm_100a): K1 for large batches, planes-only output
//
//   out[br] = LayerNorm_br(W_br c + b_br + pos_br[t + 2])     c = the stream's joints' x or y coordinates of a frame
//   (region gather model/__init__.py:133-142, x / y split model/keypoint_module.py:23-24, CoordinateMapping
//    model/layers.py:118-123, LearningPositionEmbedding model/layers.py:20-30, first_*_norm keypoint_module.py:155-162)
//
// The CUDA-core kernel (rowwise.cu: frontend_kernel) spends ~660 warp instructions per frame and stream on the
// K_s -> 256 mapping and is bound by instruction issue at 23-26 % of the HBM copy rate (profiles/r02_sweep_membound.md).
// Here the mapping is a [128 frames x 32] x [32 x 256] product on tcgen05 (coordinates and weights as fp16 hi / lo
// split planes, three product terms, fp32 accumulation: ~2^-22 relative, the same class as every other GEMM of the
// path), so the SM only gathers, normalises, splits and stores.
//
// One CTA works on ONE (stream, branch) pair - its mapping weight (B operand, 32 KB) and column parameters are staged
// once - and walks 128-frame tiles of it.  352 threads:
//   warps 0-7  epilogue, two per TMEM lane quadrant (128 columns each): pass 1 adds bias + position row and
//              accumulates the row statistics (values written back to TMEM), pass 2 normalises, splits into hi / lo
//              and hands 64-column boxes (hi tile | lo tile, 128-byte swizzle) to the TMA engine - one store per 8 KB.
//   warp 8     TMEM allocation + tcgen05.mma issue: <= 6 MMAs (N = 256, K = 16) per tile into one of TWO accumulators,
//              so the MMAs of tile i + 1 run under the epilogue of tile i.
//   warps 9-10 loaders: the region gather (thread = frame, 4-byte loads of the used joints' coordinate straight from
//              keypoints[B,T,K,2]), split into fp16 hi / lo and written as K-major, 32-byte-swizzled A tiles
//              (double-buffered) - a tile ahead of the MMAs.
// Algorithmic traffic per frame and stream: 8 n_joints bytes read (sector granularity makes that up to 32 n_joints),
// 2 x 256 x 4 bytes written.
#include <cuda.h>

#include <cstdlib>

#include "common.cuh"
#include "tc_epi.cuh"
#include "tc_host.cuh"
#include "tc_ptx.cuh"

namespace scatt {

namespace {

using namespace tc;

constexpr int kFtEpiWarps = 8;
constexpr int kFtMmaWarp = kFtEpiWarps;
constexpr int kFtLoadWarps = 2;
constexpr int kFtThreads = 32 * (kFtEpiWarps + 1 + kFtLoadWarps);
constexpr int kFtD = 256;
constexpr uint32_t kFtATile = 128 * 32;  // [128 frames x 16 joints] fp16, 32-byte rows
constexpr uint32_t kFtWTile = 256 * 32;  // [256 channels x 16 joints]
// shared memory map (relative to the 1024-aligned base)
constexpr uint32_t kFtWOff = 0;                             // [ks 2][plane 2] W tiles
constexpr uint32_t kFtAOff = kFtWOff + 4 * kFtWTile;        // [stage 2][ks 2][plane 2] A tiles
constexpr uint32_t kFtOutOff = kFtAOff + 2 * 4 * kFtATile;  // [warp 8][buffer 2][hi 4 KB | lo 4 KB] output boxes
constexpr uint32_t kFtColOff = kFtOutOff + kFtEpiWarps * 2 * 8192;  // float[3][256]: bias, gamma, beta
constexpr uint32_t kFtIdxOff = kFtColOff + 3 * kFtD * 4;    // int[32] joint indices
constexpr uint32_t kFtStatOff = kFtIdxOff + 128;            // float2[parity 2][half 2][128] row statistics of the two column halves
constexpr uint32_t kFtBarOff = kFtStatOff + 2 * 2 * 128 * 8;    // a_full[2], a_empty[2], acc_full[2], acc_empty[2], tmem ptr
constexpr uint32_t kFtSmemBytes = kFtBarOff + 128 + 1024;   // + alignment slack
static_assert(kFtOutOff % 1024 == 0, "128-byte-swizzled boxes need 1024-byte alignment");
static_assert(kFtSmemBytes <= 227 * 1024, "frontend_tc: shared memory map exceeds 227 KB");

struct FtPair {  // one (stream, branch)
  const int32_t* joint_idx;
  const float* wt;   // [n_joints][256] transposed mapping weight
  const float* bias; // [256]
  const float* pos;  // [max_pos + 2][256]
  const float* ln_g;
  const float* ln_b;
  int32_t n_joints, coord;
};

struct alignas(64) FtParams {
  CUtensorMap map_out[2 * SCATT_MAX_GROUP];  // planes [2][M][256]: box 64 x 32 x 2
  FtPair pair[2 * SCATT_MAX_GROUP];
  const float* kp;
  int64_t M;
  int32_t T, K, npairs, ctas_per_pair, tiles_m;
  int32_t ablate;  // dev builds (-DSCATT_FT_ABLATE=1, env SCATT_FT_DBG): 1 = no position loads, 2 = no gather loads, 4 = no stores
};
#ifndef SCATT_FT_ABLATE
#define SCATT_FT_ABLATE 0
#endif

// K-major tile with 32-byte rows (16 K-elements), 32-byte swizzle, 8-row groups 256 B apart
__device__ __forceinline__ uint64_t ft_desc_sw32(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= uint64_t((smem_addr & 0x3FFFFu) >> 4);
  d |= uint64_t(1) << 16;
  d |= uint64_t(256 >> 4) << 32;
  d |= uint64_t(1) << 46;
  d |= uint64_t(6) << 61;
  return d;
}
// byte offset of the 16-byte chunk q (8 K-elements) of row r inside such a tile
__device__ __forceinline__ uint32_t ft_chunk_off(int r, int q) {
  return uint32_t((r >> 3) * 256 + (r & 7) * 32 + ((q ^ ((r >> 2) & 1)) << 4));
}

__device__ __forceinline__ void ft_epi_bar() { asm volatile("bar.sync 1, %0;" ::"n"(32 * kFtEpiWarps) : "memory"); }

__global__ void __launch_bounds__(kFtThreads, 1) frontend_tc_kernel(const __grid_constant__ FtParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  const uint32_t bar0 = base + kFtBarOff;
  auto a_full = [&](uint32_t s) { return bar0 + 8u * s; };
  auto a_empty = [&](uint32_t s) { return bar0 + 16u + 8u * s; };
  auto acc_full = [&](uint32_t s) { return bar0 + 32u + 8u * s; };
  auto acc_empty = [&](uint32_t s) { return bar0 + 48u + 8u * s; };
  const uint32_t tmem_ptr_addr = bar0 + 64u;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pi = int(blockIdx.x) / P.ctas_per_pair, slot = int(blockIdx.x) % P.ctas_per_pair;
  const FtPair& Q = P.pair[pi];
  const int nj = Q.n_joints, nks = nj > 16 ? 2 : 1;
  const int ntiles = slot < P.tiles_m ? (P.tiles_m - slot + P.ctas_per_pair - 1) / P.ctas_per_pair : 0;

  if (threadIdx.x == 0) {
    for (uint32_t s = 0; s < 2; ++s) {
      mbar_init(a_full(s), 32 * kFtLoadWarps);
      mbar_init(a_empty(s), 1);
      mbar_init(acc_full(s), 1);
      mbar_init(acc_empty(s), 32 * kFtEpiWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&P.map_out[pi]) : "memory");
  }
  if (warp == kFtMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_addr), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  } else {
    // static data of the pair: W^T [nj][256] fp32 -> [ks][plane] K-major tiles (zero beyond nj), column parameters, indices
    const int tid = warp < kFtMmaWarp ? int(threadIdx.x) : int(threadIdx.x) - 32;
    constexpr int kStagers = kFtThreads - 32;
    for (int item = tid; item < kFtD * 2 * nks; item += kStagers) {
      const int n = item % kFtD, c = item / kFtD, ks = c >> 1, q = c & 1;
      float w[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int k = ks * 16 + q * 8 + e;
        w[e] = k < nj ? Q.wt[k * kFtD + n] : 0.f;
      }
      uint4 hi, lo;
      split8<SCATT_PLANE_F16>(make_float4(w[0], w[1], w[2], w[3]), make_float4(w[4], w[5], w[6], w[7]), hi, lo);
      const uint32_t off = kFtWOff + uint32_t(ks) * 2u * kFtWTile + ft_chunk_off(n, q);
      *reinterpret_cast<uint4*>(sm + off) = hi;
      *reinterpret_cast<uint4*>(sm + off + kFtWTile) = lo;
    }
    float* col = reinterpret_cast<float*>(sm + kFtColOff);
    for (int i = tid; i < kFtD; i += kStagers) {
      col[i] = Q.bias[i];
      col[kFtD + i] = Q.ln_g[i];
      col[2 * kFtD + i] = Q.ln_b[i];
    }
    if (tid < 32) reinterpret_cast<int32_t*>(sm + kFtIdxOff)[tid] = tid < nj ? Q.joint_idx[tid] : 0;
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_launch_dependents();
  pdl_wait();  // everything above is static; keypoints / outputs follow stream order
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(sm + kFtBarOff + 64);

  if (warp > kFtMmaWarp) {  // ================================================= loaders: gather + split -> A tiles
    const int tid = int(threadIdx.x) - 32 * (kFtMmaWarp + 1);  // 0..63: frames tid and tid + 64 of the tile
    const int32_t* idx = reinterpret_cast<const int32_t*>(sm + kFtIdxOff);
    const float* kpc = P.kp + Q.coord;
    for (int it = 0; it < ntiles; ++it) {
      const uint32_t s = uint32_t(it) & 1u;
      const int64_t m0 = int64_t(slot + it * P.ctas_per_pair) * 128;
      if (it >= 2) mbar_wait(a_empty(s), ((uint32_t(it) >> 1) & 1u) ^ 1u);
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        const int r = tid + 64 * h;
        const int64_t row = m0 + r;
        const float* src = kpc + row * int64_t(P.K) * 2;
        const bool valid = row < P.M;
#pragma unroll 1
        for (int c = 0; c < 2 * nks; ++c) {
          float v[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int k = c * 8 + e;
            v[e] = (valid && k < nj && !(SCATT_FT_ABLATE && (P.ablate & 2))) ? __ldg(src + idx[k] * 2) : 0.f;
          }
          uint4 hi, lo;
          split8<SCATT_PLANE_F16>(make_float4(v[0], v[1], v[2], v[3]), make_float4(v[4], v[5], v[6], v[7]), hi, lo);
          const uint32_t off = kFtAOff + s * 4u * kFtATile + uint32_t(c >> 1) * 2u * kFtATile + ft_chunk_off(r, c & 1);
          *reinterpret_cast<uint4*>(sm + off) = hi;
          *reinterpret_cast<uint4*>(sm + off + kFtATile) = lo;
        }
      }
      fence_proxy_async();  // generic-proxy writes -> visible to the tensor core's operand reads
      mbar_arrive(a_full(s));
    }
  } else if (warp == kFtMmaWarp) {  // ========================================== MMA issuer
    const uint32_t idesc = (1u << 4) | (uint32_t(kFtD >> 3) << 17) | (uint32_t(128 >> 4) << 24);  // f16 x f16 -> f32, M 128, N 256
    for (int it = 0; it < ntiles; ++it) {
      const uint32_t s = uint32_t(it) & 1u, ph = (uint32_t(it) >> 1) & 1u;
      if (it >= 2) mbar_wait(acc_empty(s), ph ^ 1u);  // the epilogue has drained this accumulator
      mbar_wait(a_full(s), ph);
      tc_fence_after();
      const uint32_t d = tmem + s * 256u;
      if (elect_one()) {
        uint32_t acc = 0;
        for (int ks = 0; ks < nks; ++ks) {
          const uint32_t a = base + kFtAOff + s * 4u * kFtATile + uint32_t(ks) * 2u * kFtATile;
          const uint32_t w = base + kFtWOff + uint32_t(ks) * 2u * kFtWTile;
          const uint64_t ah = ft_desc_sw32(a), al = ft_desc_sw32(a + kFtATile);
          const uint64_t wh = ft_desc_sw32(w), wl = ft_desc_sw32(w + kFtWTile);
          tc_mma_f16(d, ah, wl, idesc, acc);
          tc_mma_f16(d, al, wh, idesc, 1);
          tc_mma_f16(d, ah, wh, idesc, 1);
          acc = 1;
        }
        tc_commit(a_empty(s));
        tc_commit(acc_full(s));
      }
      __syncwarp();
    }
  } else {  // ================================================================= epilogue warps 0..7
    const int quad = warp & 3, half = warp >> 2;
    const int r = quad * 32 + lane;
    const uint32_t lane_addr = uint32_t(quad * 32) << 16;
    const float* col = reinterpret_cast<const float*>(sm + kFtColOff);
    float2* stats = reinterpret_cast<float2*>(sm + kFtStatOff);
    uint8_t* obuf = sm + kFtOutOff + uint32_t(warp) * 16384u;
    const uint32_t obuf_addr = base + kFtOutOff + uint32_t(warp) * 16384u;
    uint32_t stores = 0;
    for (int it = 0; it < ntiles; ++it) {
      const uint32_t s = uint32_t(it) & 1u, ph = (uint32_t(it) >> 1) & 1u;
      const int64_t m0 = int64_t(slot + it * P.ctas_per_pair) * 128;
      const int64_t row = m0 + r;
      const bool valid = row < P.M;
      const float* prow = Q.pos + (int64_t(valid ? row % P.T : 0) + 2) * kFtD + half * 128;
      mbar_wait(acc_full(s), ph);
      tc_fence_after();
      const uint32_t acc = tmem + s * 256u + lane_addr + uint32_t(half * 128);
      float v[32];
      // ---- pass 1: (dot + bias) + position row, as the reference adds them; shifted sums of the 128 columns
      float shift = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll 1
      for (int i = 0; i < 4; ++i) {
        tc_ld32(acc + uint32_t(32 * i), v);
        add_cols(v, col + half * 128 + 32 * i);
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 p = (valid && !(SCATT_FT_ABLATE && (P.ablate & 1))) ? __ldg(reinterpret_cast<const float4*>(prow + 32 * i + j)) : make_float4(0.f, 0.f, 0.f, 0.f);
          v[j] += p.x, v[j + 1] += p.y, v[j + 2] += p.z, v[j + 3] += p.w;
        }
        if (i == 0) shift = v[0];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float d = v[j] - shift;
          s1 += d;
          s2 = fmaf(d, d, s2);
        }
        tc_st32(acc + uint32_t(32 * i), v);
      }
      const float dm = s1 * (1.0f / 128.0f);
      const float my_mean = shift + dm, my_m2 = fmaxf(s2 - s1 * dm, 0.f);
      float2* st = stats + (it & 1) * 256;  // alternating buffers: a warp may be a whole tile ahead of its partner's read
      st[half * 128 + r] = make_float2(my_mean, my_m2);
      ft_epi_bar();
      const float2 other = st[(half ^ 1) * 128 + r];
      const float mean = 0.5f * (my_mean + other.x);
      const float da = my_mean - mean, db = other.x - mean;
      const float m2 = my_m2 + other.y + 128.0f * (da * da + db * db);  // Chan et al.
      const float rstd = rsqrtf(m2 * (1.0f / float(kFtD)) + 1e-5f);
      // ---- pass 2: normalise, split, 64 columns per TMA store (hi tile | lo tile of 32 rows x 128 bytes)
#pragma unroll 1
      for (int i = 0; i < 2; ++i) {
        const uint32_t buf = (stores & 1u) * 8192u;
        if (stores >= 2) {  // the box pair last written into this buffer has been read out
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          __syncwarp();
        }
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int cl = half * 128 + 64 * i + 32 * hh;
          tc_ld32(acc + uint32_t(64 * i + 32 * hh), v);
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 g = *reinterpret_cast<const float4*>(col + kFtD + cl + j);
            const float4 b = *reinterpret_cast<const float4*>(col + 2 * kFtD + cl + j);
            v[j] = (v[j] - mean) * rstd * g.x + b.x;
            v[j + 1] = (v[j + 1] - mean) * rstd * g.y + b.y;
            v[j + 2] = (v[j + 2] - mean) * rstd * g.z + b.z;
            v[j + 3] = (v[j + 3] - mean) * rstd * g.w + b.w;
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 hi, lo;
            split8<SCATT_PLANE_F16>(make_float4(v[8 * j], v[8 * j + 1], v[8 * j + 2], v[8 * j + 3]),
                                    make_float4(v[8 * j + 4], v[8 * j + 5], v[8 * j + 6], v[8 * j + 7]), hi, lo);
            const uint32_t off = uint32_t(lane * 128 + (((4 * hh + j) ^ (lane & 7)) << 4));
            *reinterpret_cast<uint4*>(obuf + buf + off) = hi;
            *reinterpret_cast<uint4*>(obuf + buf + 4096u + off) = lo;
          }
        }
        if (i == 1) {  // last TMEM read of this tile: the accumulator may be overwritten by tile it + 2
          tc_fence_before();
          mbar_arrive(acc_empty(s));
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0 && !(SCATT_FT_ABLATE && (P.ablate & 4))) {
          tma_store_3d(&P.map_out[pi], obuf_addr + buf, half * 128 + 64 * i, int(m0) + quad * 32, 0);  // rows past M are clipped
        }
        if (lane == 0) asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        ++stores;
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kFtMmaWarp) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

}  // namespace

// Returns SCATT_OK after launching, or a positive value when the request is outside this kernel's envelope
// (the caller then runs the CUDA-core kernel): fp32 outputs, the exact gathered copy, bf16 planes, tiny inputs.
int launch_frontend_tc(const float* kp, int B, int T, int K, const scatt_frontend_stream* streams, int n, int fmt, cudaStream_t s) {
  const int64_t M = int64_t(B) * T;
  if (fmt != SCATT_PLANE_F16 || M < 128 || M > 0x7fffffff - 128) return 1;
  for (int i = 0; i < n; ++i) {
    if (streams[i].gathered || streams[i].out[0] || streams[i].out[1] || !streams[i].out_planes[0] || !streams[i].out_planes[1]) return 1;
    if ((reinterpret_cast<uintptr_t>(streams[i].pos[0]) | reinterpret_cast<uintptr_t>(streams[i].pos[1])) & 15) return 1;
  }
  FtParams P{};
  P.kp = kp, P.M = M, P.T = T, P.K = K, P.npairs = 2 * n;
  P.tiles_m = int((M + 127) / 128);
  if (SCATT_FT_ABLATE) {
    const char* e = std::getenv("SCATT_FT_DBG");
    P.ablate = e ? std::atoi(e) : 0;
  }
  for (int i = 0; i < n; ++i)
    for (int br = 0; br < 2; ++br) {
      FtPair& q = P.pair[2 * i + br];
      q.joint_idx = streams[i].joint_idx, q.n_joints = streams[i].n_joints, q.coord = streams[i].coord[br];
      q.wt = streams[i].map_wt[br], q.bias = streams[i].map_b[br], q.pos = streams[i].pos[br];
      q.ln_g = streams[i].ln_g[br], q.ln_b = streams[i].ln_b[br];
      const int rc = encode_planes_map(&P.map_out[2 * i + br], streams[i].out_planes[br], M, kFtD, 32, fmt, 2);
      if (rc != SCATT_OK) return rc;
    }
  static PerDeviceFlag attr_done;
  if (!attr_done.load()) {
    SCATT_CUDA(cudaFuncSetAttribute(frontend_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kFtSmemBytes)));
    attr_done.store(true);
  }
  const int per = 148 / P.npairs;
  P.ctas_per_pair = P.tiles_m < per ? P.tiles_m : per;
  (void)launch_kernel(frontend_tc_kernel, dim3(P.ctas_per_pair * P.npairs), dim3(kFtThreads), kFtSmemBytes, s, P);
  return after_launch("frontend_tc_kernel");
}

}  // namespace scatt
