// Micro-benchmarks behind the design of block_tc.cu (dev tool; nvcc -gencode arch=compute_100a,code=sm_100a):
//   (a) L2 -> shared memory streaming rate of ONE SM through TMA as a function of the bytes in flight
//       (ring of S slots of [rows x 64] 16-bit tiles, 128-byte swizzle) and of how many SMs stream at once;
//   (b) tcgen05.mma issue / execution rate for M = 128, K = 16, N = 128 | 256, A from shared memory (SS) or
//       tensor memory (TS), operands resident (no loads).
// Prints one line per configuration: bytes / clk / SM, cycles per MMA.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../scattennet_b200/csrc/tc_ptx.cuh"

using namespace scatt::tc;

#define CK(x)                                                                         \
  do {                                                                                \
    cudaError_t e = (x);                                                              \
    if (e != cudaSuccess) {                                                           \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      exit(1);                                                                        \
    }                                                                                 \
  } while (0)

struct alignas(64) StreamParams {
  CUtensorMap map;
  int slots, box_rows, tiles, rows_total, k_total;
  int nprod, planes, bulk1d;  // producer warps; planes per box (3rd box dimension); 1: cp.async.bulk of contiguous bytes instead of a tensor box
  const void* src;
  long long* out;  // [grid] cycles
};

// (a): warp 0 produces, warp 1 consumes (waits full, arrives empty - no MMA)
__global__ void __launch_bounds__(160, 1) tma_stream_kernel(const __grid_constant__ StreamParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw), base = (raw + 1023u) & ~1023u;
  const uint32_t slot_bytes = uint32_t(P.box_rows) * 128u * uint32_t(P.planes);
  const uint32_t bar0 = base + uint32_t(P.slots) * slot_bytes;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int s = 0; s < P.slots; ++s) {
      mbar_init(bar0 + 8u * s, 1);
      mbar_init(bar0 + 8u * (P.slots + s), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&P.map) : "memory");
  }
  __syncthreads();
  const long long t0 = clock64();
  const int kbs = P.k_total / 64, rbs = P.rows_total / P.box_rows;
  if (warp >= 1) {
    if (warp - 1 < P.nprod) {
      for (int it = warp - 1; it < P.tiles; it += P.nprod) {
        const int s = it % P.slots;
        mbar_wait(bar0 + 8u * (P.slots + s), ((it / P.slots) & 1) ^ 1);
        if (elect_one()) {
          mbar_expect_tx(bar0 + 8u * s, slot_bytes);
          // every CTA walks the same tiles in the same order (like the row tiles of one weight matrix), offset by blockIdx
          const int tile = (it + blockIdx.x * 7) % (kbs * rbs);
          if (P.bulk1d) {
            const char* g = reinterpret_cast<const char*>(P.src) + size_t(tile) * slot_bytes;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(base + s * slot_bytes),
                         "l"(g), "r"(slot_bytes), "r"(bar0 + 8u * s)
                         : "memory");
          } else {
            tma_load_3d(base + s * slot_bytes, &P.map, bar0 + 8u * s, (tile % kbs) * 64, (tile / kbs) * P.box_rows, 0);
          }
        }
        __syncwarp();
      }
    }
  } else {
    for (int it = 0; it < P.tiles; ++it) {
      const int s = it % P.slots;
      mbar_wait(bar0 + 8u * s, (it / P.slots) & 1);
      if (elect_one()) mbar_arrive(bar0 + 8u * (P.slots + s));
      __syncwarp();
    }
    if (threadIdx.x == 0) P.out[blockIdx.x] = clock64() - t0;
  }
}

// (b): one warp issues n_mma MMAs on resident operands; a commit per `group` MMAs is waited for at the end only
struct MmaParams {
  int n_mma, N, ts, group;
  long long* out;  // [0] cycles until the last issue returned, [1] until the last commit arrived
};
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(MmaParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw), base = (raw + 1023u) & ~1023u;
  const uint32_t a_addr = base, b_addr = base + 16384, bar = base + 16384 + 32768, tptr = bar + 16;
  const int warp = threadIdx.x >> 5;
  for (uint32_t i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw + (base - raw))[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(bar + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tptr), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem_raw + (tptr - raw));
  if (warp == 0) {
    const uint32_t idesc = (1u << 4) | (uint32_t(P.N >> 3) << 17) | (uint32_t(128 >> 4) << 24);
    const uint64_t ad = umma_desc_sw128(a_addr), bd = umma_desc_sw128(b_addr);
    const long long t0 = clock64();
    if (elect_one()) {
      for (int i = 0; i < P.n_mma; ++i) {
        const uint64_t adv = uint64_t((i & 3) * 2);
        if (P.ts) {
          asm volatile(
              "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
              "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem),
              "r"(tmem + 256u + uint32_t(i & 3) * 8u), "l"(bd + adv), "r"(idesc), "r"(1u)
              : "memory");
        } else {
          tc_mma_f16(tmem, ad + adv, bd + adv, idesc, 1);
        }
        if ((i + 1) % P.group == 0 && i + 1 < P.n_mma) tc_commit(bar + 8);  // a commit per group, like a ring release (never waited)
      }
      tc_commit(bar);
    }
    __syncwarp();
    const long long t1 = clock64();
    mbar_wait(bar, 0);
    const long long t2 = clock64();
    if (threadIdx.x == 0) P.out[0] = t1 - t0, P.out[1] = t2 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                              const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
  EncodeFn enc = reinterpret_cast<EncodeFn>(fp);
  long long* out;
  CK(cudaMalloc(&out, 1024 * sizeof(long long)));
  std::vector<long long> host(1024);
  CK(cudaFuncSetAttribute(tma_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  CK(cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));

  // ---- (a) weight-like split planes [2][rows][K] 16-bit, K = 256 (row stride 512 B)
  {
    const int K = 256, rows = 768 * 3;
    void* w;
    CK(cudaMalloc(&w, size_t(2) * rows * K * 2));
    CK(cudaMemset(w, 0, size_t(2) * rows * K * 2));
    struct Cfg { int box_rows, planes, bulk; };
    for (Cfg c : {Cfg{64, 1, 0}, Cfg{128, 1, 0}, Cfg{256, 1, 0}, Cfg{128, 2, 0}, Cfg{256, 2, 0}, Cfg{128, 1, 1}, Cfg{256, 1, 1}}) {
      CUtensorMap map;
      const cuuint64_t dims[3] = {cuuint64_t(K), cuuint64_t(rows), 2};
      const cuuint64_t strides[2] = {cuuint64_t(K) * 2, cuuint64_t(rows) * K * 2};
      const cuuint32_t box[3] = {64, cuuint32_t(c.box_rows), cuuint32_t(c.planes)}, estr[3] = {1, 1, 1};
      if (enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, w, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
        printf("encode failed\n");
        return 1;
      }
      for (int grid : {39, 148}) {
        for (int nprod : {1, 2, 4}) {
          for (int slots : {2, 4, 8}) {
            const int slot_bytes = c.box_rows * 128 * c.planes;
            if (slots * slot_bytes > 200 * 1024 || slots < nprod) continue;
            StreamParams P{};
            P.map = map, P.slots = slots, P.box_rows = c.box_rows, P.tiles = 400, P.rows_total = rows, P.k_total = K, P.out = out;
            P.nprod = nprod, P.planes = c.planes, P.bulk1d = c.bulk, P.src = w;
            if (c.bulk) P.rows_total = rows * K * 2 * 2 / slot_bytes, P.k_total = 64;  // tiles = contiguous chunks
            const size_t smem = size_t(slots) * slot_bytes + 1024 + 16 * slots + 64;
            for (int rep = 0; rep < 2; ++rep) tma_stream_kernel<<<grid, 160, smem>>>(P);
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(host.data(), out, grid * sizeof(long long), cudaMemcpyDeviceToHost));
            double avg = 0;
            for (int i = 0; i < grid; ++i) avg += double(host[i]) / grid;
            printf("tma_stream %s box=%dx64x%d (%2d KB) ctas=%3d producers=%d slots=%d: %6.1f B/clk/SM, %5.0f cyc per op\n",
                   c.bulk ? "bulk-1d" : "tensor ", c.box_rows, c.planes, slot_bytes / 1024, grid, nprod, slots,
                   double(P.tiles) * slot_bytes / avg, avg / P.tiles);
          }
        }
      }
    }
    CK(cudaFree(w));
  }
  // ---- (b)
  for (int N : {128, 256}) {
    for (int ts : {0, 1}) {
      for (int group : {4, 12, 1000000}) {
        MmaParams P{};
        P.n_mma = 240, P.N = N, P.ts = ts, P.group = group, P.out = out;
        for (int rep = 0; rep < 2; ++rep) mma_rate_kernel<<<1, 128, 60 * 1024>>>(P);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(host.data(), out, 2 * sizeof(long long), cudaMemcpyDeviceToHost));
        printf("mma_rate N=%d A=%s commit every %d: issue %5.1f cyc/MMA, complete %5.1f cyc/MMA (floor %d)\n", N, ts ? "tmem" : "smem",
               group > 1000 ? 0 : group, double(host[0]) / P.n_mma, double(host[1]) / P.n_mma, N / 2);
      }
    }
  }
  return 0;
}
