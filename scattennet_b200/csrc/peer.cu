// Logits all-gather over NVLink peer memory (scope row e: the one exchange step of the path).
//
// Every rank owns a symmetric buffer [world][shard] that its peers have mapped (CUDA VMM handles exchanged by
// torch.distributed's symmetric memory); rank r PUSHES its shard into slot r of every peer's buffer with plain
// 16-byte stores - NVSwitch gives each GPU full bandwidth to every peer, so the W - 1 copies proceed at once -
// and the same kernel runs the barrier that tells the consumers every slot has landed:
//
//   all CTAs    copy src -> peer[p] + rank * shard   (p = 0 .. W-1, own slot included), __threadfence_system
//   last CTA    flag[rank] := seq on every peer (st.release.sys), then spins until flag[p] >= seq for all p
//               on its own pad (ld.acquire.sys)
//
// `seq` grows by one per call and the callers alternate between two buffers (seq parity): a rank can overwrite
// a buffer only after the barrier of the call in between, which every peer enters after it has consumed that
// buffer in stream order.  Against NCCL's ring all-gather this removes the (W - 1) dependent hops: the step
// costs one NVLink round trip plus shard / link-bandwidth whatever W is.
#include "common.cuh"

namespace scatt {

namespace {

struct PeerPtrs {
  uint8_t* buf[SCATT_MAX_PEERS];       // peer p's gather buffer (this call's parity)
  uint64_t* flags[SCATT_MAX_PEERS];    // peer p's flag pad: uint64[world]
};

__device__ __forceinline__ void st_release_sys(uint64_t* p, uint64_t v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t ld_acquire_sys(const uint64_t* p) {
  uint64_t v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(256) peer_allgather_kernel(const uint8_t* __restrict__ src, int64_t bytes, PeerPtrs pp, int world,
                                                            int rank, unsigned int* __restrict__ counter, uint64_t seq) {
  const int64_t nvec = bytes >> 4;
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  const uint4* s4 = reinterpret_cast<const uint4*>(src);
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    const uint4 v = s4[i];
#pragma unroll 1
    for (int q = 0; q < world; ++q) {
      const int p = (rank + q) % world;  // start with the own slot, then walk the peers from a rank-dependent offset
      reinterpret_cast<uint4*>(pp.buf[p] + int64_t(rank) * bytes)[i] = v;
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const bool last = atomicAdd(counter, 1u) == gridDim.x - 1;
    if (last) {
      __threadfence_system();
      *counter = 0;  // ready for the next call (stream order)
      for (int p = 0; p < world; ++p) st_release_sys(pp.flags[p] + rank, seq);
      for (int p = 0; p < world; ++p) {
        uint32_t spins = 0;
        while (ld_acquire_sys(pp.flags[rank] + p) < seq) {
          if (++spins > (1u << 28)) __trap();  // a peer never arrived: fail the launch instead of hanging the GPU
          __nanosleep(64);
        }
      }
    }
  }
}

}  // namespace

int launch_peer_allgather(const void* src, int64_t bytes, void* const* peer_bufs, void* const* peer_flags, int world, int rank,
                          void* counter, uint64_t seq, cudaStream_t s) {
  SCATT_REQUIRE(world >= 1 && world <= SCATT_MAX_PEERS && rank >= 0 && rank < world, "peer_allgather: world 1..%d, rank inside it",
                SCATT_MAX_PEERS);
  SCATT_REQUIRE(src && peer_bufs && peer_flags && counter && bytes % 16 == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0,
                "peer_allgather: null pointer or a shard that is not a multiple of 16 bytes");
  PeerPtrs pp{};
  for (int p = 0; p < world; ++p) {
    SCATT_REQUIRE(peer_bufs[p] && peer_flags[p], "peer_allgather: null pointer for peer %d", p);
    pp.buf[p] = reinterpret_cast<uint8_t*>(peer_bufs[p]);
    pp.flags[p] = reinterpret_cast<uint64_t*>(peer_flags[p]);
  }
  int64_t blocks = ((bytes >> 4) + 256 * 4 - 1) / (256 * 4);
  if (blocks < 1) blocks = 1;
  if (blocks > 64) blocks = 64;  // a few dozen CTAs saturate the NVLink ports; the rest of the GPU stays free
  // plain stream order (no programmatic dependent launch): the shard is usually the output of a whole CUDA-graph
  // launch, and a barrier kernel must not become resident before its producer has finished everywhere
  peer_allgather_kernel<<<dim3(unsigned(blocks)), dim3(256), 0, s>>>(reinterpret_cast<const uint8_t*>(src), bytes, pp, world, rank,
                                                                    reinterpret_cast<unsigned int*>(counter), seq);
  return after_launch("peer_allgather_kernel");
}

}  // namespace scatt
