// Recurrent half of the bidirectional LSTM of AlignmentModule (reference
// model/alignment_module.py:25-33,66-72; the consumer of fuse_embed, scope row f-2).
//
// The input projections x_t W_ih^T + b_ih + b_hh of all time steps and both
// directions are one scatt_linear call (N = 2*4H); what is left is the strictly
// sequential part  g_t = gates_x[t] + h_{t-1} W_hh^T,  (c_t, h_t) = cell(g_t, c_{t-1}).
//
// One persistent co-operative kernel runs all T steps.  W_hh (4H x H fp32 = 4 MB
// per direction) never leaves the chip: 64 CTAs per direction each keep the 32 gate
// rows of 8 hidden units in shared memory (64.5 KB) for the whole sequence, so a
// step moves only h_{t-1} (16 KB per 8 sequences) through L2.  The h exchange uses
// flagged 8-byte words (value, step) written with one store and polled directly by
// the consumers - one L2 round trip per step instead of store + fence + atomic +
// poll + load.  With more than 8 sequences the batch is walked in chunks of 8 and
// the flagged words of the next chunk are fetched while the current chunk's
// products run (its h_{t-1} was published a whole step ago), which hides the round
// trip.  Arithmetic is fp32 FMA (the recurrence amplifies product error over T
// steps, and at 8 sequences per step the tensor cores would idle anyway).
//
// Measured on B200 (tools/time_alignment.py): see DESIGN.md section 9.  A variant that
// gathered h once per 8-CTA cluster and forwarded it through DSMEM behind
// barrier.cluster was 2x slower (the release/acquire cluster barrier costs ~2.3 us
// per step next to in-flight global stores) and was dropped.
#include <cstdlib>

#include "common.cuh"

namespace scatt {

namespace {

constexpr int kH = 512;                   // hidden units per direction
constexpr int kCtasPerDir = 64;           // CTAs sharing one direction
constexpr int kUnits = kH / kCtasPerDir;  // 8 hidden units per CTA
constexpr int kRows = 4 * kUnits;         // 32 gate rows (i, f, g, o of each unit)
constexpr int kWStride = kH + 4;          // padded row stride: conflict-free 128-bit reads across rows
constexpr int kBc = 8;                    // sequences per chunk
constexpr int kThreads = 256;
constexpr int kMaxBatch = 2048;           // cell state of the CTA's units lives in shared memory

struct LstmSmem {
  float w[kRows * kWStride];
  float h[2][kBc][kH];
  float part[kThreads / 32][kRows][kBc];
  float gate[kRows][kBc];
};

__device__ __forceinline__ uint4 ld_flagged(const uint4* p) {
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

__device__ __forceinline__ void st_flagged(uint2* p, float value, uint32_t flag) {
  asm volatile("st.volatile.global.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(__float_as_uint(value)), "r"(flag) : "memory");
}

__device__ __forceinline__ float sigmoid_acc(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float tanh_acc(float x) {
  // 1 - 2 / (e^{2x} + 1): absolute error at the fp32 rounding level over the whole range
  return 1.0f - __fdividef(2.0f, __expf(2.0f * x) + 1.0f);
}

// The flagged words one thread fetches per round: producer CTA tid>>2, sequences 2*(tid&3) and +1 of the
// chunk, 8 units each = 2 x 64 bytes.
struct Fetch {
  uint4 v[8];
};

// grid (64, 2): blockIdx.y = direction (0 forward, 1 reverse), blockIdx.x = slice of 8 hidden units.
// gates_x [B*T, ldg]: row b*T + t, columns dir*4H + gate*H + unit (torch gate order i, f, g, o).
// y [B*T, 2H] row b*T + t, columns dir*H + unit (forward | reverse, as nn.LSTM concatenates them).
// ll: flagged exchange buffer [dir][parity][B][H] of (value, step) words, zeroed before the launch.
__global__ void __launch_bounds__(kThreads, 1)
    lstm_bidir_kernel(const float* __restrict__ gates_x, int64_t ldg, const float* __restrict__ w_hh,
                      float* __restrict__ y, uint16_t* __restrict__ y_planes, uint2* __restrict__ ll, int B, int T,
                      int fmt) {
  extern __shared__ __align__(16) unsigned char lstm_smem_raw[];
  LstmSmem& S = *reinterpret_cast<LstmSmem*>(lstm_smem_raw);
  float* cell = reinterpret_cast<float*>(lstm_smem_raw + sizeof(LstmSmem));  // [B][kUnits]

  const int tid = threadIdx.x, lane = tid & 31, warp = scatt_warp_idx();
  const int dir = blockIdx.y, slice = blockIdx.x, j0 = slice * kUnits;

  // resident weights: local row r = gate*8 + unit  <-  W_hh[dir][gate*H + j0 + unit][:]
  for (int idx = tid; idx < kRows * (kH / 4); idx += kThreads) {
    const int r = idx / (kH / 4), c4 = idx % (kH / 4);
    const int64_t grow = int64_t(dir) * 4 * kH + (r / kUnits) * kH + j0 + (r % kUnits);
    const float4 v = __ldg(reinterpret_cast<const float4*>(w_hh + grow * kH) + c4);
    *reinterpret_cast<float4*>(&S.w[r * kWStride + c4 * 4]) = v;
  }
  for (int idx = tid; idx < B * kUnits; idx += kThreads) cell[idx] = 0.0f;
  __syncthreads();

  const int nchunks = (B + kBc - 1) / kBc;
  const int rounds = T * nchunks;   // round r = (step r / nchunks, chunk r % nchunks)
  const bool lead = nchunks >= 2;   // round r+1 does not depend on round r: fetch it one round ahead
  const int64_t plane_stride = int64_t(B) * T * (2 * kH);
  // matvec mapping: thread = (hidden unit mu with its 4 gate rows, k slice of 16)
  const int mu = tid & 7, ks = tid >> 3;
  // fetch mapping: thread = (producer CTA, pair of sequences)
  const int prod = tid >> 2, bq = (tid & 3) * 2;
  // gate mapping: thread = (local row, sequence)
  const int gr = tid >> 3, gb = tid & 7;

  // issue the loads of round r's h_{t-1} words; returns whether every flag already carries step t
  auto fetch = [&](int r, Fetch& f) -> bool {
    const int t = r / nchunks, b0 = (r - t * nchunks) * kBc;
    const uint32_t want = uint32_t(t);
    const uint2* base = ll + (int64_t(dir * 2 + ((t - 1) & 1)) * B + b0) * kH + prod * kUnits;
    const bool v0 = b0 + bq < B, v1 = b0 + bq + 1 < B;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      f.v[i] = v0 ? ld_flagged(reinterpret_cast<const uint4*>(base + int64_t(bq) * kH) + i) : make_uint4(0, want, 0, want);
      f.v[4 + i] = v1 ? ld_flagged(reinterpret_cast<const uint4*>(base + int64_t(bq + 1) * kH) + i) : make_uint4(0, want, 0, want);
    }
    bool ok = true;
#pragma unroll
    for (int i = 0; i < 8; ++i) ok = ok && f.v[i].y == want && f.v[i].w == want;
    return ok;
  };
  // spin until round r's words are all there, then lay them out as h[buffer][sequence][unit]
  auto land = [&](int r, Fetch& f, bool ok) {
    long long t_start = 0;
    for (int spin = 0; !ok; ++spin) {
      ok = fetch(r, f);
      if (spin == 64) t_start = clock64();
      if (spin > 64 && clock64() - t_start > 4000000000ll) __trap();  // ~2 s: a peer CTA never arrived
    }
    float(*hb)[kH] = S.h[r & 1];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      *reinterpret_cast<float2*>(&hb[bq][prod * kUnits + 2 * i]) = make_float2(__uint_as_float(f.v[i].x), __uint_as_float(f.v[i].z));
      *reinterpret_cast<float2*>(&hb[bq + 1][prod * kUnits + 2 * i]) =
          make_float2(__uint_as_float(f.v[4 + i].x), __uint_as_float(f.v[4 + i].z));
    }
  };

  Fetch f;
  bool landed = false;  // S.h[r & 1] already holds round r (fetched during round r - 1)
  for (int r = 0; r < rounds; ++r) {
    const int t = r / nchunks, b0 = (r - t * nchunks) * kBc;
    const int tt = dir ? T - 1 - t : t;
    // input-projection term of this thread's (row, sequence): issued first so its latency hides
    float gx = 0.0f;
    if (b0 + gb < B)
      gx = __ldg(gates_x + (int64_t(b0 + gb) * T + tt) * ldg + dir * 4 * kH + (gr / kUnits) * kH + j0 + (gr % kUnits));
    if (t > 0 && !landed) land(r, f, false);
    __syncthreads();  // S.h[r & 1] complete
    const bool ahead = lead && r + 1 < rounds && (r + 1) / nchunks > 0;
    bool ahead_ok = false;
    if (ahead) ahead_ok = fetch(r + 1, f);  // in flight during the products below

    // partial products: the 4 gates of unit mu x 8 sequences over this thread's 16 k
    float acc[4][kBc];
#pragma unroll
    for (int g = 0; g < 4; ++g)
#pragma unroll
      for (int b = 0; b < kBc; ++b) acc[g][b] = 0.0f;
    if (t > 0) {
      const float(*hb)[kH] = S.h[r & 1];
#pragma unroll
      for (int k4 = 0; k4 < 4; ++k4) {
        float4 wv[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) wv[g] = *reinterpret_cast<const float4*>(&S.w[(g * kUnits + mu) * kWStride + ks * 16 + 4 * k4]);
#pragma unroll
        for (int b = 0; b < kBc; ++b) {
          const float4 hv = *reinterpret_cast<const float4*>(&hb[b][ks * 16 + 4 * k4]);
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            acc[g][b] = fmaf(wv[g].x, hv.x, acc[g][b]), acc[g][b] = fmaf(wv[g].y, hv.y, acc[g][b]);
            acc[g][b] = fmaf(wv[g].z, hv.z, acc[g][b]), acc[g][b] = fmaf(wv[g].w, hv.w, acc[g][b]);
          }
        }
      }
    }
    // the warp's 4 k slices by reduce-scatter (lane bits 3-4): the lane ends with gate lane>>3 of unit
    // lane&7, i.e. local row == lane; then the 8 warps through shared memory
    {
      const bool up16 = lane & 16, up8 = lane & 8;
      float fin[kBc];
#pragma unroll
      for (int b = 0; b < kBc; ++b) {
        const float k0 = (up16 ? acc[2][b] : acc[0][b]) + __shfl_xor_sync(0xffffffffu, up16 ? acc[0][b] : acc[2][b], 16);
        const float k1 = (up16 ? acc[3][b] : acc[1][b]) + __shfl_xor_sync(0xffffffffu, up16 ? acc[1][b] : acc[3][b], 16);
        fin[b] = (up8 ? k1 : k0) + __shfl_xor_sync(0xffffffffu, up8 ? k0 : k1, 8);
      }
      float* dst = &S.part[warp][lane][0];
      *reinterpret_cast<float4*>(dst) = make_float4(fin[0], fin[1], fin[2], fin[3]);
      *reinterpret_cast<float4*>(dst + 4) = make_float4(fin[4], fin[5], fin[6], fin[7]);
    }
    __syncthreads();
    {
      float g = gx;
#pragma unroll
      for (int w = 0; w < kThreads / 32; ++w) g += S.part[w][gr][gb];
      S.gate[gr][gb] = g;
    }
    __syncthreads();

    // cell update of this CTA's 8 units x 8 sequences and publication of h_t
    if (tid < kUnits * kBc) {
      const int u = tid & (kUnits - 1), b = tid / kUnits;
      if (b0 + b < B) {
        const float gi = sigmoid_acc(S.gate[u][b]);
        const float gf = sigmoid_acc(S.gate[kUnits + u][b]);
        const float gg = tanh_acc(S.gate[2 * kUnits + u][b]);
        const float go = sigmoid_acc(S.gate[3 * kUnits + u][b]);
        const float cn = fmaf(gf, cell[(b0 + b) * kUnits + u], gi * gg);
        cell[(b0 + b) * kUnits + u] = cn;
        const float hn = go * tanh_acc(cn);
        if (t + 1 < T) st_flagged(ll + (int64_t(dir * 2 + (t & 1)) * B + b0 + b) * kH + j0 + u, hn, uint32_t(t + 1));
        const int64_t off = (int64_t(b0 + b) * T + tt) * (2 * kH) + dir * kH + j0 + u;
        if (y) y[off] = hn;
        if (y_planes) {
          uint16_t hi, lo;
          split16_rt(hn, fmt, hi, lo);
          y_planes[off] = hi;
          y_planes[plane_stride + off] = lo;
        }
      }
    }
    // the other buffer was last read by round r - 1, which every thread left two barriers ago
    landed = ahead;
    if (ahead) land(r + 1, f, ahead_ok);
    // S.part / S.gate are rewritten only behind the next round's barriers
  }
}

}  // namespace

size_t lstm_workspace_bytes(int64_t B, int H) { return size_t(2) * 2 * size_t(B) * size_t(H) * sizeof(uint2); }

int launch_lstm_bidir(const float* gates_x, int64_t ldg, const float* w_hh, float* y, void* y_planes, void* workspace,
                      int64_t B, int T, int H, int fmt, cudaStream_t s) {
  SCATT_REQUIRE(H == kH, "lstm_bidir: hidden size per direction must be %d (got %d)", kH, H);
  SCATT_REQUIRE(B >= 0 && B <= kMaxBatch, "lstm_bidir: batch %lld exceeds %d sequences per call", (long long)B, kMaxBatch);
  SCATT_REQUIRE(ldg >= 8 * kH, "lstm_bidir: gates_x rows must hold 2 x 4H = %d columns", 8 * kH);
  SCATT_REQUIRE(gates_x && w_hh && workspace && (y || y_planes), "lstm_bidir: null argument");
  if (B == 0 || T == 0) return SCATT_OK;
  const size_t smem = sizeof(LstmSmem) + size_t(B) * kUnits * sizeof(float);
  SCATT_REQUIRE(smem <= 227 * 1024, "lstm_bidir: B=%lld needs %zu bytes of shared memory", (long long)B, smem);
  static PerDeviceFlag configured;  // the attribute belongs to the device's context: once per device, to the limit
  if (!configured.load()) {
    SCATT_CUDA(cudaFuncSetAttribute(lstm_bidir_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured.store(true);
  }
  SCATT_CUDA(cudaMemsetAsync(workspace, 0, lstm_workspace_bytes(B, H), s));
  // all 128 CTAs spin on each other: the launch must be co-resident (co-operative), and is not programmatic
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(kCtasPerDir, 2), cfg.blockDim = dim3(kThreads), cfg.dynamicSmemBytes = smem, cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr, cfg.numAttrs = 1;
  const int Bi = int(B);
  (void)cudaLaunchKernelEx(&cfg, lstm_bidir_kernel, gates_x, ldg, w_hh, y, reinterpret_cast<uint16_t*>(y_planes),
                           reinterpret_cast<uint2*>(workspace), Bi, T, fmt);
  return after_launch("lstm_bidir_kernel");
}

}  // namespace scatt
