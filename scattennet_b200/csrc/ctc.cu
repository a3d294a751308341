// CTC prefix beam search on the per-frame logits (scope row f-4): the device-side replacement of the
// reference's utils.py:164-189 ctc_decode, which ships the logits to the host and runs TensorFlow's
// ctc_beam_search_decoder (beam 5, top path) there.  Same algorithm as TensorFlow's
// ctc_beam_search.h Step()/TopPaths() with merge_repeated = false and the default scorer, restated from
// its published description (TensorFlow itself is not available here: see oracle/ctc_oracle.py), followed
// by the reference's own post-processing: blank = class 0, ids in the original numbering, consecutive
// duplicates collapsed (itertools.groupby).
//
// One CTA per sequence, ONE block-wide barrier per time step.  The beam (<= W <= 16 live prefixes) lives in
// registers, replicated in every warp: lane i holds prefix i (trie node, parent node, last label, parent's last
// label, log-probabilities total / blank / label), and every warp advances and re-selects it redundantly with
// warp shuffles - no thread-0 serial section, no hand-over of the beam through shared memory.  A step is:
//   * the logits of frame t+1 are loaded into registers at the top of the step and staged (raw) into the other
//     half of a double buffer at its end, together with per-warp (max, sum exp) partials of the frame's softmax
//     normaliser - the normaliser of frame t is combined from the 8 partials after the barrier of step t-1,
//     so the global loads and the softmax reduction of the next frame run under this frame's search;
//   * advance: lane i finds its parent among the live prefixes by shuffles and updates (blank, label, total);
//   * all threads score the W * (V - 1) one-label extensions against the W-th best live prefix (an extension
//     below it can never enter the beam); a whole branch is skipped when its best possible extension (largest
//     non-blank log-probability of the frame + the branch's probability) is below that bar - the usual case once
//     the beam is full and the frame is a blank; each warp keeps its best W in registers (lane j holds the
//     j-th; rare insertions are serialised by ballot);
//   * barrier; every warp merges the 8 warp lists with the live prefixes into the next beam: W rounds of
//     arg-max by redux.sync on an order-preserving integer image of the score, ties to the lower insertion code
//     (order: probability, ties to the earlier insertion - the sequential push / pop-bottom of the original
//     reduces to exactly this top-W selection because an extension can never beat the prefix it extends).
//     Winners come out in descending order, so the beam stays sorted and the most probable prefix is lane 0.
// The prefix trie (parent, label per node; at most 1 + W * T nodes) lives in shared memory, written by warp 0
// and read only by the final back-walk; only token ids leave the chip.
#include <cfloat>

#include "common.cuh"

namespace scatt {

namespace {

constexpr int kCtcThreads = 256;
constexpr int kCtcWarps = kCtcThreads / 32;
constexpr int kMaxBeam = 16;
constexpr int kRowRegs = 8;  // logits of the next frame held in registers per thread (V <= 2048; the rest is re-read)
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ float lse2(float a, float b) {
  if (a == -INFINITY) return b;
  if (b == -INFINITY) return a;
  return fmaxf(a, b) + log1pf(expf(-fabsf(a - b)));
}

// (score, order): higher score first, then lower order (earlier insertion)
__device__ __forceinline__ bool better(float s0, int o0, float s1, int o1) { return s0 > s1 || (s0 == s1 && o0 < o1); }

// order-preserving map float -> uint32 (no NaNs here; -0 is folded into +0 first so that equal scores get equal keys)
__device__ __forceinline__ uint32_t order_key(float s) {
  const uint32_t b = __float_as_uint(s + 0.f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float order_key_inv(uint32_t k) { return __uint_as_float((k & 0x80000000u) ? (k ^ 0x80000000u) : ~k); }

// softmax partials (m = max, s = sum exp(x - m)) of two disjoint sets
__device__ __forceinline__ void lse_merge(float& m, float& s, float m2, float s2) {
  const float mm = fmaxf(m, m2);
  const float e1 = m == -INFINITY ? 0.f : expf(m - mm), e2 = m2 == -INFINITY ? 0.f : expf(m2 - mm);
  s = s * e1 + s2 * e2;
  m = mm;
}

__global__ void __launch_bounds__(kCtcThreads) ctc_beam_kernel(const float* __restrict__ logits, int T, int V,
                                                               const int* __restrict__ lengths, int W, int node_cap,
                                                               int* __restrict__ out_ids, int* __restrict__ out_len,
                                                               float* __restrict__ out_score) {
  extern __shared__ __align__(16) unsigned char ctc_smem[];
  const int Vp = (V + 3) & ~3;
  float* raw = reinterpret_cast<float*>(ctc_smem);                           // [2][Vp] logits of frame t / t+1
  float* part = raw + 2 * Vp;                                                // [2][3][warps]: max, sum exp, non-blank max
  int* node_parent = reinterpret_cast<int*>(part + 2 * 3 * kCtcWarps);       // [node_cap]
  int* node_label = node_parent + node_cap;                                  // [node_cap]
  float* cand_score = reinterpret_cast<float*>(node_label + node_cap);       // [2][warps * W] per-warp best extensions
  int* cand_code = reinterpret_cast<int*>(cand_score + 2 * kCtcWarps * W);   // [2][warps * W]

  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int len = lengths ? min(max(lengths[b], 0), T) : T;
  const float* x = logits + int64_t(b) * T * V;
  const int n_lab = V - 1;  // label k of the search = class k + 1 (class 0 is the blank)

  // stage a frame held in registers (+ the tail past kRowRegs * 256 classes, re-read) and its softmax partials
  auto stage = [&](const float (&nx)[kRowRegs], const float* row, int buf) {
    float* dst = raw + buf * Vp;
    float m = -INFINITY, mnb = -INFINITY;
#pragma unroll
    for (int r = 0; r < kRowRegs; ++r) {
      const int i = tid + r * kCtcThreads;
      if (i < V) {
        dst[i] = nx[r];
        m = fmaxf(m, nx[r]);
        if (i > 0) mnb = fmaxf(mnb, nx[r]);
      }
    }
    for (int i = tid + kRowRegs * kCtcThreads; i < V; i += kCtcThreads) {
      const float v = row[i];
      dst[i] = v;
      m = fmaxf(m, v), mnb = fmaxf(mnb, v);
    }
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < kRowRegs; ++r)
      if (tid + r * kCtcThreads < V) s += expf(nx[r] - m);
    for (int i = tid + kRowRegs * kCtcThreads; i < V; i += kCtcThreads) s += expf(row[i] - m);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float m2 = __shfl_xor_sync(kFull, m, o), s2 = __shfl_xor_sync(kFull, s, o);
      lse_merge(m, s, m2, s2);
      mnb = fmaxf(mnb, __shfl_xor_sync(kFull, mnb, o));
    }
    if (lane == 0) {
      float* pp = part + buf * 3 * kCtcWarps;
      pp[warp] = m, pp[kCtcWarps + warp] = s, pp[2 * kCtcWarps + warp] = mnb;
    }
  };
  auto fetch = [&](float (&nx)[kRowRegs], const float* row) {
#pragma unroll
    for (int r = 0; r < kRowRegs; ++r) {
      const int i = tid + r * kCtcThreads;
      nx[r] = i < V ? __ldg(row + i) : 0.f;
    }
  };

  // the beam, replicated per warp: lane i = live prefix i (sorted by descending probability)
  int n = 1, next_node = 1;
  int node = lane == 0 ? 0 : -2, par = -1, lab = -1, plab = -1;
  float total = lane == 0 ? 0.f : -INFINITY, blank = total, label = -INFINITY;
  if (tid == 0) node_parent[0] = -1, node_label[0] = -1;
  if (len > 0) {
    float nx[kRowRegs];
    fetch(nx, x);
    stage(nx, x, 0);
  }
  __syncthreads();

  for (int t = 0; t < len; ++t) {
    const int buf = t & 1;
    const float* lraw = raw + buf * Vp;
    const bool more = t + 1 < len;
    const float* nrow = x + int64_t(t + 1) * V;
    float nx[kRowRegs];
    if (more) fetch(nx, nrow);

    // ---- normaliser of the frame from the 8 per-warp partials (fixed order: every thread gets the same bits)
    float norm, mlp;
    {
      const float* pp = part + buf * 3 * kCtcWarps;
      float m = pp[0], s = pp[kCtcWarps], mnb = pp[2 * kCtcWarps];
#pragma unroll
      for (int w = 1; w < kCtcWarps; ++w) {
        lse_merge(m, s, pp[w], pp[kCtcWarps + w]);
        mnb = fmaxf(mnb, pp[2 * kCtcWarps + w]);
      }
      norm = m + logf(s);
      mlp = mnb - norm;  // largest log-probability of a label (rounding is monotonic: no logp below exceeds it)
    }

    // ---- the live prefixes at t: lane i looks its parent up among them
    const float ot = total, ob = blank;
    int fj = -1;
    float prev = -INFINITY;
    for (int j = 0; j < n; ++j) {
      const int nj = __shfl_sync(kFull, node, j);
      const float otj = __shfl_sync(kFull, ot, j), obj = __shfl_sync(kFull, ob, j);
      // the parent prefix is live: paths that reach this prefix from it at t
      if (lane < n && par >= 0 && nj == par) fj = j, prev = (lab == plab) ? obj : otj;
    }
    if (lane < n) {
      if (par >= 0) {
        float nl = label;
        if (fj >= 0) nl = lse2(nl, prev);
        label = nl + (lraw[lab + 1] - norm);
      }
      blank = ot + (lraw[0] - norm);
      total = lse2(blank, label);
    }
    // extension codes (branch * (V-1) + label) whose prefix is already live
    const int forbidden = (lane < n && fj >= 0) ? fj * n_lab + lab : -1;
    int forb[kMaxBeam];
#pragma unroll
    for (int j = 0; j < kMaxBeam; ++j) forb[j] = __shfl_sync(kFull, forbidden, j);
    // an extension must beat the W-th best live prefix (ties lose: live prefixes were inserted first)
    float thr_s = -INFINITY;
    if (n == W) {
      float lo = lane < n ? total : INFINITY;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) lo = fminf(lo, __shfl_xor_sync(kFull, lo, o));
      thr_s = lo;
    }
    int thr_c = -1;

    // ---- one-label extensions.  Warp-level top-W in registers: lane j holds the warp's j-th best (ls, lc).
    float ls = -INFINITY;
    int lc = 0x7fffffff;
    for (int bi = 0; bi < n; ++bi) {
      const int lab_b = __shfl_sync(kFull, lab, bi);
      const float otb = __shfl_sync(kFull, ot, bi), obb = __shfl_sync(kFull, ob, bi);
      if (fmaxf(otb, obb) + mlp < thr_s) continue;  // no extension of this branch reaches the bar (warp-uniform)
      for (int k0 = 0; k0 < n_lab; k0 += kCtcThreads) {
        const int k = k0 + tid;
        float sc = -INFINITY;
        int cd = 0x7fffffff;
        bool pending = false;
        if (k < n_lab) {
          cd = bi * n_lab + k;
          sc = (lraw[k + 1] - norm) + (k == lab_b ? obb : otb);
          pending = sc != -INFINITY && better(sc, cd, thr_s, thr_c);
          if (pending) {
            bool live_child = false;
#pragma unroll
            for (int j = 0; j < kMaxBeam; ++j) live_child |= forb[j] == cd;
            if (live_child) pending = false, sc = -INFINITY;  // already advanced above
          }
        }
        unsigned mask = __ballot_sync(kFull, pending);
        while (mask) {
          const int leader = __ffs(mask) - 1;
          const float bsc = __shfl_sync(kFull, sc, leader);
          const int bcd = __shfl_sync(kFull, cd, leader);
          // sorted insert: elements better than the newcomer form a prefix of the lanes
          const int pos = __popc(__ballot_sync(kFull, better(ls, lc, bsc, bcd)));
          const float up_s = __shfl_up_sync(kFull, ls, 1);
          const int up_c = __shfl_up_sync(kFull, lc, 1);
          if (lane == pos) ls = bsc, lc = bcd;
          else if (lane > pos) ls = up_s, lc = up_c;
          if (lane >= W) ls = -INFINITY, lc = 0x7fffffff;
          const float ws = __shfl_sync(kFull, ls, W - 1);
          const int wc = __shfl_sync(kFull, lc, W - 1);
          if (ws != -INFINITY && better(ws, wc, thr_s, thr_c)) thr_s = ws, thr_c = wc;  // list full: its tail is the bar
          if (lane == leader) pending = false;
          else pending = pending && better(sc, cd, thr_s, thr_c);
          mask = __ballot_sync(kFull, pending);
        }
      }
    }
    const int n_items = kCtcWarps * W;
    float* cs = cand_score + buf * n_items;
    int* cc = cand_code + buf * n_items;
    if (lane < W) cs[warp * W + lane] = ls, cc[warp * W + lane] = lc;
    if (more) stage(nx, nrow, buf ^ 1);
    __syncthreads();

    // ---- next beam: the best W of live prefixes + the warps' lists.  Lane-local items: its live prefix (code
    // i - kMaxBeam: live prefixes were inserted first) and candidates lane, lane + 32, ... of the 8 W entries
    float is[1 + kCtcWarps * kMaxBeam / 32];
    int ic[1 + kCtcWarps * kMaxBeam / 32];
    is[0] = lane < n ? total : -INFINITY, ic[0] = lane < n ? lane - kMaxBeam : 0x7fffffff;
#pragma unroll
    for (int r = 0; r < kCtcWarps * kMaxBeam / 32; ++r) {
      const int i = lane + 32 * r;
      is[r + 1] = i < n_items ? cs[i] : -INFINITY, ic[r + 1] = i < n_items ? cc[i] : 0x7fffffff;
    }
    float win_s = -INFINITY;
    int win_c = 0x7fffffff, n_win = 0;
    for (int round = 0; round < W; ++round) {
      float bs = is[0];
      int bc = ic[0];
#pragma unroll
      for (int r = 1; r <= kCtcWarps * kMaxBeam / 32; ++r)
        if (better(is[r], ic[r], bs, bc)) bs = is[r], bc = ic[r];
      const uint32_t key = order_key(bs);
      const uint32_t kmax = __reduce_max_sync(kFull, key);
      if (kmax == order_key(-INFINITY)) break;  // fewer than W items exist
      const int cmin = __reduce_min_sync(kFull, key == kmax ? bc : 0x7fffffff);
#pragma unroll
      for (int r = 0; r <= kCtcWarps * kMaxBeam / 32; ++r)
        if (ic[r] == cmin) is[r] = -INFINITY;  // taken (codes of valid items are unique)
      if (lane == n_win) win_s = order_key_inv(kmax), win_c = cmin;
      ++n_win;
    }

    // ---- rebuild: lane w becomes winner w; live prefixes keep (blank, label), extensions become new trie nodes
    {
      const bool mine = lane < n_win;
      const bool is_ext = mine && win_c >= 0;
      int bi = 0, k = 0;
      if (is_ext) bi = win_c / n_lab, k = win_c - bi * n_lab;
      const int src = mine ? (win_c < 0 ? win_c + kMaxBeam : bi) : 0;
      const int s_node = __shfl_sync(kFull, node, src), s_par = __shfl_sync(kFull, par, src);
      const int s_lab = __shfl_sync(kFull, lab, src), s_plab = __shfl_sync(kFull, plab, src);
      const float s_blank = __shfl_sync(kFull, blank, src), s_label = __shfl_sync(kFull, label, src);
      const unsigned ext_mask = __ballot_sync(kFull, is_ext);
      if (is_ext) {
        const int nd = min(next_node + __popc(ext_mask & ((1u << lane) - 1u)), node_cap - 1);  // cap is 1 + W * T: never exceeded
        node = nd, par = s_node, plab = s_lab, lab = k, blank = -INFINITY, label = win_s;
        if (warp == 0) node_parent[nd] = s_node, node_label[nd] = k;
      } else if (mine) {
        node = s_node, par = s_par, lab = s_lab, plab = s_plab, blank = s_blank, label = s_label;
      } else {
        node = -2, par = -1, lab = -1, plab = -1, blank = -INFINITY, label = -INFINITY;
      }
      total = mine ? win_s : -INFINITY;
      next_node += __popc(ext_mask);
      n = n_win;
    }
  }
  __syncthreads();

  // ---- top path: walk the trie back from the most probable live prefix, then emit forward with the
  // reference's groupby (consecutive duplicates collapse)
  if (tid == 0) {
    int nd = node, depth = 0;  // the beam is sorted: the most probable live prefix is lane 0
    for (int p = nd; node_parent[p] >= 0; p = node_parent[p]) ++depth;
    int* ids = out_ids + int64_t(b) * T;
    // raw labels are written back to front into ids[0..depth), then compacted in place
    int pos = depth;
    for (int p = nd; node_parent[p] >= 0; p = node_parent[p]) ids[--pos] = node_label[p] + 1;
    int m = 0;
    for (int i = 0; i < depth; ++i)
      if (i == 0 || ids[i] != ids[i - 1]) ids[m++] = ids[i];  // in place: m <= i always
    for (int i = m; i < T; ++i) ids[i] = -1;
    out_len[b] = m;
    if (out_score) out_score[b] = len > 0 ? total : 0.f;
  }
}

}  // namespace

int launch_ctc_beam(const float* logits, int B, int T, int V, const int* lengths, int beam, int* out_ids, int* out_len,
                    float* out_score, cudaStream_t s) {
  SCATT_REQUIRE(logits && out_ids && out_len, "ctc_beam_decode: null argument");
  SCATT_REQUIRE(beam >= 1 && beam <= kMaxBeam, "ctc_beam_decode: beam width %d outside 1..%d", beam, kMaxBeam);
  SCATT_REQUIRE(V >= 2 && T >= 0 && B >= 0, "ctc_beam_decode: bad shape");
  if (B == 0) return SCATT_OK;
  const int node_cap = 1 + beam * (T > 0 ? T : 1);
  const size_t smem = size_t((V + 3) & ~3) * 8 + size_t(2 * 3 * kCtcWarps) * 4 + size_t(node_cap) * 8 + size_t(2 * kCtcWarps) * beam * 8 + 16;
  SCATT_REQUIRE(smem <= 200 * 1024, "ctc_beam_decode: T=%d, V=%d, beam=%d need %zu bytes of shared memory (limit 200 KB)", T, V,
                beam, smem);
  static PerDeviceFlag configured;  // the attribute belongs to the device's context: once per device, to the limit
  if (!configured.load()) {
    SCATT_CUDA(cudaFuncSetAttribute(ctc_beam_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured.store(true);
  }
  (void)launch_kernel(ctc_beam_kernel, dim3(B), dim3(kCtcThreads), smem, s, logits, T, V, lengths, beam, node_cap, out_ids,
                      out_len, out_score);
  return after_launch("ctc_beam_kernel");
}

}  // namespace scatt
