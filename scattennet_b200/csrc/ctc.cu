// CTC prefix beam search on the per-frame logits (scope row f-4): the device-side replacement of the
// reference's utils.py:164-189 ctc_decode, which ships the logits to the host and runs TensorFlow's
// ctc_beam_search_decoder (beam 5, top path) there.  Same algorithm as TensorFlow's
// ctc_beam_search.h Step()/TopPaths() with merge_repeated = false and the default scorer, restated from
// its published description (TensorFlow itself is not available here: see oracle/ctc_oracle.py), followed
// by the reference's own post-processing: blank = class 0, ids in the original numbering, consecutive
// duplicates collapsed (itertools.groupby).
//
// One CTA per sequence.  A time step is: block-wide log-softmax of the V logits into shared memory;
// one thread advances the <= W live prefixes (a W-entry problem); all threads score the W * (V - 1) one-label
// extensions against the W-th best live prefix (an extension below it can never enter the beam), each warp
// keeping its best W in registers (lane j holds the j-th; rare insertions are serialised by ballot); one warp
// merges the 8 warp lists with the live prefixes into the next beam (order: probability, ties to the earlier
// insertion - the sequential push / pop-bottom of the original reduces to exactly this top-W selection because
// an extension can never beat the prefix it extends).  The prefix trie
// (parent, label per node; at most 1 + W * T nodes) lives in shared memory, only token ids leave the chip.
#include <cfloat>

#include "common.cuh"

namespace scatt {

namespace {

constexpr int kCtcThreads = 256;
constexpr int kMaxBeam = 16;

__device__ __forceinline__ float lse2(float a, float b) {
  if (a == -INFINITY) return b;
  if (b == -INFINITY) return a;
  return fmaxf(a, b) + log1pf(expf(-fabsf(a - b)));
}

// (score, order): higher score first, then lower order (earlier insertion)
__device__ __forceinline__ bool better(float s0, int o0, float s1, int o1) { return s0 > s1 || (s0 == s1 && o0 < o1); }

struct Beam {
  int n;
  int node[kMaxBeam];
  float old_total[kMaxBeam], old_blank[kMaxBeam];
  float total[kMaxBeam], blank[kMaxBeam], label[kMaxBeam];
  int next_node;
  int forbidden[kMaxBeam];  // extension codes (branch * (V-1) + label) whose prefix is already live
  float thr_score;          // W-th best live prefix (-inf while the beam is not full)
  // selection result of the step: order code of each winner (negative: live prefix i = code + kMaxBeam)
  int win_code[kMaxBeam];
  float win_score[kMaxBeam];
  int n_win;
};

__global__ void __launch_bounds__(kCtcThreads) ctc_beam_kernel(const float* __restrict__ logits, int T, int V,
                                                               const int* __restrict__ lengths, int W, int node_cap,
                                                               int* __restrict__ out_ids, int* __restrict__ out_len,
                                                               float* __restrict__ out_score) {
  extern __shared__ __align__(16) unsigned char ctc_smem[];
  float* logp = reinterpret_cast<float*>(ctc_smem);                          // [V]
  int* node_parent = reinterpret_cast<int*>(logp + ((V + 3) & ~3));          // [node_cap]
  int* node_label = node_parent + node_cap;                                  // [node_cap]
  float* cand_score = reinterpret_cast<float*>(node_label + node_cap);       // [warps * W] per-warp best extensions
  int* cand_code = reinterpret_cast<int*>(cand_score + (kCtcThreads / 32) * W);
  Beam& bm = *reinterpret_cast<Beam*>(cand_code + (kCtcThreads / 32) * W);
  __shared__ float red[kCtcThreads / 32];
  __shared__ float s_norm;

  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int len = lengths ? min(max(lengths[b], 0), T) : T;
  const float* x = logits + int64_t(b) * T * V;
  const int n_lab = V - 1;  // label k of the search = class k + 1 (class 0 is the blank)

  if (tid == 0) {
    bm.n = 1, bm.node[0] = 0, bm.total[0] = 0.f, bm.blank[0] = 0.f, bm.label[0] = -INFINITY, bm.next_node = 1;
    node_parent[0] = -1, node_label[0] = -1;
  }
  __syncthreads();

  for (int t = 0; t < len; ++t) {
    // ---- log-softmax of the frame
    const float* row = x + int64_t(t) * V;
    float mx = -INFINITY;
    for (int i = tid; i < V; i += kCtcThreads) mx = fmaxf(mx, row[i]);
    mx = warp_max(mx);
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    mx = red[0];
#pragma unroll
    for (int w = 1; w < kCtcThreads / 32; ++w) mx = fmaxf(mx, red[w]);
    float sum = 0.f;
    for (int i = tid; i < V; i += kCtcThreads) sum += expf(row[i] - mx);
    sum = warp_sum(sum);
    __syncthreads();
    if (lane == 0) red[warp] = sum;
    __syncthreads();
    if (tid == 0) {
      float s = 0.f;
      for (int w = 0; w < kCtcThreads / 32; ++w) s += red[w];
      s_norm = mx + logf(s);
    }
    __syncthreads();
    const float norm = s_norm;
    for (int i = tid; i < V; i += kCtcThreads) logp[i] = row[i] - norm;
    __syncthreads();

    // ---- the live prefixes at t (one thread: W <= 16 entries)
    if (tid == 0) {
      const int n = bm.n;
      // branches = live prefixes by descending probability (stable insertion sort)
      for (int i = 1; i < n; ++i) {
        const int nd = bm.node[i];
        const float tt = bm.total[i], bb = bm.blank[i], ll = bm.label[i];
        int j = i - 1;
        while (j >= 0 && bm.total[j] < tt) {
          bm.node[j + 1] = bm.node[j], bm.total[j + 1] = bm.total[j], bm.blank[j + 1] = bm.blank[j], bm.label[j + 1] = bm.label[j];
          --j;
        }
        bm.node[j + 1] = nd, bm.total[j + 1] = tt, bm.blank[j + 1] = bb, bm.label[j + 1] = ll;
      }
      for (int i = 0; i < n; ++i) bm.old_total[i] = bm.total[i], bm.old_blank[i] = bm.blank[i];
      const float lp_blank = logp[0];
      for (int i = 0; i < n; ++i) {
        const int nd = bm.node[i], par = node_parent[nd];
        if (par >= 0) {
          const int lab = node_label[nd];
          float nl = bm.label[i];
          for (int j = 0; j < n; ++j)
            if (bm.node[j] == par) {  // the parent prefix is live: paths that reach this prefix from it at t
              const float prev = (lab == node_label[par]) ? bm.old_blank[j] : bm.old_total[j];
              nl = lse2(nl, prev);
              break;
            }
          bm.label[i] = nl + logp[lab + 1];
        }
        bm.blank[i] = bm.old_total[i] + lp_blank;
        bm.total[i] = lse2(bm.blank[i], bm.label[i]);
      }
      for (int i = 0; i < kMaxBeam; ++i) bm.forbidden[i] = -1;
      float lowest = INFINITY;
      for (int i = 0; i < n; ++i) {
        lowest = fminf(lowest, bm.total[i]);
        const int par = node_parent[bm.node[i]];
        for (int j = 0; j < n && par >= 0; ++j)
          if (bm.node[j] == par) bm.forbidden[i] = j * n_lab + node_label[bm.node[i]];
      }
      bm.thr_score = n == W ? lowest : -INFINITY;
    }
    __syncthreads();

    // ---- one-label extensions.  Warp-level top-W in registers: lane j holds the warp's j-th best (ls, lc).
    const int n = bm.n;
    int forb[kMaxBeam];
#pragma unroll
    for (int j = 0; j < kMaxBeam; ++j) forb[j] = bm.forbidden[j];
    float ls = -INFINITY;
    int lc = 0x7fffffff;
    // an extension must beat the W-th best live prefix (ties lose: live prefixes were inserted first)
    float thr_s = bm.thr_score;
    int thr_c = -1;
    for (int bi = 0; bi < n; ++bi) {
      const int lab_b = node_label[bm.node[bi]];
      const float ot = bm.old_total[bi], ob = bm.old_blank[bi];
      for (int k0 = 0; k0 < n_lab; k0 += kCtcThreads) {
        const int k = k0 + tid;
        float sc = -INFINITY;
        int cd = 0x7fffffff;
        if (k < n_lab) {
          cd = bi * n_lab + k;
          sc = logp[k + 1] + (k == lab_b ? ob : ot);
          bool live_child = false;
#pragma unroll
          for (int j = 0; j < kMaxBeam; ++j) live_child |= forb[j] == cd;
          if (live_child) sc = -INFINITY;  // already advanced above
        }
        bool pending = sc != -INFINITY && better(sc, cd, thr_s, thr_c);
        unsigned mask = __ballot_sync(0xffffffffu, pending);
        while (mask) {
          const int leader = __ffs(mask) - 1;
          const float bsc = __shfl_sync(0xffffffffu, sc, leader);
          const int bcd = __shfl_sync(0xffffffffu, cd, leader);
          // sorted insert: elements better than the newcomer form a prefix of the lanes
          const int pos = __popc(__ballot_sync(0xffffffffu, better(ls, lc, bsc, bcd)));
          const float up_s = __shfl_up_sync(0xffffffffu, ls, 1);
          const int up_c = __shfl_up_sync(0xffffffffu, lc, 1);
          if (lane == pos) ls = bsc, lc = bcd;
          else if (lane > pos) ls = up_s, lc = up_c;
          if (lane >= W) ls = -INFINITY, lc = 0x7fffffff;
          const float ws = __shfl_sync(0xffffffffu, ls, W - 1);
          const int wc = __shfl_sync(0xffffffffu, lc, W - 1);
          if (ws != -INFINITY && better(ws, wc, thr_s, thr_c)) thr_s = ws, thr_c = wc;  // list full: its tail is the bar
          if (lane == leader) pending = false;
          else pending = pending && better(sc, cd, thr_s, thr_c);
          mask = __ballot_sync(0xffffffffu, pending);
        }
      }
    }
    if (lane < W) cand_score[warp * W + lane] = ls, cand_code[warp * W + lane] = lc;
    __syncthreads();

    // ---- next beam: the best W of live prefixes + the warps' lists (warp 0; a handful of items per lane)
    if (warp == 0) {
      const int total_items = (kCtcThreads / 32) * W;
      int n_win = 0;
      for (int round = 0; round < W; ++round) {
        float bs = -INFINITY;
        int bc = 0x7fffffff, bpos = -1;
        for (int i = lane; i < total_items + n; i += 32) {
          float s;
          int c;
          if (i < n) s = bm.total[i], c = i - kMaxBeam;  // live prefixes were inserted first
          else s = cand_score[i - n], c = cand_code[i - n];
          if (s != -INFINITY && better(s, c, bs, bc)) bs = s, bc = c, bpos = i;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float os = __shfl_xor_sync(0xffffffffu, bs, o);
          const int oc = __shfl_xor_sync(0xffffffffu, bc, o);
          const int op = __shfl_xor_sync(0xffffffffu, bpos, o);
          if (os != -INFINITY && better(os, oc, bs, bc)) bs = os, bc = oc, bpos = op;
        }
        if (bpos < 0) break;  // fewer than W items exist
        if (lane == 0) {
          bm.win_code[n_win] = bc, bm.win_score[n_win] = bs;
          if (bpos < n) bm.total[bpos] = -INFINITY;  // taken (its values are re-read from the saved copies below)
          else cand_score[bpos - n] = -INFINITY;
        }
        ++n_win;
        __syncwarp();
      }
      if (lane == 0) bm.n_win = n_win;
    }
    __syncthreads();

    if (tid == 0) {
      // rebuild the beam from the winners; live prefixes keep (blank, label), extensions become new trie nodes
      int nn = 0;
      int node2[kMaxBeam];
      float t2[kMaxBeam], b2[kMaxBeam], l2[kMaxBeam];
      for (int w = 0; w < bm.n_win; ++w) {
        const int c = bm.win_code[w];
        if (c < 0) {
          const int i = c + kMaxBeam;
          node2[nn] = bm.node[i], t2[nn] = bm.win_score[w], b2[nn] = bm.blank[i], l2[nn] = bm.label[i];
        } else {
          const int bi = c / n_lab, k = c - bi * n_lab;
          const int nd = bm.next_node < node_cap ? bm.next_node++ : node_cap - 1;  // cap is 1 + W * T: never exceeded
          node_parent[nd] = bm.node[bi], node_label[nd] = k;
          node2[nn] = nd, t2[nn] = bm.win_score[w], b2[nn] = -INFINITY, l2[nn] = bm.win_score[w];
        }
        ++nn;
      }
      for (int i = 0; i < nn; ++i) bm.node[i] = node2[i], bm.total[i] = t2[i], bm.blank[i] = b2[i], bm.label[i] = l2[i];
      bm.n = nn;
    }
    __syncthreads();
  }

  // ---- top path: walk the trie back from the most probable live prefix, then emit forward with the
  // reference's groupby (consecutive duplicates collapse)
  if (tid == 0) {
    int best = 0;
    for (int i = 1; i < bm.n; ++i)
      if (bm.total[i] > bm.total[best]) best = i;
    int nd = bm.node[best], depth = 0;
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
    if (out_score) out_score[b] = len > 0 ? bm.total[best] : 0.f;
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
  const size_t smem = size_t((V + 3) & ~3) * 4 + size_t(node_cap) * 8 + size_t(kCtcThreads / 32) * beam * 8 + sizeof(Beam) + 16;
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
