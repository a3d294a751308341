// CTC prefix beam search on the per-frame logits (scope row f-4): the device-side replacement of the
// reference's utils.py:164-189 ctc_decode, which ships the logits to the host and runs TensorFlow's
// ctc_beam_search_decoder (beam 5, top path) there.  Same algorithm as TensorFlow's
// ctc_beam_search.h Step()/TopPaths() with merge_repeated = false and the default scorer, restated from
// its published description (TensorFlow itself is not available here: see oracle/ctc_oracle.py), followed
// by the reference's own post-processing: blank = class 0, ids in the original numbering, consecutive
// duplicates collapsed (itertools.groupby).
//
// One CTA per sequence: seven PRODUCER warps and one SEARCH warp, no block-wide barrier inside the time loop.
//
// Producers (warps 1..7, one frame each, round-robin): everything about a frame that does not depend on the beam -
// the softmax normaliser, the blank logit, and the frame's L = min(V - 1, 2 W) best labels in the order
// (logit descending, label ascending), found by L rounds of a warp arg-max (lane-local scan of <= 36 registers,
// redux.sync on an order-preserving integer image of the value, redux.sync min on the class index among the
// equals).  A record of 2 + 2 L words per frame goes to shared memory behind a ready flag.  Why 2 W labels are
// enough: a prefix can put at most W extensions into the next beam, they are its W best labels that are not
// already live children of it, at most W - 1 labels are (every other live prefix), and one label - the prefix's own
// last label - is scored from the blank-ending mass instead of the total and can sink out of that order.
// Search (warp 0): the beam (<= W <= 16 live prefixes) lives in registers, lane i = prefix i (trie node, parent,
// last label, parent's last label, log-probabilities total / blank / label), always sorted by descending total.
// A time step is: wait for the frame's record; advance the live prefixes (lane i finds its parent among them by
// shuffles); score the <= n * L candidate extensions (branch, label) - MAXR registers per lane; decide which
// branches the sequential original actually grows (see the comment at that pass: a popped prefix can be
// deactivated before its turn); W rounds of the same redux arg-max over live prefixes + candidates of the grown
// branches (order: probability, ties to the earlier insertion code - the sequential push / pop-bottom of the original
// reduces to exactly this top-W selection over what it generates); lane w becomes winner w.  The only global loads
// on the serial path - the next frame's logits of the labels that can be live then (own last labels + the L
// candidates) - are issued at the top of the step and consumed by the next.
// The prefix trie (parent, label per node; at most 1 + W * T nodes) lives in shared memory, written and walked back by
// the search warp; only token ids leave the chip.
// oracle/ctc_oracle.py::beam_search_set_form is this algorithm on the CPU; tests/test_oracle_ctc.py checks it against
// the sequential restatement (all W paths and scores).
//
// Rounding note: candidates are ranked per frame by logit, the search ranks extensions by fl(logp + prefix mass);
// two labels whose logits differ can collide in that sum, and the tie then goes to the lower label.  The two orders
// can only disagree about the 2W-th label of a frame, which matters only if one prefix both owns all other
// live prefixes as children and wins every slot of the next beam - and the host oracle (double precision) has no
// defined answer for such collisions either.
#include <cfloat>

#include "common.cuh"

namespace scatt {

namespace {

constexpr int kCtcThreads = 256;
constexpr int kCtcWarps = kCtcThreads / 32;
constexpr int kProducers = kCtcWarps - 1;
constexpr int kMaxBeam = 16;
constexpr int kChunkRegs = 36;  // classes per lane and pass of a producer (V <= 1152 is one pass)
constexpr unsigned kFull = 0xffffffffu;
constexpr int kNone = 0x7fffffff;

// dev tool (tools/trace_ctc.py): cycle accumulators of CTA 0 - search warp [0..7], producer warp 1 [8..9]
__device__ long long* g_trace_ctc = nullptr;

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
constexpr uint32_t kKeyNegInf = 0x007fffffu;  // order_key(-inf)

// softmax partials (m = max, s = sum exp(x - m)) of two disjoint sets
__device__ __forceinline__ void lse_merge(float& m, float& s, float m2, float s2) {
  const float mm = fmaxf(m, m2);
  const float e1 = m == -INFINITY ? 0.f : expf(m - mm), e2 = m2 == -INFINITY ? 0.f : expf(m2 - mm);
  s = s * e1 + s2 * e2;
  m = mm;
}

struct CtcSmem {
  int* ready;         // [T] frame record published
  float* fr_norm;     // [T] log-sum-exp of the frame
  float* fr_blank;    // [T] raw blank logit
  float* fr_topv;     // [T][L] raw logits of the frame's best labels
  int* fr_topk;       // [T][L] their labels (class - 1)
  int* node_parent;   // [node_cap]
  int* node_label;    // [node_cap]
};
__host__ __device__ inline size_t ctc_smem_bytes(int T, int L, int node_cap) {
  return size_t(4) * (size_t(3) * T + size_t(2) * T * L + size_t(2) * node_cap) + 16;
}

// ---------------------------------------------------------------- producer: one frame's record
__device__ __forceinline__ void ctc_produce_frame(const float* __restrict__ row, int V, int L, int t, const CtcSmem& S, int lane) {
  float m_run = -INFINITY, s_run = 0.f, blank_raw = 0.f;
  float lv = -INFINITY;  // running list: lane j holds the j-th best label so far
  int li = kNone;
  for (int c0 = 0; c0 < V; c0 += 32 * kChunkRegs) {
    float xv[kChunkRegs];
#pragma unroll
    for (int r = 0; r < kChunkRegs; ++r) {
      const int i = c0 + lane + 32 * r;
      xv[r] = i < V ? __ldg(row + i) : -INFINITY;
    }
    float mc = -INFINITY;
#pragma unroll
    for (int r = 0; r < kChunkRegs; ++r) mc = fmaxf(mc, xv[r]);
    float sc = 0.f;
    if (mc != -INFINITY) {
#pragma unroll
      for (int r = 0; r < kChunkRegs; ++r) sc += expf(xv[r] - mc);  // exp(-inf) = 0 for the classes past V
    }
    lse_merge(m_run, s_run, mc, sc);
    if (c0 == 0 && lane == 0) blank_raw = xv[0], xv[0] = -INFINITY;  // class 0 is the blank: not a label
    float nlv = -INFINITY;
    int nli = kNone;
    for (int round = 0; round < L; ++round) {
      // four independent scan chains over consecutive quarters, merged in ascending order: within a lane the class
      // index grows with r, and a strict > keeps the first maximum = the lowest label
      float qv[4];
      int qr[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        qv[q] = -INFINITY, qr[q] = -1;
#pragma unroll
        for (int r = q * (kChunkRegs / 4); r < (q + 1) * (kChunkRegs / 4); ++r)
          if (xv[r] > qv[q]) qv[q] = xv[r], qr[q] = r;
      }
      float bv = qv[0];
      int br = qr[0];
#pragma unroll
      for (int q = 1; q < 4; ++q)
        if (qv[q] > bv) bv = qv[q], br = qr[q];
      int bi = br >= 0 ? c0 + lane + 32 * br : kNone;
      if (better(lv, li, bv, bi)) bv = lv, bi = li, br = -1;
      const uint32_t key = order_key(bv);
      const uint32_t kmax = __reduce_max_sync(kFull, key);
      if (kmax == kKeyNegInf) break;  // fewer than L labels with a finite logit
      const int imin = __reduce_min_sync(kFull, key == kmax ? bi : kNone);
      if (lane == round) nlv = order_key_inv(kmax), nli = imin;
      if (key == kmax && bi == imin) {  // the owner takes it out
        if (br < 0) lv = -INFINITY, li = kNone;
#pragma unroll
        for (int r = 0; r < kChunkRegs; ++r)
          if (r == br) xv[r] = -INFINITY;
      }
    }
    lv = nlv, li = nli;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float m2 = __shfl_xor_sync(kFull, m_run, o), s2 = __shfl_xor_sync(kFull, s_run, o);
    lse_merge(m_run, s_run, m2, s2);
  }
  if (lane == 0) S.fr_norm[t] = m_run + logf(s_run), S.fr_blank[t] = blank_raw;
  if (lane < L) S.fr_topv[t * L + lane] = lv, S.fr_topk[t * L + lane] = li == kNone ? 0 : li - 1;
  __syncwarp();  // orders the lanes' record stores before lane 0's release store
  if (lane == 0) asm volatile("st.release.cta.shared.b32 [%0], %1;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(S.ready + t))), "r"(1) : "memory");
}

// MAXR: candidate registers per lane of the search warp (>= W * L / 32); FW: beam widths covered (>= W)
template <int MAXR, int FW>
__global__ void __launch_bounds__(kCtcThreads, 1) ctc_beam_kernel(const float* __restrict__ logits, int T, int V,
                                                               const int* __restrict__ lengths, int W, int L, int node_cap,
                                                               int* __restrict__ out_ids, int* __restrict__ out_len,
                                                               float* __restrict__ out_score) {
  extern __shared__ __align__(16) unsigned char ctc_smem[];
  CtcSmem S;
  S.ready = reinterpret_cast<int*>(ctc_smem);
  S.fr_norm = reinterpret_cast<float*>(S.ready + T);
  S.fr_blank = S.fr_norm + T;
  S.fr_topv = S.fr_blank + T;
  S.fr_topk = reinterpret_cast<int*>(S.fr_topv + size_t(T) * L);
  S.node_parent = S.fr_topk + size_t(T) * L;
  S.node_label = S.node_parent + node_cap;

  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
  // warp index through a shuffle: the compiler then treats the role branch below as warp-uniform and issues the
  // shuffles / redux of both roles directly instead of through its divergence-safe collective sequences (10x slower)
  const int warp = __shfl_sync(kFull, tid >> 5, 0);
  const int len = lengths ? min(max(lengths[b], 0), T) : T;
  const float* x = logits + int64_t(b) * T * V;
  const int n_lab = V - 1;  // label k of the search = class k + 1 (class 0 is the blank)

  for (int t = tid; t < len; t += kCtcThreads) S.ready[t] = 0;
  if (tid == 0) S.node_parent[0] = -1, S.node_label[0] = -1;
  __syncthreads();

  long long* const trace = blockIdx.x == 0 ? g_trace_ctc : nullptr;
  if (warp > 0) {
    const long long p0 = clock64();
    int frames = 0;
    for (int t = warp - 1; t < len; t += kProducers, ++frames) ctc_produce_frame(x + int64_t(t) * V, V, L, t, S, lane);
    if (trace && warp == 1 && lane == 0) trace[8] = clock64() - p0, trace[9] = frames;
    return;
  }

  // ---------------------------------------------------------------- search warp
  int n = 1, next_node = 1;
  int node = lane == 0 ? 0 : -2, par = -1, lab = -1, plab = -1;
  float total = lane == 0 ? 0.f : -INFINITY, blank = total, label = -INFINITY;
  float own_raw = 0.f;  // raw logit of this prefix's last label in the frame about to be searched
  int cand_at[MAXR];    // candidate lane + 32 r is (branch << 8) | index into the frame's label list
#pragma unroll
  for (int r = 0; r < MAXR; ++r) {
    const int c = lane + 32 * r;
    cand_at[r] = ((c / L) << 8) | (c % L);
  }
  long long c_wait = 0, c_adv = 0, c_score = 0, c_sel = 0, c_reb = 0;
  const long long c_begin = clock64();

  for (int t = 0; t < len; ++t) {
    const long long c0 = clock64();
    const bool more = t + 1 < len;
    const float* nrow = x + int64_t(t + 1) * V;
    if (lane == 0) {  // (the acquire drains this thread's outstanding loads: the prefetches are issued after it)
      const uint32_t flag_addr = static_cast<uint32_t>(__cvta_generic_to_shared(S.ready + t));
      uint32_t ok;
      do {
        asm volatile("ld.acquire.cta.shared.b32 %0, [%1];" : "=r"(ok) : "r"(flag_addr) : "memory");
      } while (ok == 0);
    }
    __syncwarp();
    const long long c1 = clock64();
    const float norm = S.fr_norm[t];
    const float lpb = S.fr_blank[t] - norm;
    float tlp = -INFINITY;  // lane li < L: log-probability and label of the frame's li-th best label
    int tk = 0;
    if (lane < L) tlp = S.fr_topv[t * L + lane] - norm, tk = S.fr_topk[t * L + lane];
    // the next frame's logits of the labels that can be live then: consumed by the rebuild at the end of this step
    float pf_own = 0.f, pf_top = 0.f;
    if (more && lane < n && lab >= 0) pf_own = __ldg(nrow + lab + 1);
    if (more && lane < L) pf_top = __ldg(nrow + tk + 1);

    // ---- the live prefixes at t: lane i looks its parent up among them
    const float ot = total, ob = blank;
    int fj = -1;
    float prev = -INFINITY;
#pragma unroll
    for (int j = 0; j < FW; ++j) {  // lanes >= n hold node = -2: they never match
      const int nj = __shfl_sync(kFull, node, j);
      const float otj = __shfl_sync(kFull, ot, j), obj = __shfl_sync(kFull, ob, j);
      // the parent prefix is live: paths that reach this prefix from it at t
      if (par >= 0 && nj == par) fj = j, prev = (lab == plab) ? obj : otj;
    }
    if (lane < n) {
      if (par >= 0) {
        float nl = label;
        if (fj >= 0) nl = lse2(nl, prev);
        label = nl + (own_raw - norm);
      }
      blank = ot + lpb;
      total = lse2(blank, label);
    }
    // extension codes (branch * (V-1) + label) whose prefix is already live
    const int forbidden = (lane < n && fj >= 0) ? fj * n_lab + lab : -1;
    int forb[FW];
#pragma unroll
    for (int j = 0; j < FW; ++j) forb[j] = __shfl_sync(kFull, forbidden, j);

    const long long c2 = clock64();
    // ---- items of the selection: this lane's live prefix (code i - kMaxBeam: live prefixes were inserted first)
    // and candidates lane, lane + 32, ... of the n * L (branch, label) pairs
    float is[MAXR + 1];
    int ic[MAXR + 1], ia[MAXR + 1];
    is[0] = lane < n ? total : -INFINITY, ic[0] = lane < n ? lane - kMaxBeam : kNone, ia[0] = 0;
    const int n_cand = n * L;
#pragma unroll
    for (int r = 0; r < MAXR; ++r) {
      is[r + 1] = -INFINITY, ic[r + 1] = kNone, ia[r + 1] = cand_at[r];
      if (32 * r < n_cand) {  // warp-uniform
        const int bi = cand_at[r] >> 8, li = cand_at[r] & 255;  // bi < n for the valid ones; the others read some lane
        const int k = __shfl_sync(kFull, tk, li);
        const float lpk = __shfl_sync(kFull, tlp, li);
        const int lab_b = __shfl_sync(kFull, lab, bi);
        const float otb = __shfl_sync(kFull, ot, bi), obb = __shfl_sync(kFull, ob, bi);
        const float sc = lpk + (k == lab_b ? obb : otb);
        const int cd = bi * n_lab + k;
        bool ok = lane + 32 * r < n_cand && sc != -INFINITY;
        if (ok) {
#pragma unroll
          for (int j = 0; j < FW; ++j) ok = ok && forb[j] != cd;  // a live child was advanced above
        }
        if (ok) is[r + 1] = sc, ic[r + 1] = cd;
      }
    }

    // ---- which branches the original actually grows.  Its step is sequential - branches in beam order, labels in
    // ascending order, every candidate better than the current worst leaf pops that leaf - and a live prefix p that was
    // popped before its parent's label loop reaches label(p) is deactivated there INCLUDING the old probabilities its own
    // turn as a branch would grow from (ctc_beam_search.h: "Deactivate child"): the extensions of p are then never
    // generated.  Everything else of the sequence reduces to "best W of live prefixes + generated extensions", so only
    // the set of grown branches has to be reproduced, in branch order m = 0 .. n-1:
    //   grown(m)  = p_m not deactivated;
    //   for every live child p_j (j > m) of a grown p_m: p_j is deactivated iff at least W items generated before the
    //   loop of m reaches label(p_j) beat its total: live prefixes, extensions of grown branches < m, and extensions
    //   (m, k) with k < label(p_j).
    // Counts over a whole branch come from the 2 W labels in registers and saturate correctly: if every listed label of a
    // branch beats a total there are >= W of them.  Only the count over k < label(p_j) depends on label ORDER and scans
    // the frame's logits (every tenth step or so).
    // (The original also skips a branch whose old total does not beat the current worst leaf - none of its extensions, nor
    // those of its later children, could enter: that skip needs no reproduction.  And a step without a live child ordered
    // after its live parent has nothing to deactivate.)
    unsigned grown = 0, dead = 0;
    const unsigned pairs = __ballot_sync(kFull, lane < n && fj >= 0 && fj < lane);
    for (int m = 0; m < n; ++m) {
      const float o_m = __shfl_sync(kFull, ot, m);
      if (((dead >> m) & 1u) || o_m == -INFINITY) continue;
      grown |= 1u << m;
      if (!pairs) continue;
      unsigned kids = __ballot_sync(kFull, lane < n && fj == m && lane > m);
      while (kids) {
        const int j = __ffs(kids) - 1;
        kids &= kids - 1;
        const float a_j = __shfl_sync(kFull, total, j);
        const int lab_j = __shfl_sync(kFull, lab, j);
        int c1 = lane < n && lane != j && (total > a_j || (total == a_j && lane > j));
        int cm = 0;
#pragma unroll
        for (int r = 0; r < MAXR; ++r) {
          const int bi = cand_at[r] >> 8;
          c1 += bi < m && ((grown >> bi) & 1u) && is[r + 1] > a_j;
          cm += bi == m && is[r + 1] > a_j;
        }
        const int base = __reduce_add_sync(kFull, c1);
        bool out = base >= W;
        if (!out && base + __reduce_add_sync(kFull, cm) >= W) {
          // ambiguous: count the extensions (m, k), k < label(p_j), that beat p_j - over all labels, in label order
          const float ot_m = o_m, ob_m = __shfl_sync(kFull, ob, m);
          const int lab_m = __shfl_sync(kFull, lab, m);
          const float* row = x + int64_t(t) * V;
          int cp = 0;
#pragma unroll 4
          for (int k = lane; k < lab_j; k += 32) {
            const float sc = (__ldg(row + k + 1) - norm) + (k == lab_m ? ob_m : ot_m);
            bool hit = sc > a_j;
            const int cd = m * n_lab + k;
#pragma unroll
            for (int q = 0; q < FW; ++q) hit = hit && forb[q] != cd;
            cp += hit;
          }
          out = base + __reduce_add_sync(kFull, cp) >= W;
        }
        if (out) dead |= 1u << j;
      }
    }
#pragma unroll
    for (int r = 0; r < MAXR; ++r) {
      const int bi = cand_at[r] >> 8;  // slots past n * L carry branch numbers >= n
      if (bi >= kMaxBeam || !((grown >> bi) & 1u)) is[r + 1] = -INFINITY, ic[r + 1] = kNone;
    }

    const long long c3 = clock64();
    // ---- next beam: the best W items, in order
    float win_s = -INFINITY;
    int win_c = kNone, win_a = 0, n_win = 0;
    for (int round = 0; round < W; ++round) {
      float bs = is[0];
      int bc = ic[0], ba = ia[0];
#pragma unroll
      for (int r = 1; r <= MAXR; ++r)
        if (better(is[r], ic[r], bs, bc)) bs = is[r], bc = ic[r], ba = ia[r];
      const uint32_t key = order_key(bs);
      const uint32_t kmax = __reduce_max_sync(kFull, key);
      if (kmax == kKeyNegInf) break;  // fewer than W items exist
      const int cmin = __reduce_min_sync(kFull, key == kmax ? bc : kNone);
      // the owner (codes of valid items are unique) takes it out and tells which (branch, label slot) it was
      const bool owner = key == kmax && bc == cmin;
      const int a_w = __shfl_sync(kFull, ba, __ffs(__ballot_sync(kFull, owner)) - 1);
#pragma unroll
      for (int r = 0; r <= MAXR; ++r)
        if (owner && ic[r] == cmin) is[r] = -INFINITY;
      if (lane == n_win) win_s = order_key_inv(kmax), win_c = cmin, win_a = a_w;
      ++n_win;
    }

    const long long c4 = clock64();
    // ---- rebuild: lane w becomes winner w; live prefixes keep (blank, label), extensions become new trie nodes
    {
      const bool mine = lane < n_win;
      const bool is_ext = mine && win_c >= 0;
      const int bi = is_ext ? win_a >> 8 : 0, li_w = is_ext ? win_a & 255 : 0;
      const int k = __shfl_sync(kFull, tk, li_w);
      const int src = mine ? (win_c < 0 ? win_c + kMaxBeam : bi) : 0;
      const int s_node = __shfl_sync(kFull, node, src), s_par = __shfl_sync(kFull, par, src);
      const int s_lab = __shfl_sync(kFull, lab, src), s_plab = __shfl_sync(kFull, plab, src);
      const float s_blank = __shfl_sync(kFull, blank, src), s_label = __shfl_sync(kFull, label, src);
      const float s_own = __shfl_sync(kFull, pf_own, src), s_top = __shfl_sync(kFull, pf_top, li_w);
      const unsigned ext_mask = __ballot_sync(kFull, is_ext);
      if (is_ext) {
        const int nd = min(next_node + __popc(ext_mask & ((1u << lane) - 1u)), node_cap - 1);  // cap is 1 + W * T: never exceeded
        node = nd, par = s_node, plab = s_lab, lab = k, blank = -INFINITY, label = win_s, own_raw = s_top;
        S.node_parent[nd] = s_node, S.node_label[nd] = k;
      } else if (mine) {
        node = s_node, par = s_par, lab = s_lab, plab = s_plab, blank = s_blank, label = s_label, own_raw = s_own;
      } else {
        node = -2, par = -1, lab = -1, plab = -1, blank = -INFINITY, label = -INFINITY, own_raw = 0.f;
      }
      total = mine ? win_s : -INFINITY;
      next_node += __popc(ext_mask);
      n = n_win;
    }
    const long long c5 = clock64();
    c_wait += c1 - c0, c_adv += c2 - c1, c_score += c3 - c2, c_sel += c4 - c3, c_reb += c5 - c4;
  }
  __syncwarp();
  if (trace && lane == 0)
    trace[0] = clock64() - c_begin, trace[1] = c_wait, trace[2] = c_adv, trace[3] = c_score, trace[4] = c_sel, trace[5] = c_reb, trace[6] = len;

  // ---- top path: walk the trie back from the most probable live prefix, then emit forward with the
  // reference's groupby (consecutive duplicates collapse)
  if (tid == 0) {
    int nd = node, depth = 0;  // the beam is sorted: the most probable live prefix is lane 0
    for (int p = nd; S.node_parent[p] >= 0; p = S.node_parent[p]) ++depth;
    int* ids = out_ids + int64_t(b) * T;
    // raw labels are written back to front into ids[0..depth), then compacted in place
    int pos = depth;
    for (int p = nd; S.node_parent[p] >= 0; p = S.node_parent[p]) ids[--pos] = S.node_label[p] + 1;
    int m = 0;
    for (int i = 0; i < depth; ++i)
      if (i == 0 || ids[i] != ids[i - 1]) ids[m++] = ids[i];  // in place: m <= i always
    for (int i = m; i < T; ++i) ids[i] = -1;
    out_len[b] = m;
    if (out_score) out_score[b] = len > 0 ? total : 0.f;
  }
}

}  // namespace

int debug_set_trace_ctc(void* dev_buf) {
  long long* p = reinterpret_cast<long long*>(dev_buf);
  SCATT_CUDA(cudaMemcpyToSymbol(g_trace_ctc, &p, sizeof(p)));
  return SCATT_OK;
}

int launch_ctc_beam(const float* logits, int B, int T, int V, const int* lengths, int beam, int* out_ids, int* out_len,
                    float* out_score, cudaStream_t s) {
  SCATT_REQUIRE(logits && out_ids && out_len, "ctc_beam_decode: null argument");
  SCATT_REQUIRE(beam >= 1 && beam <= kMaxBeam, "ctc_beam_decode: beam width %d outside 1..%d", beam, kMaxBeam);
  SCATT_REQUIRE(V >= 2 && T >= 0 && B >= 0, "ctc_beam_decode: bad shape");
  if (B == 0) return SCATT_OK;
  const int node_cap = 1 + beam * (T > 0 ? T : 1);
  const int L = std::min(V - 1, 2 * beam);  // labels of a frame that can enter the beam (see the header)
  const size_t smem = ctc_smem_bytes(T, L, node_cap);
  SCATT_REQUIRE(smem <= 200 * 1024, "ctc_beam_decode: T=%d, V=%d, beam=%d need %zu bytes of shared memory (limit 200 KB)", T, V,
                beam, smem);
  static PerDeviceFlag configured;  // the attribute belongs to the device's context: once per device, to the limit
  if (!configured.load()) {
    SCATT_CUDA(cudaFuncSetAttribute(ctc_beam_kernel<2, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    SCATT_CUDA(cudaFuncSetAttribute(ctc_beam_kernel<4, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    SCATT_CUDA(cudaFuncSetAttribute(ctc_beam_kernel<16, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured.store(true);
  }
  const int n_cand = beam * L;  // candidate extensions per step: MAXR registers per lane of the search warp
  if (n_cand <= 64 && beam <= 8)
    (void)launch_kernel(ctc_beam_kernel<2, 8>, dim3(B), dim3(kCtcThreads), smem, s, logits, T, V, lengths, beam, L, node_cap, out_ids,
                        out_len, out_score);
  else if (n_cand <= 128 && beam <= 8)
    (void)launch_kernel(ctc_beam_kernel<4, 8>, dim3(B), dim3(kCtcThreads), smem, s, logits, T, V, lengths, beam, L, node_cap, out_ids,
                        out_len, out_score);
  else
    (void)launch_kernel(ctc_beam_kernel<16, 16>, dim3(B), dim3(kCtcThreads), smem, s, logits, T, V, lengths, beam, L, node_cap, out_ids,
                        out_len, out_score);
  return after_launch("ctc_beam_kernel");
}

}  // namespace scatt
