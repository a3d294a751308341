/*
 * scatt.h - C ABI of libscatt.so, the sm_100a kernels behind scattennet_b200.
 *
 * The reference (tinh2044/SCAttenNet) has no FFI / plugin interface: its
 * boundary is the Python nn.Module surface.  This header
 * is the "thin C-ABI extension" the host-side PyTorch modules call through
 * ctypes; every entry point names the reference code it replaces
 * (paths relative to the reference repo root).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - the caller owns every buffer (inputs, outputs, scratch); the library
 *     never allocates, never synchronises, never touches the default stream:
 *     work is enqueued on `stream` (a cudaStream_t passed as void*), so calls
 *     are capturable into CUDA graphs;
 *   - return value: 0 on success, a negative SCATT_ERR_* code otherwise;
 *     scatt_last_error() returns a thread-local message for the last failure;
 *   - re-entrant across host threads and streams; no global mutable state
 *     besides per-kernel attribute initialisation (idempotent);
 *   - sm_100a only.  There is no CPU fallback: on a machine without a B200
 *     the compute entry points return SCATT_ERR_CUDA.
 *
 * "Split planes": activations and weights that feed a tensor-core GEMM are
 * stored as 16-bit hi/lo planes, `planes[0] = rn16(x)`, `planes[1] =
 * rn16(x - planes[0])`, laid out [2][rows][cols] row-major (fp16 or bf16, see
 * scatt_plane_fmt).  A GEMM with `terms` = 1 uses hi*hi only, 2 adds lo*hi
 * (activation low part), 3 adds hi*lo as well (fp32-grade products).
 */
#ifndef SCATT_H_
#define SCATT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SCATT_ABI_VERSION 3
#define SCATT_MAX_GROUP 4 /* problems per grouped launch (the 3 anatomical streams + spare) */
#define SCATT_MAX_PEERS 8  /* GPUs of one NVSwitch box taking part in scatt_peer_allgather */
#define SCATT_MAX_FINITE 16 /* tensors per scatt_finite_check call */

enum scatt_error {
  SCATT_OK = 0,
  SCATT_ERR_INVALID = -1,     /* bad argument (shape, alignment, enum) */
  SCATT_ERR_UNSUPPORTED = -2, /* valid request this build does not implement */
  SCATT_ERR_CUDA = -3,        /* CUDA runtime / driver error (message has the string) */
  SCATT_ERR_KERNEL = -4       /* a kernel reported a device-side failure (pipeline time-out) */
};

enum scatt_engine {
  SCATT_ENGINE_SIMT = 0,   /* fp32 FMA on CUDA cores: exact-order reference engine, fp32 tier */
  SCATT_ENGINE_TCGEN05 = 1 /* TMA -> smem -> tcgen05.mma (TMEM accumulators), split-plane operands */
};

enum scatt_plane_fmt { SCATT_PLANE_F16 = 0, SCATT_PLANE_BF16 = 1 };

enum scatt_act { SCATT_ACT_NONE = 0, SCATT_ACT_GELU = 1 /* exact erf */, SCATT_ACT_RELU = 2 };

enum scatt_residual_mode { SCATT_RES_NONE = 0, SCATT_RES_BEFORE_LN = 1, SCATT_RES_AFTER_LN = 2 };

enum scatt_attention_kind { SCATT_ATTN_SELF = 0, SCATT_ATTN_CAUSAL = 1, SCATT_ATTN_CROSS = 2 };

/* ------------------------------------------------------------------ misc */

int scatt_abi_version(void);
const char* scatt_version(void);
const char* scatt_last_error(void);
/* Name of the kernel the calling host thread launched last, with its template arguments as a demangler prints them
 * ("linear_tc_kernel<256, 1, 0>"): lets bench.py attach algorithmic flops / bytes to the symbols of a profiler trace. */
const char* scatt_last_kernel(void);
/* Number of kernels this library has launched from the calling process (all
 * threads); bench.py's `gpu_launches` is the difference across the timed region. */
uint64_t scatt_launch_count(void);
/* 0 if the current device is sm_100 and the kernels are loadable. */
int scatt_device_check(void);
/* Developer aid: when `dev_buf` (device memory, >= 16 x int64) is non-NULL, CTA 0 of
 * every tcgen05 linear launch stores clock64() stamps of its pipeline phases there
 * (tools/trace_linear.py); NULL switches tracing off. */
int scatt_debug_set_trace(void* dev_buf);

/* ------------------------------------------------------------------ split planes */

/* planes[0..1][rows][cols] <- split(scale * x[rows][ldx]); used to pack weights
 * once (scale = 0.5 folds the `key_value_states / 2` of CrossAttention,
 * model/attention.py:103, into W_v exactly) and to import fp32 activations. */
int scatt_split_planes(const float* x, int64_t rows, int64_t cols, int64_t ldx, float scale,
                       void* planes, int plane_fmt, void* stream);

/* Hint: pull `n` static device buffers (packed weight planes; 16-byte aligned, < 2 GiB each) into L2 with
 * cp.async.bulk.prefetch.L2, one launch per 1024 buffers.  The reference has no counterpart (ATen leaves cache
 * residency to the hardware); MSCAEncoder issues it on a parallel graph branch at the top of a small-batch step so
 * that the ~60 dependent launches that follow do not each pay DRAM latency on their first weight tile. */
int scatt_l2_prefetch(const void* const* ptrs, const int64_t* nbytes, int n, void* stream);

/* ------------------------------------------------------------------ K1: front end */

/* One anatomical stream of the fused front end.  Replaces, in one pass over
 * keypoints[B,T,K,2]: the region gather `keypoints[:, :, idx, :]`
 * (model/__init__.py:133-142), the x / y split (model/keypoint_module.py:23-24),
 * CoordinateMapping (model/layers.py:118-123), LearningPositionEmbedding
 * (model/layers.py:20-30, table row t+2) and first_self_norm /
 * first_causal_norm (model/keypoint_module.py:155-162).
 * Branch 0 is the "self" branch, branch 1 the "causal" branch; `coord[br]`
 * says which coordinate (0 = x, 1 = y) feeds it (cfg self_attn_x). */
typedef struct scatt_frontend_stream {
  const int32_t* joint_idx; /* [n_joints] indices into K */
  int32_t n_joints;         /* <= 32 */
  int32_t coord[2];
  const float* map_wt[2];   /* [n_joints, D]: TRANSPOSED mapping weight (nn.Linear.weight^T) of the coordinate feeding branch br */
  const float* map_b[2];    /* [D] */
  const float* pos[2];      /* [max_pos + 2, D] position tables (self, causal) */
  const float* ln_g[2];     /* [D] */
  const float* ln_b[2];     /* [D] */
  float* out[2];            /* [B*T, D] fp32, may be NULL */
  void* out_planes[2];      /* [2][B*T][D] split planes, may be NULL */
  float* gathered;          /* optional [B,T,n_joints,2] exact copy of the gathered region, may be NULL */
} scatt_frontend_stream;

int scatt_frontend(const float* keypoints, int B, int T, int K, int D, const scatt_frontend_stream* streams_host,
                   int n_streams, int max_pos, int plane_fmt, void* stream);

/* x[B,T,D] + table[t+2] -> LayerNorm -> out (+ planes).  The same two steps
 * for callers that enter below KeypointModule (SeparativeCoordinateAttention /
 * Encoder called directly: model/keypoint_module.py:155-162, model/encoder.py:80-82). */
int scatt_posembed_layernorm(const float* x, const float* table, const float* ln_g, const float* ln_b, float* out,
                             void* out_planes, int B, int T, int D, int max_pos, int plane_fmt, void* stream);

/* ------------------------------------------------------------------ K2: linear + epilogue */

typedef struct scatt_epilogue {
  int32_t act_pre;       /* scatt_act applied to (acc + bias) * colscale */
  int32_t residual_mode; /* scatt_residual_mode */
  int32_t layer_norm;    /* 1: LayerNorm over the N outputs of a row (eps, affine) */
  int32_t act_post;      /* scatt_act applied last */
  float clamp;           /* > 0: clamp to [-clamp, clamp] (RecognitionHead, model/__init__.py:56-60) */
  int32_t scale_cols;    /* columns [0, scale_cols) are multiplied by `scale` after the bias ... */
  float scale;           /* ... (q = (x Wq + bq) * head_dim^-0.5, model/attention.py:49) */
  float ln_eps;          /* 1e-5 */
} scatt_epilogue;

/* One problem of a grouped launch: y = epilogue(x W^T + bias).
 * Replaces nn.Linear call sites of the path with their trailing elementwise /
 * LayerNorm ops fused: q/k/v/out projections (model/attention.py:49-51,74),
 * FeedForward (model/layers.py:103-108), residual + LayerNorm
 * (model/keypoint_module.py:62-72,98-107), ResidualBlock linears + LayerNorm +
 * ReLU (model/residual.py:25-38), CoordinatesFusion / InvertedResidual linears
 * (model/fusion.py:37-53,67-78), RecognitionHead classifiers + clamp
 * (model/__init__.py:49-60). */
typedef struct scatt_linear_problem {
  const float* x;        /* SIMT engine: [M, K] fp32, row stride ldx */
  const void* x_planes;  /* TCGEN05 engine: [2][M][K] split planes (contiguous) */
  const float* w;        /* SIMT engine: [N, K] fp32 contiguous (nn.Linear.weight) */
  const void* w_planes;  /* TCGEN05 engine: [2][N][K] split planes */
  const float* bias;     /* [N] or NULL */
  const float* residual; /* [M, N] fp32 (row stride ldres) or NULL */
  const float* ln_g;     /* [N] */
  const float* ln_b;     /* [N] */
  float* y;              /* [M, N] fp32, row stride ldy; may be NULL if y_planes is given and no LayerNorm scratch is needed */
  void* y_planes;        /* [2][M][N] split planes or NULL */
  const void* residual_planes; /* the residual as [2][M][N] split planes (hi + lo is added) when `residual` is NULL:
                                * lets a residual stream live in planes only.  LayerNorm GEMMs of the tcgen05 engine
                                * only (scatt_linear_ln_fused): the cluster kernels stage the planes by TMA; the
                                * one-CTA-per-row-tile kernel (N = 256, more than 74 row tiles) takes them with
                                * RES_BEFORE_LN and no pre-activation / column scaling */
} scatt_linear_problem;

int scatt_linear(const scatt_linear_problem* problems_host, int group, int64_t M, int N, int K, int64_t ldx,
                 int64_t ldres, int64_t ldy, const scatt_epilogue* epilogue_host, int engine, int plane_fmt, int terms,
                 void* stream);

/* scatt_linear with a caller-owned scratch buffer.  Launches over few row tiles with a long K loop (the fusion block of
 * model/fusion.py at small batches: M = B*T' rows, K = 1024 / 3072) are split along K when the scratch is given: S
 * slices run as S problem slots of one GEMM launch that writes fp32 partial sums into `workspace`, and a row-wise
 * kernel adds them in slot order and applies the epilogue.  scatt_linear_workspace_bytes returns the bytes such a
 * launch wants (0: the launch is not split; any buffer, or none, may be passed). */
size_t scatt_linear_workspace_bytes(int group, int64_t M, int N, int K, int engine);
int scatt_linear_ws(const scatt_linear_problem* problems_host, int group, int64_t M, int N, int K, int64_t ldx,
                    int64_t ldres, int64_t ldy, const scatt_epilogue* epilogue_host, int engine, int plane_fmt,
                    int terms, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------ K2b: fused row-local layer tail
 *
 * y = LayerNorm2(h + fc2(GELU(fc1(h)))),  h = LayerNorm1(x + ctx Wo^T + bo)
 * i.e. everything of a self / merge layer behind the attention core in ONE kernel (tcgen05 engine): out_proj
 * (model/attention.py:74,126), residual + attn_layer_norm (model/keypoint_module.py:62-66,98-102), FeedForward
 * (model/layers.py:103-108), residual + last_layer_norm (model/keypoint_module.py:68-72,104-107); the same tail
 * closes an EncoderLayer (model/encoder.py:38-55).  h and the F-wide hidden activation stay on the SM (shared /
 * tensor memory); per 128-row tile only ctx and x are read and y is written.  D must be 256, F a multiple of 128
 * up to 1024 (scatt_attn_block_supported); other shapes run as three scatt_linear calls. */
typedef struct scatt_block_problem {
  const void* ctx_planes;      /* [2][M][D] attention output, heads concatenated (scatt_attention_planes out_planes) */
  const void* residual_planes; /* [2][M][D] the layer input x (hi + lo is added), 16-byte aligned */
  const void* wo_planes;       /* [2][D][D] out_proj.weight */
  const float* bo;             /* [D] */
  const float* ln1_g;          /* [D] attn_layer_norm */
  const float* ln1_b;
  const void* w1_planes;       /* [2][F][D] fc1.weight */
  const float* b1;             /* [F] */
  const void* w2_planes;       /* [2][D][F] fc2.weight */
  const float* b2;             /* [D] */
  const float* ln2_g;          /* [D] last_layer_norm */
  const float* ln2_b;
  float* y;                    /* [M][D] fp32 contiguous or NULL */
  void* y_planes;              /* [2][M][D] split planes or NULL */
} scatt_block_problem;

int scatt_attn_block(const scatt_block_problem* problems_host, int group, int64_t M, int D, int F, float ln_eps,
                     int plane_fmt, int terms, void* stream);
int scatt_attn_block_supported(int64_t M, int D, int F);
/* The causal layer in front of a merge layer as ONE launch: out_proj + residual + attn_layer_norm of
 * CoordinateAttention(causal) (model/keypoint_module.py:62-66, no FeedForward on that branch) and the q_proj of the
 * CoordinatesMerge layer that consumes it (CrossAttention.forward, model/attention.py:99-101: q = q_proj(h) * scaling).
 *   h = LayerNorm(x + ctx Wo^T + bo) -> h_planes,   q = (h Wq^T + bq) * q_scale -> q_planes
 * Same kernel as scatt_attn_block (h stays in shared memory as the A operand of the second product). D = N = 256. */
typedef struct scatt_outq_problem {
  const void* ctx_planes;      /* [2][M][D] attention output of the causal layer */
  const void* residual_planes; /* [2][M][D] the causal layer's input x */
  const void* wo_planes;       /* [2][D][D] out_proj.weight */
  const float* bo;             /* [D] */
  const float* ln_g;           /* [D] attn_layer_norm */
  const float* ln_b;
  const void* wq_planes;       /* [2][N][D] the merge layer's q_proj.weight */
  const float* bq;             /* [N] */
  void* h_planes;              /* out [2][M][D] */
  void* q_planes;              /* out [2][M][N] */
} scatt_outq_problem;

int scatt_attn_out_q(const scatt_outq_problem* problems_host, int group, int64_t M, int D, int N, float ln_eps, float q_scale,
                     int plane_fmt, int terms, void* stream);
int scatt_attn_out_q_supported(int64_t M, int D, int N);
/* Developer aid: 1 | 2 forces one CTA / a 2-CTA cluster per 128-row tile in scatt_attn_block, 0 restores the automatic
 * choice (clusters while there are at most 74 row tiles). */
int scatt_debug_set_block_cluster(int cluster);

/* 1 when scatt_linear with a LayerNorm epilogue normalises inside the GEMM kernel for this shape (tcgen05
 * engine: N = 256 always; N = 512 / 1024 by 4- / 8-CTA clusters while ceil(M / 128) * group * N / 128 <= 148),
 * 0 when it runs the GEMM and then the row-wise tail in place on y - y is then required as scratch. */
int scatt_linear_ln_fused(int64_t M, int N, int group, int engine);

/* Row-wise tail on an fp32 [M, N] matrix: optional LayerNorm, residual after,
 * activation, clamp, split-plane export.  In-place (y == z) is allowed. */
int scatt_rowwise(const float* z, int64_t M, int N, int64_t ldz, const float* residual, int64_t ldres,
                  const float* ln_g, const float* ln_b, const scatt_epilogue* epilogue_host, float* y, int64_t ldy,
                  void* y_planes, int plane_fmt, void* stream);

/* ------------------------------------------------------------------ K3: stream attention */

/* softmax(q k^T + mask) v per head for one stream; q is already scaled.
 * Replaces the score / mask / softmax / PV block of SelfAttention,
 * CrossAttention and SelfCausalAttention (model/attention.py:53-73,105-125,
 * 155-179) and the mask builders (model/utils.py:3-28) without materialising
 * scores or masks.  Mask semantics: `key_mask` ([B,Tk] uint8, 1 = valid) gives
 * a padded key the logit finfo(float32).min exactly like the additive mask of
 * the reference (so a row whose permitted keys are all padded is uniform over
 * them); `additive` ([B,1,Tq,Tk] fp32) is the generic low-level-interface
 * mask and is added verbatim; causal rows see keys j <= i only. */
typedef struct scatt_attention_problem {
  const float* q; /* [B*Tq, H*hd] fp32, row stride ldq */
  const float* k; /* [B*Tk, H*hd] row stride ldk */
  const float* v; /* [B*Tk, H*hd] row stride ldv */
  const uint8_t* key_mask;
  const float* additive;
  float* out;       /* [B*Tq, H*hd] fp32 contiguous or NULL */
  void* out_planes; /* [2][B*Tq][H*hd] or NULL */
} scatt_attention_problem;

/* `engine` = SCATT_ENGINE_TCGEN05 runs both contractions on the tensor cores
 * (split planes built in shared memory, `terms` product terms) when Tk <= 256
 * and no dense additive mask is given; otherwise, and for SCATT_ENGINE_SIMT,
 * the fp32 CUDA-core kernel runs. */
int scatt_attention(const scatt_attention_problem* problems_host, int group, int B, int Tq, int Tk, int H, int hd,
                    int64_t ldq, int64_t ldk, int64_t ldv, int kind, int engine, int plane_fmt, int terms, void* stream);

/* Same attention, operands taken as the split planes a projection GEMM wrote
 * (`scatt_linear` with y_planes): planes[2][rows][ld], head h at columns
 * [col + hd*h, col + hd*(h+1)).  Tiles are fetched by TMA straight into tensor-core
 * layout - nothing is converted or transposed - so this is the fast path used by
 * SeparativeCoordinateAttention / Encoder.  Requires hd = 16 and Tk <= 1568 (use
 * scatt_attention otherwise); q must already carry the head_dim^-0.5 scaling. */
typedef struct scatt_attn_operand {
  const void* planes;
  int64_t rows; /* rows of the plane matrix (>= B*T) */
  int64_t ld;   /* row stride in elements, multiple of 8 */
  int32_t col;  /* first column of head 0, multiple of 8 */
  int32_t reserved;
} scatt_attn_operand;

typedef struct scatt_attention_planes_problem {
  scatt_attn_operand q, k, v;
  const uint8_t* key_mask; /* [B,Tk] 1 = valid, or NULL */
  float* out;              /* [B*Tq, H*hd] fp32 or NULL */
  void* out_planes;        /* [2][B*Tq][H*hd] or NULL */
} scatt_attention_planes_problem;

int scatt_attention_planes(const scatt_attention_planes_problem* problems_host, int group, int B, int Tq, int Tk, int H,
                           int hd, int kind, int plane_fmt, int terms, void* stream);
/* Developer aid: the schedule of scatt_attention_planes - 0 by item count (one (batch, head, query tile) item per CTA
 * below 4096 items, the persistent two-group kernel from there on), 1 forces the persistent kernel (where its
 * shared-memory map fits: Tk <= 672), 2 one item per CTA. */
int scatt_debug_set_attn_persist(int mode);

/* ------------------------------------------------------------------ K5: fusion attention */

/* out[b] = softmax(q[b] k[b]^T) v[b], single head of width D, no mask, no
 * scaling (model/fusion.py:46-49: q = right_out, k = left_out, v = body_out). */
int scatt_fusion_attention(const float* q, const float* k, const float* v, int B, int T, int D, float* out,
                           void* out_planes, int plane_fmt, void* stream);

/* Same contraction on the tensor cores (tcgen05): q / k / v are the split planes [2][B*T][D] written by the
 * left_se / right_se / body_se GEMMs (`scatt_linear` with y_planes, GELU applied), `terms` product terms as in
 * scatt_linear.  Requires T <= 256 and D a multiple of 256 (`scatt_fusion_attention_planes_supported`); longer
 * sequences use scatt_fusion_attention. */
int scatt_fusion_attention_planes_supported(int T, int D);
int scatt_fusion_attention_planes(const void* q_planes, const void* k_planes, const void* v_planes, int B, int T, int D,
                                  float* out, void* out_planes, int plane_fmt, int terms, void* stream);

/* ------------------------------------------------------------------ K4: temporal pooling */

/* MaxPool1d(2,2) over time of x[B,T,C] -> y[B,floor(T/2),C] (model/residual.py:40-43),
 * with optional split-plane export of the pooled rows. */
int scatt_pool_pairs(const float* x, int B, int T, int C, float* y, void* y_planes, int plane_fmt, void* stream);
/* Same for `group` same-shaped tensors (the anatomical streams) in one launch; the three arrays are host
 * arrays of device pointers (`planes_host` or its entries may be NULL). */
int scatt_pool_pairs_group(const float* const* xs_host, float* const* ys_host, void* const* planes_host, int group, int B,
                           int T, int C, int plane_fmt, void* stream);

/* ------------------------------------------------------------------ consumers of the path */

/* Recurrent part of one bidirectional LSTM layer (nn.LSTM of AlignmentModule, model/alignment_module.py:25-33,66-72;
 * hidden size per direction H = 512, zero initial state, every sequence runs all T steps exactly like the
 * reference, which does not pack padded sequences).
 *   gates_x  [B*T, ldg] fp32, row b*T + t, columns dir*4H + gate*H + unit in torch's gate order (i, f, g, o):
 *            the input projection x_t W_ih^T + b_ih + b_hh of both directions (one scatt_linear call, N = 8H);
 *   w_hh     [2][4H][H] fp32: weight_hh_l{k} then weight_hh_l{k}_reverse;
 *   y        [B*T, 2H] fp32 (row b*T + t; forward states in columns [0,H), reverse in [H,2H)) and / or
 *   y_planes [2][B*T][2H] split planes of the same matrix (operand of the next GEMM);
 *   workspace: scatt_lstm_workspace_bytes(B, H) bytes of device memory owned by the caller (the h exchange
 *            buffer; zeroed on `stream` by this call).
 * One persistent co-operative launch of 128 CTAs runs all T steps with W_hh resident in shared memory. */
size_t scatt_lstm_workspace_bytes(int64_t B, int H);
int scatt_lstm_bidir(const float* gates_x, int64_t ldg, const float* w_hh, float* y, void* y_planes, void* workspace,
                     int64_t B, int T, int H, int plane_fmt, void* stream);

/* out = clamp(log_softmax(logits, dim=-1), clamp_min, clamp_max) for logits [B*T, V] (row b*T + t, row stride ld),
 * written batch-major ([B,T,V]) or, with time_major != 0, as [T,B,V] - the permute + log_softmax + clamp(-100, 0)
 * that feeds nn.CTCLoss in MSCA_Net.compute_loss (model/__init__.py:243-250). */
int scatt_log_softmax(const float* logits, int64_t ld, int V, int B, int T, int time_major, float clamp_min,
                      float clamp_max, float* out, void* stream);

/* *flags_dev = OR over i of (1 << i) for every tensor i (host arrays of `count` device pointers / element counts)
 * holding a NaN or an infinity: one launch and one 4-byte read-back instead of the 16 host-synchronising
 * isnan / isinf checks of MSCA_Net.forward (model/__init__.py:130-167). */
int scatt_finite_check(const float* const* tensors_host, const int64_t* sizes_host, int count, int* flags_dev,
                       void* stream);

/* CTC prefix beam search, top path, on logits [B, T, V] fp32 (contiguous; class 0 = blank, as nn.CTCLoss(blank=0)
 * and the reference's logits have it): the device-side form of utils.py:164-189 ctc_decode - rotation of the blank
 * to TensorFlow's position, tf.nn.ctc_beam_search_decoder(beam_width = beam, top_paths = 1), + 1, and the
 * groupby that collapses consecutive duplicates - so that only gloss ids leave the GPU instead of the logits.
 *   lengths  [B] int32 valid frames per sequence (NULL = T for all);
 *   out_ids  [B, T] int32: the decoded gloss ids (original class numbering, 1..V-1), padded with -1;
 *   out_len  [B] int32 number of ids; out_score [B] log-probability of the winning prefix (may be NULL).
 * beam <= 16; shared memory grows with V, T and beam (V = 1120, T = 512, beam = 5: ~75 KB). */
int scatt_ctc_beam_decode(const float* logits, int B, int T, int V, const int32_t* lengths, int beam, int32_t* out_ids,
                          int32_t* out_len, float* out_score, void* stream);

/* ------------------------------------------------------------------ K6: logits all-gather over NVLink peer memory
 *
 * The path's one exchange step (the reference has no parallelism, its NCCL group is never used:
 * utils.py:237-265).  Rank `rank` pushes `bytes` (multiple of 16) from `src` into slot `rank` of every peer's
 * gather buffer - peer_bufs_host[p] is the base of rank p's buffer [world][bytes], mapped into this process
 * (torch.distributed symmetric memory / CUDA VMM) - and the same launch runs the barrier: it publishes `seq` in
 * entry `rank` of every peer's flag pad (peer_flags_host[p]: uint64[world], zero-initialised) and returns, in
 * stream order, once all `world` entries of its own pad have reached `seq`.  `seq` must grow by one per call and
 * the caller must alternate between two buffers (seq parity), see csrc/peer.cu.  `counter_dev`: a zero-initialised
 * uint32 of this rank.  A peer that never arrives traps the launch after a bounded spin instead of hanging. */
int scatt_peer_allgather(const void* src, int64_t bytes, void* const* peer_bufs_host, void* const* peer_flags_host, int world,
                         int rank, void* counter_dev, uint64_t seq, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SCATT_H_ */
