// extern "C" surface of libscatt.so (see include/scatt.h).  Argument checking
// and engine selection only; kernels live in the other translation units.
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace scatt {

std::atomic<uint64_t> g_launches{0};
std::atomic<int> g_pdl{[] {
  const char* e = std::getenv("SCATT_PDL");
  return (e && e[0] == '0') ? 0 : 1;
}()};

namespace {
thread_local char t_error[512] = "";
thread_local char t_kernel[128] = "";
}

void set_last_kernel(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_kernel, sizeof(t_kernel), fmt, ap);
  va_end(ap);
}

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_error, sizeof(t_error), fmt, ap);
  va_end(ap);
}

}  // namespace scatt

using namespace scatt;

namespace {
inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
inline bool fmt_ok(int fmt) { return fmt == SCATT_PLANE_F16 || fmt == SCATT_PLANE_BF16; }
}  // namespace

extern "C" {

int scatt_abi_version(void) { return SCATT_ABI_VERSION; }

const char* scatt_version(void) { return "scatt-b200 0.1 (sm_100a; tcgen05+TMA linear engine, fp32 SIMT engine)"; }

const char* scatt_last_error(void) { return t_error; }

const char* scatt_last_kernel(void) { return t_kernel; }

uint64_t scatt_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int scatt_device_check(void) {
  int dev = 0;
  SCATT_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  SCATT_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10) {
    set_error("device %d is sm_%d%d; libscatt is built for sm_100a only", dev, prop.major, prop.minor);
    return SCATT_ERR_UNSUPPORTED;
  }
  return SCATT_OK;
}

int scatt_debug_set_trace(void* dev_buf) {
  const int rc = debug_set_trace(dev_buf);
  if (rc != SCATT_OK) return rc;
  const int rc2 = debug_set_trace_attention(dev_buf);
  if (rc2 != SCATT_OK) return rc2;
  const int rc3 = debug_set_trace_fa(dev_buf);
  if (rc3 != SCATT_OK) return rc3;
  const int rc4 = debug_set_trace_ctc(dev_buf);
  return rc4 != SCATT_OK ? rc4 : debug_set_trace_block(dev_buf);
}

int scatt_split_planes(const float* x, int64_t rows, int64_t cols, int64_t ldx, float scale, void* planes, int plane_fmt,
                       void* stream) {
  SCATT_REQUIRE(x && planes && fmt_ok(plane_fmt), "split_planes: null pointer or bad plane format");
  return launch_split_planes(x, rows, cols, ldx, scale, planes, plane_fmt, as_stream(stream));
}

int scatt_l2_prefetch(const void* const* ptrs, const int64_t* nbytes, int n, void* stream) {
  SCATT_REQUIRE(n == 0 || (ptrs && nbytes), "l2_prefetch: null argument");
  return launch_l2_prefetch(ptrs, nbytes, n, as_stream(stream));
}

int scatt_frontend(const float* keypoints, int B, int T, int K, int D, const scatt_frontend_stream* streams_host,
                   int n_streams, int max_pos, int plane_fmt, void* stream) {
  SCATT_REQUIRE(keypoints && streams_host && fmt_ok(plane_fmt), "frontend: null pointer or bad plane format");
  SCATT_REQUIRE(B >= 0 && T >= 0 && K >= 1, "frontend: bad shape");
  return launch_frontend(keypoints, B, T, K, D, streams_host, n_streams, max_pos, plane_fmt, as_stream(stream));
}

int scatt_posembed_layernorm(const float* x, const float* table, const float* ln_g, const float* ln_b, float* out,
                             void* out_planes, int B, int T, int D, int max_pos, int plane_fmt, void* stream) {
  SCATT_REQUIRE(x && table && ln_g && ln_b && (out || out_planes) && fmt_ok(plane_fmt), "posembed_layernorm: bad argument");
  SCATT_REQUIRE(T <= max_pos, "posembed_layernorm: T=%d exceeds max_position_embeddings=%d", T, max_pos);
  return launch_posembed_ln(x, table, ln_g, ln_b, out, out_planes, B, T, D, plane_fmt, as_stream(stream));
}

size_t scatt_linear_workspace_bytes(int group, int64_t M, int N, int K, int engine) {
  return (engine == SCATT_ENGINE_TCGEN05 && M >= 0 && N >= 1 && K >= 1) ? linear_tc_workspace_bytes(M, N, K, group) : 0;
}

int scatt_linear(const scatt_linear_problem* problems_host, int group, int64_t M, int N, int K, int64_t ldx, int64_t ldres,
                 int64_t ldy, const scatt_epilogue* epilogue_host, int engine, int plane_fmt, int terms, void* stream) {
  return scatt_linear_ws(problems_host, group, M, N, K, ldx, ldres, ldy, epilogue_host, engine, plane_fmt, terms, nullptr, 0, stream);
}

int scatt_linear_ws(const scatt_linear_problem* problems_host, int group, int64_t M, int N, int K, int64_t ldx, int64_t ldres,
                    int64_t ldy, const scatt_epilogue* epilogue_host, int engine, int plane_fmt, int terms, void* workspace,
                    size_t workspace_bytes, void* stream) {
  SCATT_REQUIRE(problems_host && epilogue_host && fmt_ok(plane_fmt), "linear: null pointer or bad plane format");
  SCATT_REQUIRE(group >= 1 && group <= SCATT_MAX_GROUP, "linear: group must be 1..%d", SCATT_MAX_GROUP);
  SCATT_REQUIRE(M >= 0 && N >= 1 && K >= 1, "linear: bad shape M=%lld N=%d K=%d", (long long)M, N, K);
  if (engine == SCATT_ENGINE_SIMT)
    return launch_linear_simt(problems_host, group, M, N, K, ldx, ldres, ldy, *epilogue_host, plane_fmt, as_stream(stream));
  if (engine == SCATT_ENGINE_TCGEN05)
    return launch_linear_tc(problems_host, group, M, N, K, ldres, ldy, *epilogue_host, plane_fmt, terms, workspace, workspace_bytes,
                            as_stream(stream));
  set_error("linear: unknown engine %d", engine);
  return SCATT_ERR_INVALID;
}

int scatt_attn_block(const scatt_block_problem* problems_host, int group, int64_t M, int D, int F, float ln_eps, int plane_fmt,
                     int terms, void* stream) {
  SCATT_REQUIRE(problems_host && fmt_ok(plane_fmt), "attn_block: null pointer or bad plane format");
  SCATT_REQUIRE(group >= 1 && group <= SCATT_MAX_GROUP, "attn_block: group must be 1..%d", SCATT_MAX_GROUP);
  SCATT_REQUIRE(M >= 0, "attn_block: bad shape M=%lld", (long long)M);
  return launch_attn_block(problems_host, group, M, D, F, ln_eps, plane_fmt, terms, as_stream(stream));
}

int scatt_attn_out_q(const scatt_outq_problem* problems_host, int group, int64_t M, int D, int N, float ln_eps, float q_scale,
                     int plane_fmt, int terms, void* stream) {
  SCATT_REQUIRE(problems_host && fmt_ok(plane_fmt), "attn_out_q: null pointer or bad plane format");
  SCATT_REQUIRE(group >= 1 && group <= SCATT_MAX_GROUP, "attn_out_q: group must be 1..%d", SCATT_MAX_GROUP);
  SCATT_REQUIRE(M >= 0, "attn_out_q: bad shape M=%lld", (long long)M);
  return launch_attn_out_q(problems_host, group, M, D, N, ln_eps, q_scale, plane_fmt, terms, as_stream(stream));
}

int scatt_attn_out_q_supported(int64_t M, int D, int N) { return attn_out_q_supported(M, D, N) ? 1 : 0; }

int scatt_debug_set_block_cluster(int cluster) { return debug_set_block_cluster(cluster); }
int scatt_debug_set_attn_persist(int mode) { return debug_set_attn_persist(mode); }

int scatt_attn_block_supported(int64_t M, int D, int F) { return attn_block_supported(M, D, F) ? 1 : 0; }

int scatt_linear_ln_fused(int64_t M, int N, int group, int engine) {
  if (engine != SCATT_ENGINE_TCGEN05 || M < 0 || N < 1 || group < 1) return 0;
  return linear_tc_ln_cluster(M, N, group, 1) > 0 ? 1 : 0;
}

int scatt_rowwise(const float* z, int64_t M, int N, int64_t ldz, const float* residual, int64_t ldres, const float* ln_g,
                  const float* ln_b, const scatt_epilogue* epilogue_host, float* y, int64_t ldy, void* y_planes,
                  int plane_fmt, void* stream) {
  SCATT_REQUIRE(z && epilogue_host && (y || y_planes) && fmt_ok(plane_fmt), "rowwise: bad argument");
  return launch_rowwise(z, M, N, ldz, residual, ldres, ln_g, ln_b, *epilogue_host, y, ldy, y_planes, plane_fmt,
                        as_stream(stream));
}

int scatt_attention(const scatt_attention_problem* problems_host, int group, int B, int Tq, int Tk, int H, int hd,
                    int64_t ldq, int64_t ldk, int64_t ldv, int kind, int engine, int plane_fmt, int terms, void* stream) {
  SCATT_REQUIRE(problems_host && fmt_ok(plane_fmt), "attention: null pointer or bad plane format");
  SCATT_REQUIRE(kind >= SCATT_ATTN_SELF && kind <= SCATT_ATTN_CROSS, "attention: bad kind %d", kind);
  SCATT_REQUIRE(group >= 1 && group <= SCATT_MAX_GROUP, "attention: group 1..%d", SCATT_MAX_GROUP);
  if (engine == SCATT_ENGINE_TCGEN05 && attention_tc_supported(Tq, Tk, hd, problems_host, group))
    return launch_attention_tc(problems_host, group, B, Tq, Tk, H, hd, ldq, ldk, ldv, kind, plane_fmt, terms,
                               as_stream(stream));
  return launch_attention(problems_host, group, B, Tq, Tk, H, hd, ldq, ldk, ldv, kind, plane_fmt, as_stream(stream));
}

int scatt_attention_planes(const scatt_attention_planes_problem* problems_host, int group, int B, int Tq, int Tk, int H,
                           int hd, int kind, int plane_fmt, int terms, void* stream) {
  SCATT_REQUIRE(problems_host && fmt_ok(plane_fmt), "attention_planes: null pointer or bad plane format");
  SCATT_REQUIRE(kind >= SCATT_ATTN_SELF && kind <= SCATT_ATTN_CROSS, "attention_planes: bad kind %d", kind);
  SCATT_REQUIRE(group >= 1 && group <= SCATT_MAX_GROUP, "attention_planes: group 1..%d", SCATT_MAX_GROUP);
  return launch_attention_planes(problems_host, group, B, Tq, Tk, H, hd, kind, plane_fmt, terms, as_stream(stream));
}

int scatt_fusion_attention(const float* q, const float* k, const float* v, int B, int T, int D, float* out, void* out_planes,
                           int plane_fmt, void* stream) {
  SCATT_REQUIRE(q && k && v && (out || out_planes) && fmt_ok(plane_fmt), "fusion_attention: bad argument");
  return launch_fusion_attention(q, k, v, B, T, D, out, out_planes, plane_fmt, as_stream(stream));
}

int scatt_fusion_attention_planes_supported(int T, int D) { return fusion_attention_tc_supported(T, D) ? 1 : 0; }

int scatt_fusion_attention_planes(const void* q_planes, const void* k_planes, const void* v_planes, int B, int T, int D,
                                  float* out, void* out_planes, int plane_fmt, int terms, void* stream) {
  SCATT_REQUIRE(q_planes && k_planes && v_planes && (out || out_planes) && fmt_ok(plane_fmt) && B >= 0,
                "fusion_attention_planes: bad argument");
  return launch_fusion_attention_tc(q_planes, k_planes, v_planes, B, T, D, out, out_planes, plane_fmt, terms, as_stream(stream));
}

int scatt_pool_pairs(const float* x, int B, int T, int C, float* y, void* y_planes, int plane_fmt, void* stream) {
  SCATT_REQUIRE(x && (y || y_planes) && fmt_ok(plane_fmt), "pool_pairs: bad argument");
  return launch_pool_pairs(x, B, T, C, y, y_planes, plane_fmt, as_stream(stream));
}

int scatt_pool_pairs_group(const float* const* xs_host, float* const* ys_host, void* const* planes_host, int group, int B, int T,
                           int C, int plane_fmt, void* stream) {
  SCATT_REQUIRE(xs_host && ys_host && fmt_ok(plane_fmt), "pool_pairs_group: bad argument");
  return launch_pool_pairs_group(xs_host, ys_host, planes_host, group, B, T, C, plane_fmt, as_stream(stream));
}

size_t scatt_lstm_workspace_bytes(int64_t B, int H) { return B < 0 || H < 0 ? 0 : lstm_workspace_bytes(B, H); }

int scatt_lstm_bidir(const float* gates_x, int64_t ldg, const float* w_hh, float* y, void* y_planes, void* workspace,
                     int64_t B, int T, int H, int plane_fmt, void* stream) {
  SCATT_REQUIRE(fmt_ok(plane_fmt) && T >= 0, "lstm_bidir: bad argument");
  return launch_lstm_bidir(gates_x, ldg, w_hh, y, y_planes, workspace, B, T, H, plane_fmt, as_stream(stream));
}

int scatt_log_softmax(const float* logits, int64_t ld, int V, int B, int T, int time_major, float clamp_min, float clamp_max,
                      float* out, void* stream) {
  SCATT_REQUIRE(B >= 0 && T >= 0, "log_softmax: bad argument");
  return launch_log_softmax(logits, ld, V, B, T, time_major, clamp_min, clamp_max, out, as_stream(stream));
}

int scatt_finite_check(const float* const* tensors_host, const int64_t* sizes_host, int count, int* flags_dev, void* stream) {
  return launch_finite_check(tensors_host, sizes_host, count, flags_dev, as_stream(stream));
}

int scatt_ctc_beam_decode(const float* logits, int B, int T, int V, const int32_t* lengths, int beam, int32_t* out_ids,
                          int32_t* out_len, float* out_score, void* stream) {
  return launch_ctc_beam(logits, B, T, V, lengths, beam, out_ids, out_len, out_score, as_stream(stream));
}

int scatt_peer_allgather(const void* src, int64_t bytes, void* const* peer_bufs_host, void* const* peer_flags_host, int world,
                         int rank, void* counter_dev, uint64_t seq, void* stream) {
  return launch_peer_allgather(src, bytes, peer_bufs_host, peer_flags_host, world, rank, counter_dev, seq, as_stream(stream));
}

}  // extern "C"
