"""ctypes binding of ``libscatt.so`` (the C ABI declared in ``include/scatt.h``).

No torch types cross this boundary: tensors are passed as raw device pointers
plus explicit sizes, the CUDA stream as ``void*``.  There is no CPU fallback:
if the library is missing it is built with nvcc (``scattennet_b200.build``);
if that is impossible, or a compute call is made without a B200, the call
raises.
"""

from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SCATT_LIB", os.path.join(_HERE, "libscatt.so"))  # SCATT_LIB: instrumented dev builds

MAX_GROUP = 4
ENGINE_SIMT, ENGINE_TCGEN05 = 0, 1
PLANE_F16, PLANE_BF16 = 0, 1
ACT_NONE, ACT_GELU, ACT_RELU = 0, 1, 2
RES_NONE, RES_BEFORE_LN, RES_AFTER_LN = 0, 1, 2
ATTN_SELF, ATTN_CAUSAL, ATTN_CROSS = 0, 1, 2

# every symbol include/scatt.h declares (tests check the .so exports all of them)
SYMBOLS = (
    "scatt_abi_version", "scatt_version", "scatt_last_error", "scatt_last_kernel", "scatt_launch_count", "scatt_device_check",
    "scatt_debug_set_trace",
    "scatt_split_planes", "scatt_l2_prefetch", "scatt_frontend", "scatt_posembed_layernorm", "scatt_linear", "scatt_linear_ws", "scatt_linear_workspace_bytes", "scatt_linear_ln_fused", "scatt_attn_block", "scatt_attn_block_supported", "scatt_attn_out_q", "scatt_attn_out_q_supported", "scatt_debug_set_block_cluster", "scatt_debug_set_attn_persist",
    "scatt_rowwise", "scatt_attention", "scatt_attention_planes", "scatt_fusion_attention", "scatt_fusion_attention_planes",
    "scatt_fusion_attention_planes_supported", "scatt_pool_pairs", "scatt_pool_pairs_group",
    "scatt_lstm_workspace_bytes", "scatt_lstm_bidir", "scatt_log_softmax", "scatt_finite_check",
    "scatt_ctc_beam_decode", "scatt_peer_allgather",
)


class ScattError(RuntimeError):
    """A libscatt entry point returned a negative status."""


class Epilogue(C.Structure):
    _fields_ = [
        ("act_pre", C.c_int32), ("residual_mode", C.c_int32), ("layer_norm", C.c_int32), ("act_post", C.c_int32),
        ("clamp", C.c_float), ("scale_cols", C.c_int32), ("scale", C.c_float), ("ln_eps", C.c_float),
    ]


class LinearProblem(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("x_planes", C.c_void_p), ("w", C.c_void_p), ("w_planes", C.c_void_p), ("bias", C.c_void_p),
        ("residual", C.c_void_p), ("ln_g", C.c_void_p), ("ln_b", C.c_void_p), ("y", C.c_void_p), ("y_planes", C.c_void_p), ("residual_planes", C.c_void_p),
    ]


class OutQProblem(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("ctx_planes", "residual_planes", "wo_planes", "bo", "ln_g", "ln_b", "wq_planes", "bq",
                                          "h_planes", "q_planes")]


class BlockProblem(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("ctx_planes", "residual_planes", "wo_planes", "bo", "ln1_g", "ln1_b", "w1_planes", "b1",
                                          "w2_planes", "b2", "ln2_g", "ln2_b", "y", "y_planes")]


class AttentionProblem(C.Structure):
    _fields_ = [
        ("q", C.c_void_p), ("k", C.c_void_p), ("v", C.c_void_p), ("key_mask", C.c_void_p), ("additive", C.c_void_p),
        ("out", C.c_void_p), ("out_planes", C.c_void_p),
    ]


class AttnOperand(C.Structure):
    _fields_ = [("planes", C.c_void_p), ("rows", C.c_int64), ("ld", C.c_int64), ("col", C.c_int32), ("reserved", C.c_int32)]


class AttentionPlanesProblem(C.Structure):
    _fields_ = [("q", AttnOperand), ("k", AttnOperand), ("v", AttnOperand), ("key_mask", C.c_void_p), ("out", C.c_void_p),
                ("out_planes", C.c_void_p)]


class FrontendStream(C.Structure):
    _fields_ = [
        ("joint_idx", C.c_void_p), ("n_joints", C.c_int32), ("coord", C.c_int32 * 2), ("map_wt", C.c_void_p * 2),
        ("map_b", C.c_void_p * 2), ("pos", C.c_void_p * 2), ("ln_g", C.c_void_p * 2), ("ln_b", C.c_void_p * 2),
        ("out", C.c_void_p * 2), ("out_planes", C.c_void_p * 2), ("gathered", C.c_void_p),
    ]


_lock = threading.Lock()
_lib = None


def _declare(lib):
    i32, i64, f32, vp = C.c_int, C.c_int64, C.c_float, C.c_void_p
    lib.scatt_abi_version.restype = i32
    lib.scatt_version.restype = C.c_char_p
    lib.scatt_last_error.restype = C.c_char_p
    lib.scatt_last_kernel.restype = C.c_char_p
    lib.scatt_launch_count.restype = C.c_uint64
    lib.scatt_device_check.restype = i32
    lib.scatt_debug_set_trace.argtypes = [vp]
    lib.scatt_debug_set_trace.restype = i32
    lib.scatt_split_planes.argtypes = [vp, i64, i64, i64, f32, vp, i32, vp]
    lib.scatt_l2_prefetch.argtypes = [C.POINTER(vp), C.POINTER(C.c_int64), i32, vp]
    lib.scatt_frontend.argtypes = [vp, i32, i32, i32, i32, C.POINTER(FrontendStream), i32, i32, i32, vp]
    lib.scatt_posembed_layernorm.argtypes = [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, vp]
    lib.scatt_linear.argtypes = [C.POINTER(LinearProblem), i32, i64, i32, i32, i64, i64, i64, C.POINTER(Epilogue), i32, i32,
                                 i32, vp]
    lib.scatt_linear_ws.argtypes = [C.POINTER(LinearProblem), i32, i64, i32, i32, i64, i64, i64, C.POINTER(Epilogue), i32, i32,
                                    i32, vp, C.c_size_t, vp]
    lib.scatt_linear_ws.restype = i32
    lib.scatt_linear_workspace_bytes.argtypes = [i32, i64, i32, i32, i32]
    lib.scatt_linear_workspace_bytes.restype = C.c_size_t
    lib.scatt_attn_block.argtypes = [C.POINTER(BlockProblem), i32, i64, i32, i32, f32, i32, i32, vp]
    lib.scatt_attn_block.restype = i32
    lib.scatt_attn_out_q.argtypes = [C.POINTER(OutQProblem), i32, i64, i32, i32, f32, f32, i32, i32, vp]
    lib.scatt_attn_out_q_supported.argtypes = [i64, i32, i32]
    lib.scatt_attn_out_q_supported.restype = i32
    lib.scatt_attn_block_supported.argtypes = [i64, i32, i32]
    lib.scatt_attn_block_supported.restype = i32
    lib.scatt_debug_set_block_cluster.argtypes = [i32]
    lib.scatt_debug_set_block_cluster.restype = i32
    lib.scatt_debug_set_attn_persist.argtypes = [i32]
    lib.scatt_debug_set_attn_persist.restype = i32
    lib.scatt_linear_ln_fused.argtypes = [i64, i32, i32, i32]
    lib.scatt_linear_ln_fused.restype = i32
    lib.scatt_rowwise.argtypes = [vp, i64, i32, i64, vp, i64, vp, vp, C.POINTER(Epilogue), vp, i64, vp, i32, vp]
    lib.scatt_attention.argtypes = [C.POINTER(AttentionProblem), i32, i32, i32, i32, i32, i32, i64, i64, i64, i32, i32, i32,
                                    i32, vp]
    lib.scatt_attention_planes.argtypes = [C.POINTER(AttentionPlanesProblem), i32, i32, i32, i32, i32, i32, i32, i32, i32, vp]
    lib.scatt_attention_planes.restype = i32
    lib.scatt_fusion_attention.argtypes = [vp, vp, vp, i32, i32, i32, vp, vp, i32, vp]
    lib.scatt_fusion_attention_planes.argtypes = [vp, vp, vp, i32, i32, i32, vp, vp, i32, i32, vp]
    lib.scatt_fusion_attention_planes.restype = i32
    lib.scatt_fusion_attention_planes_supported.argtypes = [i32, i32]
    lib.scatt_fusion_attention_planes_supported.restype = i32
    lib.scatt_pool_pairs.argtypes = [vp, i32, i32, i32, vp, vp, i32, vp]
    lib.scatt_pool_pairs_group.argtypes = [C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), i32, i32, i32, i32, i32, vp]
    lib.scatt_pool_pairs_group.restype = i32
    lib.scatt_lstm_workspace_bytes.argtypes = [i64, i32]
    lib.scatt_lstm_workspace_bytes.restype = C.c_size_t
    lib.scatt_lstm_bidir.argtypes = [vp, i64, vp, vp, vp, vp, i64, i32, i32, i32, vp]
    lib.scatt_lstm_bidir.restype = i32
    lib.scatt_log_softmax.argtypes = [vp, i64, i32, i32, i32, i32, f32, f32, vp, vp]
    lib.scatt_log_softmax.restype = i32
    lib.scatt_finite_check.argtypes = [C.POINTER(vp), C.POINTER(i64), i32, vp, vp]
    lib.scatt_finite_check.restype = i32
    lib.scatt_ctc_beam_decode.argtypes = [vp, i32, i32, i32, vp, i32, vp, vp, vp, vp]
    lib.scatt_ctc_beam_decode.restype = i32
    lib.scatt_peer_allgather.argtypes = [vp, i64, C.POINTER(vp), C.POINTER(vp), i32, i32, vp, C.c_uint64, vp]
    lib.scatt_peer_allgather.restype = i32
    for name in ("scatt_split_planes", "scatt_frontend", "scatt_posembed_layernorm", "scatt_linear", "scatt_rowwise",
                 "scatt_attention", "scatt_fusion_attention", "scatt_pool_pairs"):
        getattr(lib, name).restype = i32


def load(build_if_missing: bool = True):
    """Load (building first if needed) and return the ctypes handle."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            if not build_if_missing:
                raise ScattError(f"{LIB_PATH} is missing and there is no CPU fallback; run python -m scattennet_b200.build")
            from . import build as _build

            _build.build()
        lib = C.CDLL(LIB_PATH)
        missing = [s for s in SYMBOLS if not hasattr(lib, s)]
        if missing:
            raise ScattError(f"{LIB_PATH} lacks symbols {missing}")
        _declare(lib)
        if lib.scatt_abi_version() != 3:
            raise ScattError("libscatt ABI version mismatch; rebuild with python -m scattennet_b200.build --force")
        _lib = lib
        return lib


def check(status: int, what: str):
    if status != 0:
        msg = load().scatt_last_error().decode(errors="replace")
        raise ScattError(f"{what} failed ({status}): {msg}")


def launch_count() -> int:
    return int(load().scatt_launch_count())
