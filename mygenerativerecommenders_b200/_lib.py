"""ctypes binding of libgrb200.so (the C ABI in include/grb200.h).

There is deliberately no fallback of any kind: if the shared library is missing or a call
fails, the error is raised to the caller.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import torch

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libgrb200.so"

GRB_OK = 0
GRB_ERR_INVALID_ARG = -1
GRB_ERR_UNSUPPORTED = -2
GRB_ERR_CUDA = -3
GRB_ERR_WORKSPACE = -4
GRB_F32 = 0
GRB_BF16 = 1

c_i64 = C.c_int64
c_i32 = C.c_int32
c_vp = C.c_void_p


class HstuAttnArgs(C.Structure):
    _fields_ = [
        ("B", c_i64), ("N", c_i64), ("T", c_i64), ("max_len", c_i64),
        ("H", c_i32), ("dqk", c_i32), ("dv", c_i32),
        ("dtype", c_i32), ("index_bits", c_i32), ("num_buckets", c_i32),
        ("q", c_vp), ("k", c_vp), ("v", c_vp),
        ("ldq", c_i64), ("ldk", c_i64), ("ldv", c_i64),
        ("offsets", c_vp), ("timestamps", c_vp), ("ts_w", c_vp), ("pos_w", c_vp),
        ("bucket_thresholds", c_vp),
        ("out", c_vp), ("ldo", c_i64),
        ("dout", c_vp), ("lddo", c_i64),
        ("dq", c_vp), ("dk", c_vp), ("dv_grad", c_vp),
        ("lddq", c_i64), ("lddk", c_i64), ("lddv", c_i64),
        ("dq_accum", c_vp), ("d_ts_w", c_vp), ("d_pos_w", c_vp), ("d_bias_copies", c_i32),
        ("bucket_octaves", c_vp), ("bucket_cache", c_vp), ("bucket_cache_max_len", c_i64),
        ("short_schedule", c_vp), ("bucket_cache_masked", c_i32), ("zero_tail_rows", c_i32),
    ]


class HstuAttnDecodeArgs(C.Structure):
    _fields_ = [
        ("B", c_i64), ("N", c_i64), ("H", c_i32), ("dqk", c_i32), ("dv", c_i32),
        ("dtype", c_i32), ("index_bits", c_i32), ("pos_bits", c_i32), ("num_buckets", c_i32),
        ("q", c_vp), ("ldq", c_i64), ("k_cache", c_vp), ("ldk", c_i64), ("v", c_vp), ("ldv", c_i64),
        ("offsets", c_vp), ("positions", c_vp), ("timestamps", c_vp),
        ("ts_w", c_vp), ("pos_w", c_vp), ("bucket_thresholds", c_vp),
        ("out", c_vp), ("ldo", c_i64),
    ]


class JaggedInputArgs(C.Structure):
    _fields_ = [
        ("B", c_i64), ("N", c_i64), ("V", c_i64), ("rows", c_i64),
        ("D", c_i32), ("dtype", c_i32), ("index_bits", c_i32), ("reserved", c_i32),
        ("table", c_vp), ("ldt", c_i64), ("ids", c_vp), ("offsets", c_vp),
        ("pos", c_vp), ("ldp", c_i64),
        ("scale", C.c_float), ("p_drop", C.c_float), ("seed", c_vp),
        ("io", c_vp), ("ldio", c_i64), ("d_table", c_vp), ("d_pos", c_vp),
    ]


class ProjGemmArgs(C.Structure):
    _fields_ = [
        ("M", c_i64), ("N", c_i64), ("K", c_i64),
        ("a_mn", c_i32), ("b_mn", c_i32), ("epi", c_i32), ("reserved", c_i32),
        ("A", c_vp), ("lda", c_i64), ("B", c_vp), ("ldb", c_i64),
        ("out0", c_vp), ("ldo0", c_i64), ("out1", c_vp), ("ldo1", c_i64),
        ("bias", c_vp), ("res", c_vp), ("ldres", c_i64),
    ]


GEMM_EPI_PLAIN, GEMM_EPI_SILU2, GEMM_EPI_BIAS_RES, GEMM_EPI_F32_ADD = 0, 1, 2, 3


class MipsTopkArgs(C.Structure):
    _fields_ = [
        ("B", c_i64), ("X", c_i64), ("D", c_i64),
        ("k", c_i32), ("dtype", c_i32),
        ("queries", c_vp), ("ldq", c_i64),
        ("items", c_vp), ("ldi", c_i64),
        ("item_ids", c_vp),
        ("out_scores", c_vp), ("out_ids", c_vp),
        ("workspace", c_vp), ("workspace_bytes", c_i64),
        ("sample_stride", c_i64), ("cand_cap", c_i64),
        ("status", c_vp),
        ("invalid_ids", c_vp), ("ld_invalid", c_i64), ("n_invalid", c_i32),
        ("target_ids", c_vp), ("out_ranks", c_vp),
    ]


class LnGateArgs(C.Structure):
    _fields_ = [
        ("x", c_vp), ("ldx", c_i64),
        ("gate", c_vp), ("ldg", c_i64),
        ("y", c_vp), ("ldy", c_i64),
        ("mean", c_vp), ("rstd", c_vp),
        ("rows", c_i64), ("W", c_i64),
        ("eps", C.c_float), ("dtype", c_i32),
        ("p_drop", C.c_float), ("_pad0", c_i32),
        ("seed", c_vp), ("salt", c_i64),
        ("dx", c_vp), ("lddx", c_i64),
        ("dgate", c_vp), ("lddg", c_i64),
        ("res", c_vp), ("ldres", c_i64),
    ]


class SslArgs(C.Structure):
    _fields_ = [
        ("n_rows", c_i64),
        ("R", c_i32), ("D", c_i32), ("d0", c_i32), ("d1", c_i32),
        ("l2_norm", c_i32), ("dtype", c_i32),
        ("l2_eps", C.c_float), ("temperature", C.c_float),
        ("q", c_vp), ("ldq_", c_i64),
        ("p", c_vp), ("ldp", c_i64),
        ("table0", c_vp), ("ldt0", c_i64),
        ("table1", c_vp), ("ldt1", c_i64),
        ("idx0", c_vp), ("idx1", c_vp), ("pos_ids", c_vp), ("neg_ids", c_vp),
        ("loss_rows", c_vp), ("probs", c_vp),
        ("g", c_vp), ("dq", c_vp), ("dp", c_vp), ("dtable0", c_vp), ("dtable1", c_vp),
    ]


# name -> (restype, argtypes); every symbol include/grb200.h declares
SYMBOLS = {
    "grb_version": (C.c_int, []),
    "grb_last_error_string": (C.c_char_p, []),
    "grb_launch_count": (c_i64, []),
    "grb_complete_cumsum": (C.c_int, [c_vp, c_vp, c_i64, C.c_int, c_vp]),
    "grb_dense_to_jagged": (C.c_int, [c_vp, c_vp, c_vp, c_i64, c_i64, c_i64, c_i64, C.c_int, c_vp]),
    "grb_jagged_to_padded_dense": (
        C.c_int, [c_vp, c_vp, c_vp, c_i64, c_i64, c_i64, c_i64, c_vp, C.c_int, C.c_int, c_vp]),
    "grb_gather_last_rows": (C.c_int, [c_vp, c_vp, c_vp, c_i64, c_i64, c_i64, C.c_int, C.c_int, c_vp]),
    "grb_bucket_octaves": (C.c_int, [c_vp, c_i32, c_vp]),
    "grb_hstu_bucket_cache_bytes": (c_i64, [c_i64, c_i64]),
    "grb_hstu_bucket_tiles_masked": (C.c_int, [c_vp, C.c_int, c_vp, c_i64, c_i64, c_i64, c_vp, c_i32, c_vp, c_vp, c_vp]),
    "grb_hstu_short_schedule": (C.c_int, [c_vp, C.c_int, c_i64, c_i64, c_vp, c_vp]),
    "grb_hstu_bucket_tiles": (C.c_int, [c_vp, C.c_int, c_vp, c_i64, c_i64, c_i64, c_vp, c_i32, c_vp, c_vp, c_vp]),
    "grb_hstu_attn_fwd": (C.c_int, [C.POINTER(HstuAttnArgs), c_vp]),
    "grb_hstu_attn_bwd": (C.c_int, [C.POINTER(HstuAttnArgs), c_vp]),
    "grb_ln_gate_fwd": (
        C.c_int, [c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_vp, c_vp, c_i64, c_i64, C.c_float,
                  C.c_int, c_vp]),
    "grb_ln_gate_bwd": (
        C.c_int, [c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_vp, c_vp, c_vp, c_i64, c_vp, c_i64,
                  c_i64, c_i64, C.c_int, c_vp]),
    "grb_ln_gate_fwd_ex": (C.c_int, [C.POINTER(LnGateArgs), c_vp]),
    "grb_ln_gate_bwd_ex": (C.c_int, [C.POINTER(LnGateArgs), c_vp]),
    "grb_mips_topk_workspace_bytes": (c_i64, [C.POINTER(MipsTopkArgs)]),
    "grb_mips_topk": (C.c_int, [C.POINTER(MipsTopkArgs), c_vp]),
    "grb_topk_select": (
        C.c_int, [c_vp, c_vp, c_vp, c_i64, c_i64, c_i32, c_vp, c_vp, c_vp, c_vp]),
    "grb_sampled_softmax_fwd": (C.c_int, [C.POINTER(SslArgs), c_vp]),
    "grb_sampled_softmax_bwd": (C.c_int, [C.POINTER(SslArgs), c_vp]),
    "grb_sampled_softmax_bwd_csr_workspace_bytes": (c_i64, [c_i64, c_i32, c_i32, c_i64]),
    "grb_sampled_softmax_bwd_csr": (C.c_int, [C.POINTER(SslArgs), c_i64, c_vp, c_i64, c_vp]),
    "grb_l2norm_fwd": (C.c_int, [c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_i64, C.c_float, c_vp]),
    "grb_l2norm_bwd": (C.c_int, [c_vp, c_i64, c_vp, c_i64, c_vp, c_vp, c_i64, c_i64, c_i64, c_vp]),
    "grb_p2p_put_rows": (C.c_int, [c_vp, c_i64, C.POINTER(c_vp), c_i32, c_i64, c_i64, c_i64, c_i64, c_vp]),
    "grb_p2p_barrier": (C.c_int, [C.POINTER(c_vp), c_i32, c_i32, c_i32, c_i64, c_vp]),
    "grb_p2p_allreduce": (C.c_int, [C.POINTER(c_vp), c_i32, c_i32, c_i64, C.c_float, c_vp]),
    "grb_p2p_put_table_rows": (C.c_int, [c_vp, c_vp, c_i64, c_i32, c_i64, c_i64, C.c_float, C.POINTER(c_vp),
                                         C.POINTER(c_vp), c_i32, c_i64, c_vp]),
    "grb_hstu_attn_decode": (C.c_int, [C.POINTER(HstuAttnDecodeArgs), c_vp]),
    "grb_silu_fwd": (C.c_int, [c_vp, c_i64, c_vp, c_i64, c_i64, c_i32, c_i32, c_vp]),
    "grb_silu_split_bwd": (C.c_int, [c_vp, c_i64, c_i32, C.POINTER(c_vp), C.POINTER(c_i64), C.POINTER(c_i32),
                                     c_vp, c_i64, c_i64, c_i32, c_vp]),
    "grb_adamw_step": (C.c_int, [C.c_int, C.POINTER(c_vp), C.POINTER(c_vp), C.POINTER(c_vp), C.POINTER(c_vp),
                                 C.POINTER(c_i64)] + [C.c_double] * 7 + [c_vp]),
    "grb_weighted_mean_fwd": (C.c_int, [c_vp, c_vp, c_i64, c_vp, c_vp]),
    "grb_weighted_mean_bwd": (C.c_int, [c_vp, c_vp, c_vp, c_i64, c_vp, c_vp]),
    "grb_cast_f32_bf16_many": (C.c_int, [C.c_int, C.POINTER(c_vp), C.POINTER(c_vp), C.POINTER(c_i64), c_vp]),
    "grb_rows_scatter_add": (C.c_int, [c_vp, c_i64, c_vp, c_vp, c_i64, c_i32, c_i64, c_i64, c_vp]),
    "grb_rows_scale": (C.c_int, [c_vp, c_vp, c_i64, c_i32, c_i64, c_i64, C.c_float, c_vp]),
    "grb_zero_tail_rows": (C.c_int, [c_vp, c_i64, c_i64, c_i64, c_i32, c_i64, c_vp, c_i32, c_i64, c_i64, c_vp]),
    "grb_draw_negatives": (C.c_int, [c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp]),
    "grb_inbatch_distinct_ids": (C.c_int, [c_vp, c_i64, c_i64, c_vp, c_i32, c_i32, c_vp, c_i64, c_vp, c_vp,
                                           c_i64, c_vp, c_vp]),
    "grb_jagged_input_fwd": (C.c_int, [C.POINTER(JaggedInputArgs), c_vp]),
    "grb_jagged_input_bwd": (C.c_int, [C.POINTER(JaggedInputArgs), c_vp]),
    "grb_l2norm_cast_fwd": (C.c_int, [c_vp, c_i64, C.c_int, c_vp, c_i64, c_vp, c_i64, c_i64, C.c_float, c_vp]),
    "grb_l2norm_cast_bwd": (C.c_int, [c_vp, c_i64, c_vp, c_i64, c_vp, c_vp, c_i64, C.c_int, c_i64, c_i64, c_vp]),
    "grb_proj_gemm": (C.c_int, [C.POINTER(ProjGemmArgs), c_vp]),
    "grb_colsum_bf16": (C.c_int, [c_vp, c_i64, c_i64, c_i32, c_vp, c_vp]),
    "grb_colsum_f32_pair": (C.c_int, [c_vp, c_i32, c_vp, c_vp, c_i32, c_vp, c_i32, c_vp]),
    "grb_selftest_umma": (C.c_int, [C.POINTER(C.c_float), C.c_int, c_vp]),
}

_lib = None


def lib() -> C.CDLL:
    """The loaded library.  Raises if it has not been built (python -m ...build)."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with "
                "`python -m mygenerativerecommenders_b200.build` (needs nvcc). "
                "This package has no CPU or PyTorch fallback."
            )
        handle = C.CDLL(str(LIB_PATH))
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(handle, name)  # AttributeError if the .so lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc: int) -> None:
    if rc == GRB_OK:
        return
    msg = lib().grb_last_error_string().decode("utf-8", "replace")
    if rc == GRB_ERR_INVALID_ARG:
        raise ValueError(f"grb200: {msg}")
    if rc == GRB_ERR_UNSUPPORTED:
        raise NotImplementedError(f"grb200: {msg}")
    raise RuntimeError(f"grb200 (code {rc}): {msg}")


def stream_ptr(device: torch.device | None = None) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def require_cuda(*tensors: torch.Tensor) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError(
                "grb200 ops run on CUDA tensors only (no CPU fallback); got a "
                f"{t.device} tensor of shape {tuple(t.shape)}"
            )


def dtype_code(dt: torch.dtype) -> int:
    if dt == torch.float32:
        return GRB_F32
    if dt == torch.bfloat16:
        return GRB_BF16
    raise NotImplementedError(f"grb200: unsupported dtype {dt} (float32 / bfloat16 only)")


def index_bits(t: torch.Tensor) -> int:
    if t.dtype == torch.int32:
        return 32
    if t.dtype == torch.int64:
        return 64
    raise ValueError(f"grb200: offsets/lengths must be int32 or int64, got {t.dtype}")


def ptr(t: torch.Tensor | None) -> int | None:
    return None if t is None else t.data_ptr()


def launch_count() -> int:
    return int(lib().grb_launch_count())


# --------------------------------------------------------------------------------------------
# optional per-kernel timing (bench.py): CUDA events on the launching stream around one call
# --------------------------------------------------------------------------------------------
_PROFILE = None  # name -> list of (start_event, end_event)


def profile_start() -> None:
    global _PROFILE
    _PROFILE = {}


def profile_stop() -> dict:
    """Returns name -> (launch_groups, total_ms).  Synchronises the device."""
    global _PROFILE
    prof, _PROFILE = _PROFILE, None
    torch.cuda.synchronize()
    out = {}
    for name, evs in (prof or {}).items():
        out[name] = (len(evs), sum(s.elapsed_time(e) for s, e in evs))
    return out


class timed:
    """``with timed("name"):`` around a C-ABI call; free when profiling is off."""

    __slots__ = ("name", "start")

    def __init__(self, name: str):
        self.name = name
        self.start = None

    def __enter__(self):
        if _PROFILE is not None:
            self.start = torch.cuda.Event(enable_timing=True)
            self.start.record()
        return self

    def __exit__(self, *exc):
        if self.start is not None and _PROFILE is not None:
            end = torch.cuda.Event(enable_timing=True)
            end.record()
            _PROFILE.setdefault(self.name, []).append((self.start, end))
        return False
