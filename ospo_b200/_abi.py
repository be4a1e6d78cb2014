"""ctypes binding of the C ABI in ``include/ospo_head.h`` (``ospo_b200/lib/libospo_head.so``).

This module is plumbing only: it mirrors the C structs, loads the shared library and turns a
non-zero status into a Python exception.  There is no CPU fallback -- if the library is missing
or the device is not sm_100, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_LIB_PATH = Path(__file__).resolve().parent / "lib" / "libospo_head.so"

c_float_p = C.POINTER(C.c_float)
c_int64_p = C.POINTER(C.c_int64)


class Shape(C.Structure):
    _fields_ = [
        ("rows", C.c_int32),
        ("hidden", C.c_int32),
        ("embed", C.c_int32),
        ("vocab", C.c_int32),
        ("num_seqs", C.c_int32),
    ]


class Weights(C.Structure):
    _fields_ = [
        ("w1", C.c_void_p),
        ("b1", C.c_void_p),
        ("w2", C.c_void_p),
        ("b2", C.c_void_p),
    ]


class HeadArgs(C.Structure):
    _fields_ = [
        ("shape", Shape),
        ("w", Weights),
        ("x", C.c_void_p),
        ("logits", C.c_void_p),
        ("workspace", C.c_void_p),
        ("workspace_bytes", C.c_size_t),
    ]


class DpExchange(C.Structure):
    _fields_ = [
        ("world", C.c_int32),
        ("rank", C.c_int32),
        ("inbox", C.c_void_p * 8),
        ("flat", C.c_void_p * 8),
        ("flat_multicast", C.c_void_p),
    ]


class SimpoArgs(C.Structure):
    _fields_ = [
        ("shape", Shape),
        ("w", Weights),
        ("x", C.c_void_p),
        ("x_seg_rows", C.c_int32),
        ("x_seg_pitch", C.c_int32),
        ("x_seg_off", C.c_int32),
        ("labels", C.c_void_p),
        ("seq_offsets", C.c_void_p),
        ("average_log_prob", C.c_int32),
        ("beta", C.c_float),
        ("gamma_beta_ratio", C.c_float),
        ("label_smoothing", C.c_float),
        ("sft_weight", C.c_float),
        ("loss_type", C.c_int32),
        ("row_logps", C.c_void_p),
        ("seq_logps", C.c_void_p),
        ("losses", C.c_void_p),
        ("chosen_rewards", C.c_void_p),
        ("rejected_rewards", C.c_void_p),
        ("scalars", C.c_void_p),
        ("pre", C.c_void_p),
        ("act", C.c_void_p),
        ("logits", C.c_void_p),
        ("row_lse", C.c_void_p),
        ("row_ref", C.c_void_p),
        ("grad_seq", C.c_void_p),
        ("grad_loss", C.c_void_p),
        ("dx", C.c_void_p),
        ("flat_grads", C.c_void_p),
        ("workspace", C.c_void_p),
        ("workspace_bytes", C.c_size_t),
        ("bwd_stage", C.c_int32),
        ("reserve_sms", C.c_int32),
        ("wgrad_scale", C.c_float),
        ("dp", C.c_void_p),              # const ospo_dp_exchange*
    ]


class CfgArgs(C.Structure):
    _fields_ = [
        ("shape", Shape),
        ("w", Weights),
        ("h", C.c_void_p),
        ("logits", C.c_void_p),
        ("cfg_weight", C.c_float),
        ("temperature", C.c_float),
        ("merge_mode", C.c_int32),
        ("greedy", C.c_int32),
        ("num_steps", C.c_int32),
        ("uniforms", C.c_void_p),
        ("ids", C.c_void_p),
        ("merged", C.c_void_p),
        ("workspace", C.c_void_p),
        ("workspace_bytes", C.c_size_t),
        ("w1_packed", C.c_void_p),
        ("w2_packed", C.c_void_p),
        ("next_embeds", C.c_void_p),     # const ospo_aligner_args*
    ]


class AlignerArgs(C.Structure):
    _fields_ = [
        ("rows", C.c_int32),
        ("embed", C.c_int32),
        ("codebook", C.c_int32),
        ("code_dim", C.c_int32),
        ("ids", C.c_void_p),
        ("gen_embed", C.c_void_p),
        ("wa", C.c_void_p),
        ("ba", C.c_void_p),
        ("wb", C.c_void_p),
        ("bb", C.c_void_p),
        ("out", C.c_void_p),
        ("workspace", C.c_void_p),
        ("workspace_bytes", C.c_size_t),
        ("id_repeat", C.c_int32),
        ("table", C.c_void_p),
    ]


class AdamWArgs(C.Structure):
    _fields_ = [
        ("numel", C.c_int64),
        ("grads", C.c_void_p),
        ("params", C.c_void_p),
        ("exp_avg", C.c_void_p),
        ("exp_avg_sq", C.c_void_p),
        ("params_bf16", C.c_void_p),
        ("shadow_numel", C.c_int64),
        ("lr", C.c_double),
        ("beta1", C.c_double),
        ("beta2", C.c_double),
        ("eps", C.c_double),
        ("weight_decay", C.c_double),
        ("step", C.c_int32),
        ("max_norm", C.c_float),
        ("total_sqnorm", C.c_void_p),
    ]


# scalars[] indices (OSPO_SC_*)
SC_LOSS, SC_SIMPO_LOSS, SC_SFT_LOSS = 0, 1, 2
SC_REWARD_CHOSEN, SC_REWARD_REJECTED, SC_REWARD_ACC, SC_REWARD_MARGIN = 3, 4, 5, 6
SC_LOGPS_CHOSEN, SC_LOGPS_REJECTED, SC_LOGITS_CHOSEN, SC_LOGITS_REJECTED = 7, 8, 9, 10
SC_COUNT = 16

LOSS_SIGMOID, LOSS_HINGE = 0, 1
MERGE_BF16, MERGE_FP32 = 0, 1

# every symbol include/ospo_head.h declares (tests check the library exports exactly these)
EXPORTS = (
    "ospo_head_workspace_bytes",
    "ospo_head_logits",
    "ospo_head_logps_fwd",
    "ospo_head_logps_bwd",
    "ospo_head_simpo_fwd",
    "ospo_head_simpo_bwd",
    "ospo_head_dp_reduce_broadcast",
    "ospo_head_packed_weight_bytes",
    "ospo_head_pack_weight",
    "ospo_head_cfg_sample",
    "ospo_head_cfg_merge_sample",
    "ospo_head_gen_img_embeds",
    "ospo_head_grad_sqnorm",
    "ospo_head_adamw_step",
    "ospo_head_strerror",
    "ospo_head_set_cta_group",
    "ospo_head_set_decode_mode",
    "ospo_head_set_decode_merged",
    "ospo_head_set_decode_l2_ahead",
    "ospo_head_set_group_m",
    "ospo_head_set_kernel_tune",
    "ospo_head_set_wgrad_splitk",
    "ospo_head_profile_enable",
    "ospo_head_profile_read",
    "ospo_head_trace",
    "ospo_head_launch_count",
    "ospo_head_watchdog_record_host",
    "ospo_head_gemm_debug",
)


class OspoHeadError(RuntimeError):
    pass


_lib = None


def lib_path() -> Path:
    return Path(os.environ.get("OSPO_HEAD_LIB", str(_LIB_PATH)))


def load() -> C.CDLL:
    """Load libospo_head.so (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    p = lib_path()
    if not p.exists():
        raise OspoHeadError(
            f"{p} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C ospo_b200/csrc -j`.  There is no fallback implementation."
        )
    lib = C.CDLL(str(p))
    S = C.c_void_p  # stream handle
    lib.ospo_head_workspace_bytes.argtypes = [C.POINTER(Shape), C.POINTER(C.c_size_t)]
    lib.ospo_head_workspace_bytes.restype = C.c_int
    lib.ospo_head_logits.argtypes = [C.POINTER(HeadArgs), S]
    lib.ospo_head_logits.restype = C.c_int
    for name in ("ospo_head_logps_fwd", "ospo_head_logps_bwd", "ospo_head_simpo_fwd", "ospo_head_simpo_bwd"):
        fn = getattr(lib, name)
        fn.argtypes = [C.POINTER(SimpoArgs), S]
        fn.restype = C.c_int
    lib.ospo_head_dp_reduce_broadcast.argtypes = [C.POINTER(Shape), C.POINTER(DpExchange), C.c_int32, C.c_int32, S]
    lib.ospo_head_dp_reduce_broadcast.restype = C.c_int
    for name in ("ospo_head_cfg_sample", "ospo_head_cfg_merge_sample"):
        fn = getattr(lib, name)
        fn.argtypes = [C.POINTER(CfgArgs), S]
        fn.restype = C.c_int
    lib.ospo_head_gen_img_embeds.argtypes = [C.POINTER(AlignerArgs), S]
    lib.ospo_head_gen_img_embeds.restype = C.c_int
    lib.ospo_head_packed_weight_bytes.argtypes = [C.c_int32, C.c_int32, C.POINTER(C.c_size_t)]
    lib.ospo_head_packed_weight_bytes.restype = C.c_int
    lib.ospo_head_pack_weight.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, S]
    lib.ospo_head_pack_weight.restype = C.c_int
    lib.ospo_head_grad_sqnorm.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_size_t, S]
    lib.ospo_head_grad_sqnorm.restype = C.c_int
    lib.ospo_head_adamw_step.argtypes = [C.POINTER(AdamWArgs), S]
    lib.ospo_head_adamw_step.restype = C.c_int
    lib.ospo_head_strerror.argtypes = [C.c_int]
    lib.ospo_head_strerror.restype = C.c_char_p
    lib.ospo_head_set_cta_group.argtypes = [C.c_int]
    lib.ospo_head_set_cta_group.restype = C.c_int
    lib.ospo_head_set_decode_mode.argtypes = [C.c_int, C.c_int]
    lib.ospo_head_set_decode_mode.restype = C.c_int
    lib.ospo_head_set_decode_merged.argtypes = [C.c_int]
    lib.ospo_head_set_decode_merged.restype = C.c_int
    lib.ospo_head_set_decode_l2_ahead.argtypes = [C.c_int]
    lib.ospo_head_set_decode_l2_ahead.restype = C.c_int
    lib.ospo_head_set_group_m.argtypes = [C.c_int]
    lib.ospo_head_set_group_m.restype = C.c_int
    lib.ospo_head_set_kernel_tune.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int]
    lib.ospo_head_set_kernel_tune.restype = C.c_int
    lib.ospo_head_set_wgrad_splitk.argtypes = [C.c_int]
    lib.ospo_head_set_wgrad_splitk.restype = C.c_int
    lib.ospo_head_profile_enable.argtypes = [C.c_int]
    lib.ospo_head_profile_enable.restype = C.c_int
    lib.ospo_head_profile_read.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_int32), C.c_int32]
    lib.ospo_head_profile_read.restype = C.c_int
    lib.ospo_head_trace.argtypes = [C.c_void_p]
    lib.ospo_head_trace.restype = C.c_int
    lib.ospo_head_launch_count.argtypes = []
    lib.ospo_head_launch_count.restype = C.c_uint64
    lib.ospo_head_watchdog_record_host.argtypes = []
    lib.ospo_head_watchdog_record_host.restype = C.POINTER(C.c_uint32)
    lib.ospo_head_gemm_debug.argtypes = [
        C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
        C.c_int32, C.c_int32, C.c_int32, S,
    ]
    lib.ospo_head_gemm_debug.restype = C.c_int
    _lib = lib
    return lib


def strerror(status: int) -> str:
    return load().ospo_head_strerror(status).decode()


def check(status: int, what: str) -> None:
    if status != 0:
        raise OspoHeadError(f"{what} failed: {strerror(status)} (status {status})")


def workspace_bytes(rows: int, hidden: int, embed: int, vocab: int, num_seqs: int = 1) -> int:
    out = C.c_size_t(0)
    shape = Shape(rows, hidden, embed, vocab, num_seqs)
    check(load().ospo_head_workspace_bytes(C.byref(shape), C.byref(out)), "ospo_head_workspace_bytes")
    return int(out.value)


KERNEL_NAMES = ("gemm1_bias_gelu", "gemm2_logits_lse", "scalar_stage", "row_weights", "dact_gelu_bwd",
                "wgrad_w2", "colsum_db2_db1", "wgrad_w1", "dgrad_x", "gemm2_logits_plain", "decode_gemm1",
                "decode_gemm2", "cfg_merge_sample", "gen_img_embeds", "clip_adamw", "dp_exchange")


def profile_enable(on: bool) -> None:
    load().ospo_head_profile_enable(int(on))


def profile_read() -> dict:
    """{kernel name: (total_ms, spans)} since the previous read (synchronises on the recorded events)"""
    n = len(KERNEL_NAMES)
    ms = (C.c_float * n)()
    cnt = (C.c_int32 * n)()
    check(load().ospo_head_profile_read(ms, cnt, n), "ospo_head_profile_read")
    return {KERNEL_NAMES[i]: (float(ms[i]), int(cnt[i])) for i in range(n) if cnt[i] > 0}


def watchdog_record() -> list[int] | None:
    p = load().ospo_head_watchdog_record_host()
    if not p:
        return None
    return [int(p[i]) for i in range(6)]
