"""ctypes binding of the C-ABI library (``include/pnmn.h``).

The library is built in-tree (``probnmn_clevr_b200/libpnmn.so``) by ``__graft_entry__.build()`` or
``make -C probnmn_clevr_b200/csrc``.  There is no fallback: if the library is missing, or a compute
entry point is called without a CUDA device, this module raises.
"""
import ctypes
import os
import subprocess
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_int64, c_uint8, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PNMN_LIB") or os.path.join(_HERE, "libpnmn.so")   # (PNMN_LIB: experiments with build variants)
CSRC = os.path.join(_HERE, "csrc")

NSMAX = 2
MAX_MODULE_PARAMS = 12

# pnmn_token_kind
TOK_SKIP, TOK_SCENE, TOK_AND, TOK_OR, TOK_COMPARE, TOK_QUERY, TOK_RELATE, TOK_SAME, TOK_ATTENTION = range(9)

# pnmn_size_slot
SZ_ARENA16, SZ_ARENA18, SZ_ARENA22, SZ_MAPS, SZ_DMAPS, SZ_IDX, SZ_BLOB, SZ_AIN = range(8)
SZ_COUNT = 8

# ConvFlags (csrc/executor.h)
F_BIAS, F_RELU, F_STORE, F_DOTSIG, F_MASK, F_ACCUM, F_HALF = 1, 2, 4, 8, 16, 32, 64


class Buffers(Structure):
    _fields_ = [
        ("arena16", c_void_p), ("arena18", c_void_p), ("arena22", c_void_p),
        ("maps", c_void_p), ("dmaps", c_void_p), ("idx", c_void_p),
        ("blob", c_void_p), ("packed", c_void_p), ("params", c_void_p), ("grads", c_void_p),
        ("ain", c_void_p), ("scratch", c_void_p),
    ]


# ---- mirrors of the internal task records, used only by the kernel-level bring-up tests ----------
class ConvCfg(Structure):
    _fields_ = [(n, c_int) for n in (
        "n_kb", "kb_per_in", "ntaps", "dil", "S_in", "P_in", "S_out", "P_out", "S_aux", "P_aux", "flags", "lead")]


class ConvTask(Structure):
    _fields_ = [
        ("in_", (c_void_p * NSMAX) * 2), ("out", c_void_p * NSMAX), ("aux", c_void_p * NSMAX),
        ("map_out", c_void_p * NSMAX), ("w", c_void_p), ("bias", c_void_p), ("w3", c_void_p), ("b3", c_void_p),
        ("cfg", c_int), ("n_samp", c_int), ("mt0", c_int), ("n_mt", c_int),
    ]


class WgradInst(Structure):
    _fields_ = [("dz", c_void_p), ("x", c_void_p)]


class WgradTask(Structure):
    _fields_ = [
        ("inst", c_void_p), ("n_inst", c_int), ("tap_row", c_int), ("ntaps_x", c_int), ("dil", c_int),
        ("S", c_int), ("P", c_int), ("cin_total", c_int), ("cin0", c_int), ("ksize", c_int),
        ("dw", c_void_p), ("scale", c_void_p),
    ]


class PackTask(Structure):
    _fields_ = [
        ("src_off", c_int64), ("dst_off", c_int64), ("first_tile", c_int), ("n_kb", c_int), ("ntaps", c_int),
        ("flip", c_int), ("k_off", c_int), ("n_off", c_int), ("k_stride", c_int), ("n_stride", c_int),
        ("tap_stride", c_int), ("pad_", c_int),
    ]


class EltTask(Structure):
    _fields_ = [
        ("op", c_int), ("flags", c_int), ("a", c_void_p), ("b", c_void_p), ("c", c_void_p), ("g", c_void_p),
        ("o", c_void_p), ("o2", c_void_p), ("w", c_void_p), ("dw", c_void_p), ("dw2", c_void_p), ("idx", c_void_p),
        ("scale", c_void_p), ("part", c_int), ("n_parts", c_int), ("pad_", c_int64 * 3),
    ]


assert ctypes.sizeof(ConvTask) == 128 and ctypes.sizeof(EltTask) == 128


class PgDesc(Structure):
    """pnmn_pg_desc (include/pnmn.h): sizes + float offsets of the seq2seq parameters inside the flat buffer."""
    _fields_ = [
        ("vocab_src", c_int32), ("vocab_tgt", c_int32), ("hidden", c_int32), ("num_layers", c_int32),
        ("src_embed", c_int64),
        ("enc_w_ih", c_int64 * 2), ("enc_w_hh", c_int64 * 2), ("enc_b_ih", c_int64 * 2), ("enc_b_hh", c_int64 * 2),
        ("tgt_embed", c_int64),
        ("dec_w_ih", c_int64), ("dec_w_hh", c_int64), ("dec_b_ih", c_int64), ("dec_b_hh", c_int64),
        ("out_w", c_int64), ("out_b", c_int64),
    ]

class PriorDesc(Structure):
    """pnmn_prior_desc (include/pnmn.h)."""
    _fields_ = [
        ("vocab", c_int32), ("hidden", c_int32), ("num_layers", c_int32), ("pad_", c_int32),
        ("embed", c_int64),
        ("w_ih", c_int64 * 2), ("w_hh", c_int64 * 2), ("b_ih", c_int64 * 2), ("b_hh", c_int64 * 2),
        ("proj", c_int64),
    ]


PLAN_INPUT_BY_ROW = 1   # PNMN_PLAN_INPUT_BY_ROW
PLAN_FORWARD_HALF = 2   # PNMN_PLAN_FORWARD_HALF

EXPORTS = [
    "pnmn_version", "pnmn_last_error", "pnmn_model_create", "pnmn_model_destroy", "pnmn_model_packed_floats",
    "pnmn_plan_create", "pnmn_plan_destroy", "pnmn_plan_upload", "pnmn_plan_valid", "pnmn_plan_sizes", "pnmn_plan_stats",
    "pnmn_nmn_forward", "pnmn_nmn_backward", "pnmn_debug_launch_conv", "pnmn_debug_launch_wgrad",
    "pnmn_debug_pack", "pnmn_debug_nchw_to_planes", "pnmn_debug_launch_elt", "pnmn_profile_enable",
    "pnmn_profile_read", "pnmn_debug_set_trace", "pnmn_debug_host_times", "pnmn_debug_plan_meta", "pnmn_debug_plan_records", "pnmn_debug_plan_maps", "pnmn_debug_graph_stats",
    "pnmn_relu_pool_fwd", "pnmn_relu_pool_bwd", "pnmn_relu_pool_fwd_bias", "pnmn_launch_count", "pnmn_pg_workspace_bytes", "pnmn_pg_forward", "pnmn_pg_backward", "pnmn_pg_debug_layout", "pnmn_pg_forward_mixed",
    "pnmn_prior_workspace_bytes", "pnmn_prior_forward", "pnmn_clamp_adam", "pnmn_elbo_glue", "pnmn_set_reserved_sms", "pnmn_has_bringup_kernels", "pnmn_answer_loss_forward", "pnmn_answer_loss_backward", "pnmn_gemm_split", "pnmn_gemm_split_workspace", "pnmn_nmn_forward_f16", "pnmn_round_features_f16", "pnmn_plan_create_ex", "pnmn_plan_set_exec_ctas", "pnmn_model_pack_table_bytes", "pnmn_model_pack_table", "pnmn_model_ain_floats", "pnmn_nmn_prestage",
]


def build(verbose: bool = False) -> str:
    """Compile the sm_100a shared library in-tree (nvcc cross-compiles without a GPU)."""
    proc = subprocess.run(["make", "-C", CSRC, "-j8"], capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("building libpnmn.so failed:\n" + proc.stdout[-4000:] + proc.stderr[-4000:])
    if verbose:
        print(proc.stdout[-2000:])
    return LIB_PATH


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU / eager fallback for the NMN executor)")
    L = ctypes.CDLL(LIB_PATH)
    L.pnmn_version.restype = c_int
    L.pnmn_last_error.restype = c_char_p
    L.pnmn_model_create.restype = c_void_p
    L.pnmn_model_create.argtypes = [c_int, POINTER(c_int32), POINTER(c_int64), POINTER(c_int64), c_int]
    L.pnmn_model_destroy.argtypes = [c_void_p]
    L.pnmn_model_packed_floats.restype = c_int64
    L.pnmn_model_packed_floats.argtypes = [c_void_p]
    L.pnmn_plan_create.restype = c_void_p
    L.pnmn_plan_create.argtypes = [c_void_p, POINTER(c_int64), c_int, c_int, c_int]
    L.pnmn_plan_destroy.argtypes = [c_void_p]
    L.pnmn_plan_upload.argtypes = [c_void_p, c_void_p, c_void_p]
    L.pnmn_plan_valid.argtypes = [c_void_p, POINTER(c_uint8)]
    L.pnmn_plan_sizes.argtypes = [c_void_p, POINTER(c_int64)]
    L.pnmn_plan_stats.argtypes = [c_void_p, POINTER(c_int64)]
    L.pnmn_nmn_forward.argtypes = [c_void_p, POINTER(Buffers), c_void_p, c_void_p, c_void_p]
    L.pnmn_plan_create_ex.restype = c_void_p
    L.pnmn_plan_create_ex.argtypes = [c_void_p, POINTER(c_int64), c_int, c_int, c_int, c_int]
    L.pnmn_plan_set_exec_ctas.argtypes = [c_void_p, c_int]
    L.pnmn_model_pack_table_bytes.restype = c_int64
    L.pnmn_model_pack_table_bytes.argtypes = [c_void_p]
    L.pnmn_model_pack_table.argtypes = [c_void_p, c_void_p]
    L.pnmn_model_ain_floats.restype = c_int64
    L.pnmn_model_ain_floats.argtypes = [c_void_p, c_int]
    L.pnmn_nmn_prestage.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p]
    L.pnmn_nmn_forward_f16.argtypes = [c_void_p, POINTER(Buffers), c_void_p, c_void_p, c_void_p]
    L.pnmn_round_features_f16.argtypes = [c_void_p, c_void_p, c_int64, c_void_p]
    L.pnmn_nmn_backward.argtypes = [c_void_p, POINTER(Buffers), c_void_p, c_void_p]
    L.pnmn_debug_launch_conv.argtypes = [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_void_p]
    L.pnmn_debug_launch_wgrad.argtypes = [c_void_p, c_int, c_void_p, c_int, c_int, c_void_p]
    L.pnmn_debug_pack.argtypes = [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]
    L.pnmn_debug_nchw_to_planes.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int64, c_void_p]
    L.pnmn_debug_launch_elt.argtypes = [c_void_p, c_int, c_void_p]
    L.pnmn_debug_set_trace.argtypes = [c_void_p, c_int64]
    L.pnmn_debug_host_times.argtypes = [POINTER(ctypes.c_double)]
    L.pnmn_debug_plan_meta.restype = c_int64
    L.pnmn_debug_plan_meta.argtypes = [c_void_p, c_int, c_void_p, c_int64]
    L.pnmn_debug_plan_records.restype = c_int64
    L.pnmn_debug_plan_records.argtypes = [c_void_p, c_int, c_void_p, c_int64]
    L.pnmn_debug_plan_maps.restype = c_int64
    L.pnmn_debug_plan_maps.argtypes = [c_void_p, c_void_p, c_int64]
    L.pnmn_debug_graph_stats.argtypes = [POINTER(c_int64)]
    L.pnmn_pg_workspace_bytes.restype = c_int64
    L.pnmn_pg_workspace_bytes.argtypes = [POINTER(PgDesc), c_int, c_int, c_int, c_int, c_int]
    L.pnmn_pg_forward.argtypes = [POINTER(PgDesc), c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                  ctypes.c_uint64, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    L.pnmn_pg_forward_mixed.argtypes = [POINTER(PgDesc), c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                        ctypes.c_uint64, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    L.pnmn_pg_backward.argtypes = [POINTER(PgDesc), c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                   c_void_p, c_void_p]
    L.pnmn_pg_debug_layout.argtypes = [POINTER(PgDesc), c_int, c_int, c_int, c_int, c_int, POINTER(c_int64)]
    L.pnmn_prior_workspace_bytes.restype = c_int64
    L.pnmn_prior_workspace_bytes.argtypes = [POINTER(PriorDesc), c_int, c_int]
    L.pnmn_prior_forward.argtypes = [POINTER(PriorDesc), c_void_p, c_void_p, c_int, c_int, ctypes.c_uint64, c_void_p,
                                     c_void_p, c_void_p, c_void_p, c_void_p]
    c_double = ctypes.c_double
    L.pnmn_clamp_adam.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_double, c_double, c_double,
                                  c_double, c_double, c_double, c_int, c_void_p]
    L.pnmn_elbo_glue.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_float, c_float, c_float, c_int,
                                 c_void_p, c_void_p, c_void_p, c_void_p]
    L.pnmn_relu_pool_fwd.argtypes = [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p]
    L.pnmn_relu_pool_bwd.argtypes = [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p]
    L.pnmn_relu_pool_fwd_bias.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p]
    L.pnmn_launch_count.restype = ctypes.c_longlong
    L.pnmn_launch_count.argtypes = [c_int]
    L.pnmn_profile_enable.argtypes = [c_int]
    L.pnmn_set_reserved_sms.argtypes = [c_int]
    L.pnmn_gemm_split.argtypes = [c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int, c_int, c_int,
                                  c_void_p, c_int, c_void_p, c_int64, c_void_p]
    L.pnmn_gemm_split_workspace.restype = c_int64
    L.pnmn_gemm_split_workspace.argtypes = [c_int, c_int, c_int]
    L.pnmn_answer_loss_forward.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                                           c_void_p]
    L.pnmn_answer_loss_backward.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]
    L.pnmn_profile_read.argtypes = [POINTER(ctypes.c_double), POINTER(c_int64)]
    _lib = L
    return L


def check(status: int, what: str = "pnmn call") -> None:
    if status != 0:
        raise RuntimeError(f"{what} failed: {lib().pnmn_last_error().decode()}")
