"""ctypes binding of libpmce_b200.so (the C ABI declared in include/pmce_b200.h).

There is NO fallback: if the shared library is missing or an entry point is absent, importing the
library handle raises. Build it with `python -m pmce_b200.build` (or `__graft_entry__.build()`).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# PMCE_B200_PROFILING=1 selects the -DPMCE_PROFILING build (timing knobs that corrupt results; see build.py)
LIB_PATH = os.path.join(_HERE, "libpmce_b200_prof.so" if os.environ.get("PMCE_B200_PROFILING", "0") == "1" else "libpmce_b200.so")


class PmceDims(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "num_joint", "embed_dim", "depth", "seqlen", "num_vert_ds", "num_vert", "feat_dim", "gru_hidden",
        "coevo_dim", "lifter_heads")]

    def key(self):
        return tuple(getattr(self, n) for n, _ in self._fields_)


class PmceSpinConv(C.Structure):
    _fields_ = [("w_off", C.c_int64), ("b_off", C.c_int64), ("cout", C.c_int32), ("k", C.c_int32)]


class PmceSlot(C.Structure):
    _fields_ = [("offset", C.c_uint64), ("rows", C.c_int64), ("cols", C.c_int64), ("ld", C.c_int64)]


_P = C.c_void_p
_DP = C.POINTER(PmceDims)

# name -> (restype, argtypes); must list every symbol include/pmce_b200.h declares
SIGNATURES = {
    "pmce_last_error": (C.c_char_p, []),
    "pmce_abi_version": (C.c_int, []),
    "pmce_weights_bytes": (C.c_size_t, [_DP]),
    "pmce_weight_slot": (C.c_int, [_DP, C.c_char_p, C.POINTER(PmceSlot)]),
    "pmce_pack_weights": (C.c_int, [_DP, _P, _P]),
    "pmce_workspace_bytes": (C.c_size_t, [_DP, C.c_int]),
    "pmce_forward": (C.c_int, [_DP, _P, _P, _P, _P, C.c_int, _P, _P, _P, _P, C.c_size_t, _P]),
    "pmce_sliding_windows": (C.c_int, [_DP, C.c_int, C.c_int]),
    "pmce_forward_sliding": (C.c_int, [_DP, _P, _P, _P, _P, C.c_int, C.c_int, _P, _P, _P, _P, C.c_size_t, _P]),
    "pmce_io_bytes": (C.c_size_t, [_DP, C.c_int]),
    "pmce_forward_host": (C.c_int, [_DP, _P, _P, _P, _P, C.c_int, _P, _P, _P, _P, _P, C.c_size_t, _P]),
    "pmce_lifter_forward": (C.c_int, [_DP, _P, _P, _P, C.c_int, _P, _P, C.c_size_t, _P]),
    "pmce_gru_mid": (C.c_int, [_DP, _P, _P, C.c_int, _P, _P, C.c_size_t, _P]),
    "pmce_adaln_slots": (C.c_int, []),
    "pmce_adaln_gammabeta": (C.c_int, [_DP, _P, _P, C.c_int, _P, _P, C.c_size_t, _P]),
    "pmce_coevo_block": (C.c_int, [_DP, _P, C.c_int, _P, _P, _P, C.c_int, _P, _P, _P, C.c_size_t, _P]),
    "pmce_cross_attn_block": (C.c_int, [_DP, _P, C.c_int, C.c_int, _P, _P, _P, _P, C.c_int, _P, _P, C.c_size_t, _P]),
    "pmce_ca_fold_bytes": (C.c_size_t, [C.c_int]),
    "pmce_ca_vertex_fused": (C.c_int, [_DP, _P, C.c_int, _P, _P, _P, _P, C.c_int, _P, _P, _P, C.c_int, _P]),
    "pmce_ca_vertex_fused_embed": (C.c_int, [_DP, _P, C.c_int, _P, _P, _P, _P, _P, C.c_int, _P, C.c_int, _P, _P]),
    "pmce_self_attn_block": (C.c_int, [_DP, _P, C.c_int, C.c_int, _P, _P, C.c_int, _P, _P, C.c_size_t, _P]),
    "pmce_mesh_epilogue": (C.c_int, [_DP, _P, _P, _P, C.c_int, _P, _P, C.c_size_t, _P]),
    "pmce_decoder_forward": (C.c_int, [_DP, _P, _P, _P, _P, C.c_int, _P, _P, _P, _P, C.c_size_t, _P]),
    "pmce_jregress": (C.c_int, [_P, _P, _P, C.c_int, _P, C.c_int, C.c_int, C.c_float, _P, _P]),
    "pmce_eval_errors": (C.c_int, [_P, _P, _P, C.c_int, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_float, _P, _P, _P, _P]),
    "pmce_linear": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P]),
    "pmce_linear_tc_scratch_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "pmce_linear_tc": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, C.c_size_t, _P]),
    "pmce_split_bf16": (C.c_int, [_P, C.c_int, C.c_int, _P, _P, _P]),
    "pmce_linear_tc_presplit": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P, _P, _P]),
    "pmce_launch_count": (C.c_ulonglong, []),
    "pmce_spin_num_convs": (C.c_int, []),
    "pmce_spin_workspace_bytes": (C.c_size_t, [C.c_int]),
    "pmce_spin_features": (C.c_int, [_P, _P, _P, _P, C.c_int, _P, C.c_int, _P, _P, C.c_size_t, _P]),
    "smpl_blend_ld": (C.c_int, []),
    "smpl_workspace_bytes": (C.c_size_t, [C.c_int]),
    "smpl_lbs_forward_scaled": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, C.c_int, C.c_float, _P, _P, _P, C.c_size_t, _P]),
    "smpl_lbs_forward_sparse": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, C.c_int, C.c_float, _P, _P, _P, C.c_size_t, _P]),
    "smpl_lbs_forward": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, C.c_int, _P, _P, _P, C.c_size_t, _P]),
}

_lib = None


class PmceError(RuntimeError):
    pass


def load():
    """Load libpmce_b200.so and bind every entry point. Raises if anything is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PmceError(
            f"{LIB_PATH} not found: the CUDA library is required (there is no CPU/PyTorch fallback). "
            "Build it with `python -m pmce_b200.build`.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise PmceError(f"libpmce_b200.so does not export `{name}`") from e
        fn.restype = res
        fn.argtypes = args
    if lib.pmce_abi_version() != 1:
        raise PmceError("libpmce_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().pmce_last_error()
        raise PmceError(f"{what} failed (code {rc}): {msg.decode() if msg else '?'}")
