"""ctypes binding of libqasr.so (include/qasr.h).

The library is built in-tree by q-asr_b200/csrc/build.sh (see __graft_entry__.build()).
There is NO CPU fallback: if the shared library is missing or no CUDA device is
present, every op raises.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("QASR_LIB_PATH", os.path.join(_HERE, "libqasr.so"))  # override: kernel experiments

EPI_REQUANT_I8, EPI_DEQUANT_F32, EPI_RAW_I32, EPI_STATS = 0, 1, 2, 3
F_RELU0, F_RELU1, F_STAGE2, F_MASK = 1, 2, 4, 8
F_EXACT_F32, F_EXACT_DYAD, F_B_SIGNED, F_B1_SIGNED, F_BIAS64 = 16, 32, 64, 128, 256
F_FAST_F32 = 512
F_DEVICE_TABLES = 1024
F_RANGE_CHECK = 2048
F_CLAMP16 = 4096
F_STAGE2_16 = 8192
DYN_MAX_CAND = 4


class QasrError(RuntimeError):
    pass


class Epilogue(C.Structure):
    _fields_ = [
        ("mode", C.c_int32), ("flags", C.c_int32),
        ("ch0", C.c_void_p), ("ch1", C.c_void_p),
        ("prev", C.c_void_p), ("prev_ld", C.c_int32),
        ("qmin", C.c_int32), ("qmax", C.c_int32),
        ("m2", C.c_int32), ("e2", C.c_int32),
        ("qmin2", C.c_int32), ("qmax2", C.c_int32),
        ("out_lens", C.c_void_p),
        ("f32_mult", C.c_void_p),
    ]


class GemmOperand(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("wp", C.c_void_p),
        ("x_ld", C.c_int32), ("C_in", C.c_int32), ("T_in", C.c_int32), ("k", C.c_int32),
        ("stride", C.c_int32), ("dilation", C.c_int32), ("pad", C.c_int32),
    ]


class DynOperand(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("w", C.c_void_p), ("s_x", C.c_void_p), ("ch", C.c_void_p), ("wabs", C.c_void_p),
        ("x_ld", C.c_int32), ("C_in", C.c_int32), ("T_in", C.c_int32), ("k", C.c_int32), ("stride", C.c_int32),
        ("dilation", C.c_int32), ("pad", C.c_int32), ("depthwise", C.c_int32), ("in_unsigned", C.c_int32),
    ]


class DynFinalizeArgs(C.Structure):
    _fields_ = [
        ("src_kind", C.c_int32), ("src", C.c_void_p), ("src_scale", C.c_void_p), ("n_levels", C.c_int32),
        ("act_scale", C.c_void_p), ("n_prod", C.c_int32), ("prod", C.c_void_p * 2), ("prod_C", C.c_int32 * 2),
        ("prev_scale", C.c_void_p), ("prev_me", C.c_void_p), ("n_cons", C.c_int32), ("cons", C.c_void_p * 2),
        ("cons_sw", C.c_void_p * 2), ("cons_bias", C.c_void_p * 2), ("cons_C", C.c_int32 * 2),
        ("peer_slots", C.c_void_p), ("rank", C.c_int32), ("world", C.c_int32), ("slot", C.c_int32),
        ("seq", C.c_void_p),
        ("prod_wabs", C.c_void_p * 2), ("prod_qabs", C.c_int32 * 2),
        ("stats", C.c_void_p), ("stats_C", C.c_int32), ("stats_dual", C.c_int32), ("phase", C.c_int32),
    ]


_lib = None

_i32, _i64, _vp, _f32 = C.c_int32, C.c_int64, C.c_void_p, C.c_float
_SIGS = {
    "qasr_version": (C.c_int, []),
    "qasr_last_error": (C.c_char_p, []),
    "qasr_sm_count": (C.c_int, []),
    "qasr_dwconv1d_i8": (C.c_int, [_vp, _i32, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32,
                                   _i32, C.POINTER(Epilogue), _vp]),
    "qasr_dw_packed_bytes": (_i64, [_i32, _i32, _i32, _i32, _i32]),
    "qasr_pack_dw_weights": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp]),
    "qasr_conv1d_gemm_i8": (C.c_int, [C.POINTER(GemmOperand), C.POINTER(GemmOperand), _vp, _i32, _i32,
                                      _i32, _i32, C.POINTER(Epilogue), _vp]),
    "qasr_conv1d_ntc_i8": (C.c_int, [C.POINTER(GemmOperand), C.POINTER(GemmOperand), _vp, _i32, _i32,
                                     _i32, _i32, C.POINTER(Epilogue), _vp]),
    "qasr_quantize_f32_ntc": (C.c_int, [_vp, _i32, _vp, _i32, _i32, _vp, _i32, _i32, _i32, _i32, _vp, _vp]),
    "qasr_packed_weight_bytes": (_i64, [_i32, _i32, _i32]),
    "qasr_pack_conv_weights": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _vp]),
    "qasr_quantize_f32": (C.c_int, [_vp, _i32, _vp, _i32, _vp, _i32, _vp, _i32, _i32, _i32, _vp, _vp]),
    "qasr_requant_f32": (C.c_int, [_vp, _vp, _i32, _vp, _vp, _i32, _vp, _i32, _i32, _vp, _i32, _vp, _i32,
                                   _i32, _i32, _i32, _vp, _vp]),
    "qasr_dyadic_from_scales": (C.c_int, [_vp, _i32, _vp, _vp, _vp, _vp]),
    "qasr_codes_from_f32": (C.c_int, [_vp, _i32, _vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "qasr_noise_codes_i8": (C.c_int, [_vp, _i32, _vp, _vp, C.c_int64, _vp]),
    "qasr_combine_noise_f32": (C.c_int, [_vp, _vp, _i32, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "qasr_requant_i8": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _i32, _i32, _i32, _i32,
                                  _vp, _vp]),
    "qasr_minmax_workspace_bytes": (_i64, []),
    "qasr_minmax_f32": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp]),
    "qasr_quantile_workspace_bytes": (_i64, []),
    "qasr_quantile2_f32": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _vp, _f32, _f32, _vp, _vp, _vp]),
    "qasr_ctc_argmax": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _vp, _vp]),
    "qasr_log_softmax_tv": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _vp, _vp]),
    "qasr_melspec_f32": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _i32, _f32, _f32, _vp,
                                   _i32, _i32, _vp]),
    "qasr_normalize_features_f32": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _i32, _vp, _i32, _vp]),
    "qasr_ctc_collapse": (C.c_int, [_vp, _i32, _i32, _vp, _i32, _vp, _vp, _vp]),
    "qasr_dyn_refine": (C.c_int, [_vp, C.POINTER(DynOperand), C.POINTER(DynOperand), _i32, _vp, _i32, _i32, _i32, _vp,
                                  _vp, _vp]),
    "qasr_dyn_finalize": (C.c_int, [C.POINTER(DynFinalizeArgs), _vp]),
    "qasr_minmax_i8": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp]),
    "qasr_requant_i8_dev": (C.c_int, [_vp, _i32, _i32, _vp, _i32, _i32, _vp, _i32, _i32, _i32, _i32, _vp, _vp]),
    "qasr_comm_alloc": (C.c_int, [_i64, C.POINTER(C.c_void_p), _vp]),
    "qasr_comm_open": (C.c_int, [_vp, C.POINTER(C.c_void_p)]),
    "qasr_comm_close": (C.c_int, [_vp]),
    "qasr_comm_free": (C.c_int, [_vp]),
    "qasr_debug_set": (C.c_int, [C.c_int, C.c_longlong]),
    "qasr_debug_knobs_set": (C.c_int, []),
    "qasr_profile_begin": (C.c_int, [_vp, C.c_int]),
    "qasr_profile_end": (C.c_int, []),
}

# bring-up hooks are not part of include/qasr.h
EXPORTED_SYMBOLS = [s for s in _SIGS if s not in ("qasr_debug_set", "qasr_debug_knobs_set", "qasr_profile_begin",
                                                 "qasr_profile_end")]


def load_library():
    """dlopen libqasr.so and bind every symbol of include/qasr.h (no GPU needed)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise QasrError("libqasr.so not built (%s); run `python -c 'import __graft_entry__ as g; "
                            "g.build()'` or q-asr_b200/csrc/build.sh" % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def lib():
    """The library for compute calls: additionally requires a CUDA device."""
    if not torch.cuda.is_available():
        raise QasrError("qasr kernels need a CUDA device (sm_100a); there is no CPU fallback")
    return load_library()


def check(rc):
    if rc != 0:
        msg = load_library().qasr_last_error().decode("utf-8", "replace")
        if rc == -1:
            raise ValueError("qasr: " + msg)
        if rc == -3:
            raise NotImplementedError("qasr: " + msg)
        raise QasrError("qasr error %d: %s" % (rc, msg))


def ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
