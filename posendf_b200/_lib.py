"""ctypes binding of libpndf.so (include/pndf.h).  No fallback: if the CUDA library is missing or
fails, every entry point raises."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PNDF_LIBRARY") or os.path.join(_HERE, "libpndf.so")      # PNDF_LIBRARY: A/B builds (tools/)

ACT = {"relu": 0, "lrelu": 1, "softplus": 2}
MAX_HIDDEN = 8


class PndfConfig(C.Structure):
    _fields_ = [
        ("use_enc", C.c_int32), ("enc_act", C.c_int32), ("enc_beta", C.c_float),
        ("df_act", C.c_int32), ("df_beta", C.c_float), ("in_dim", C.c_int32),
        ("num_hidden", C.c_int32), ("dims", C.c_int32 * MAX_HIDDEN), ("device", C.c_int32),
    ]


# every symbol include/pndf.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "pndf_create": (C.c_int, [C.POINTER(PndfConfig), C.POINTER(C.c_void_p)]),
    "pndf_destroy": (C.c_int, [C.c_void_p]),
    "pndf_param_count": (C.c_int, [C.POINTER(PndfConfig), C.POINTER(C.c_size_t)]),
    "pndf_set_weights": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "pndf_set_weights_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "pndf_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]),
    "pndf_forward_grad": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pndf_project": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "pndf_project_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int]),
    "pndf_project_gather": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "pndf_peer_alloc": (C.c_int, [C.c_int, C.c_size_t, C.POINTER(C.c_void_p), C.c_void_p]),
    "pndf_peer_open": (C.c_int, [C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]),
    "pndf_peer_close": (C.c_int, [C.c_int, C.c_void_p]),
    "pndf_peer_free": (C.c_int, [C.c_int, C.c_void_p]),
    "pndf_peer_barrier": (C.c_int, [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_uint32, C.c_void_p]),
    "pndf_prior_grad": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pndf_denoise_prior": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pndf_debug_dump_floats": (C.c_int, [C.POINTER(C.c_size_t)]),
    "pndf_forward_grad_debug": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pndf_softplus_adjoint": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_float,
                                        C.c_int64, C.c_int, C.c_void_p, C.c_void_p]),
    "pndf_act_handoff_bytes": (C.c_int, [C.c_void_p, C.c_int64, C.POINTER(C.c_size_t)]),
    "pndf_forward_grad_export": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pndf_forward_tangent_export": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pndf_encoder_tangent": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]),
    "pndf_encoder_param_grads": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pndf_train_losses": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pndf_wgrad_accumulate": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float,
                                        C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "pndf_adam_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_double, C.c_double,
                                 C.c_double, C.c_double, C.c_double, C.c_double, C.c_int64, C.c_void_p]),
    "pndf_feed_batch": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                  C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_void_p]),
    "pndf_axis_angle_to_quaternion": (C.c_int, [C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "pndf_quaternion_to_axis_angle": (C.c_int, [C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "pndf_knn_rerank": (C.c_int, [C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pndf_knn_exact": (C.c_int, [C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pndf_set_tile_policy": (C.c_int, [C.c_void_p, C.c_int]),
    "pndf_tile_for_batch": (C.c_int, [C.c_void_p, C.c_int64, C.POINTER(C.c_int)]),
    "pndf_fp32_peak": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_double)]),
    "pndf_launch_count": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64)]),
    "pndf_num_sms": (C.c_int, [C.c_void_p, C.POINTER(C.c_int)]),
    "pndf_last_error": (C.c_char_p, []),
    "pndf_version": (C.c_char_p, []),
}

_lib = None


def load():
    """dlopen libpndf.so (built in-tree by __graft_entry__.build()).  Raises if it is not there."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(there is no CPU / PyTorch fallback for the PoseNDF hot path)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int):
    if rc != 0:
        raise RuntimeError("libpndf: " + load().pndf_last_error().decode())


def make_config(use_enc=True, enc_act="lrelu", enc_beta=100.0, df_act="lrelu", df_beta=100.0, in_dim=None,
                dims=(256, 512, 1024, 512, 256, 64), device=0) -> PndfConfig:
    cfg = PndfConfig()
    cfg.use_enc = int(bool(use_enc))
    cfg.enc_act = ACT[enc_act]
    cfg.enc_beta = float(enc_beta)
    cfg.df_act = ACT[df_act]
    cfg.df_beta = float(df_beta)
    cfg.in_dim = int(in_dim if in_dim is not None else (126 if use_enc else 84))
    if len(dims) > MAX_HIDDEN:
        raise RuntimeError("libpndf: too many hidden layers")
    cfg.num_hidden = len(dims)
    for i, d in enumerate(dims):
        cfg.dims[i] = int(d)
    cfg.device = int(device)
    return cfg
