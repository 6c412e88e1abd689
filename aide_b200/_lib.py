"""ctypes binding of libaide_b200.so (the C ABI declared in include/aide_b200.h).

The product path has NO fallback: if the shared library is missing or a kernel call fails, an
exception is raised.  Build the library with ``python __graft_entry__.py`` (or ``make -C
aide_b200/csrc``).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libaide_b200.so")

FMT_F32, FMT_TF32X2, FMT_BF16, FMT_F16X2 = 0, 1, 2, 3

_vp, _i, _f, _d, _sz = C.c_void_p, C.c_int, C.c_float, C.c_double, C.c_size_t

# name -> (restype, argtypes); mirrors include/aide_b200.h one to one
SIGNATURES = {
    "aide_last_error": (C.c_char_p, []),
    "aide_version": (_i, []),
    "aide_launch_count": (C.c_ulonglong, []),
    "aide_has_tma": (_i, []),
    "aide_f16_saturated": (_i, [_i]),
    "aide_nchw_to_nhwc": (_i, [_i, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "aide_nhwc_to_nchw": (_i, [_i, _vp, _vp, _i, _i, _vp, _i, _i, _i, _i, _vp]),
    "aide_weight_prep": (_i, [_i, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "aide_weight_prep_batch": (_i, [_i, _i, C.POINTER(_vp), C.POINTER(_i), C.POINTER(_i), C.POINTER(_vp), C.POINTER(_vp),
                                    C.POINTER(_vp), C.POINTER(_vp), _vp]),
    "aide_conv3x3_stat_rows": (_i, [_i, _i, _i, _i, _i, _i]),
    "aide_conv3x3_fwd": (_i, [_i, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "aide_conv3x3_bn_relu_ok": (_i, [_i, _i, _i, _i, _i, _i]),
    "aide_conv3x3_bn_relu_fwd": (_i, [_i, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _i, _i, _i,
                                      _vp, _vp, _i, _i, _vp, _vp, _i, _i, _vp, _vp, _i, _i, _vp]),
    "aide_bn_eval_scale_shift_batch": (_i, [_i, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp),
                                            C.POINTER(_i), C.POINTER(_vp), _f, _vp]),
    "aide_conv3x3_plan_info": (_i, [_i, _i, _i, _i, _i, _i, C.POINTER(_i)]),
    "aide_conv3x3_wgrad_workspace_bytes": (_sz, [_i, _i, _i, _i, _i, _i]),
    "aide_conv3x3_dgrad": (_i, [_i, _vp, _vp, _i, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "aide_conv3x3_wgrad": (_i, [_i, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _sz, _vp, _vp]),
    "aide_conv3x3_wgrad_ex": (_i, [_i, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _sz, _vp, _vp, _i, _i, _vp, _vp]),
    "aide_bn_finalize": (_i, [_vp, _i, _i, _d, _vp, _vp, _vp, _vp, _f, _f, _i, _vp, _vp, _vp]),
    "aide_bn_finalize_grouped": (_i, [_vp, _i, _i, _i, _d, _vp, _vp, _vp, _vp, _f, _f, _i, _vp, _vp, _vp, _vp]),
    "aide_bn_ticket_slots": (_i, [_i]),
    "aide_bn_relu_apply_grouped": (_i, [_i, _vp, _i, _i, _i, _i, _i, _vp,
                                        _vp, _vp, _i, _i, _vp, _vp, _i, _i, _vp, _vp, _i, _i, _vp]),
    "aide_bn_relu_apply": (_i, [_i, _vp, _i, _i, _i, _i, _vp,
                                _vp, _vp, _i, _i, _vp, _vp, _i, _i, _vp, _vp, _i, _i, _vp]),
    "aide_bn_bwd_rows": (_i, [_i, _i, _i, _i]),
    "aide_bn_relu_bwd_reduce": (_i, [_vp, _vp, _vp, _i, _i, _i, _i,
                                     C.POINTER(_vp), C.POINTER(_i), C.POINTER(_i), _i,
                                     C.POINTER(_vp), C.POINTER(_i), C.POINTER(_i), _i, _vp, _vp, _vp, _vp]),
    "aide_bn_relu_bwd_apply": (_i, [_i, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i,
                                    _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "aide_sa_stat_rows": (_i, [_i, _i, _i]),
    "aide_sa_fwd": (_i, [_i, _vp, _vp, _i, _i, _i, _i, _i] + [_vp] * 13 + [_i, _i, _i, _vp]),
    "aide_sa_gate_apply": (_i, [_i, _vp, _vp, _i, _i, _vp, _vp, _i, _i, _i, _i, _i, _vp,
                                _vp, _vp, _i, _i, _vp, _vp, _i, _i, _vp, _vp, _i, _i, _vp]),
    "aide_sa_bwd_rows": (_i, [_i, _i, _i, _i]),
    "aide_sa_bwd_gate": (_i, [_i, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _i, _i, _i,
                              C.POINTER(_vp), C.POINTER(_i), C.POINTER(_i), _i,
                              C.POINTER(_vp), C.POINTER(_i), C.POINTER(_i), _i, _vp, _vp, _vp, _vp]),
    "aide_sa_bwd_workspace_floats": (_sz, [_i, _i, _i, _i, _i]),
    "aide_sa_bwd_chain": (_i, [_i, _vp, _vp, _i, _i, _i, _i, _i] + [_vp] * 12 + [_i, _i, _i, _i, _vp, _sz] + [_vp] * 10 + [_vp]),
    "aide_upsample2x_fwd": (_i, [_i, _vp, _vp, _i, _i, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "aide_upsample2x_bwd": (_i, [_vp, _i, _i, _vp, _i, _i, _i, _i, _vp]),
    "aide_zero_insert2x_fwd": (_i, [_i, _vp, _vp, _i, _i, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "aide_zero_insert2x_bwd": (_i, [_vp, _i, _i, _vp, _i, _i, _i, _i, _vp]),
    "aide_conv1x1_fwd": (_i, [_i, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "aide_conv1x1_bwd_rows": (_i, [_i, _i, _i, _i]),
    "aide_conv1x1_bwd": (_i, [_i, _vp, _vp, _i, _i, _i, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "aide_comm_alloc": (_i, [_sz, C.POINTER(_vp), C.c_char_p]),
    "aide_comm_open": (_i, [C.c_char_p, C.POINTER(_vp)]),
    "aide_comm_close": (_i, [_vp]),
    "aide_comm_free": (_i, [_vp]),
    "aide_comm_pad_words": (_i, [_i, _i]),
    "aide_allreduce_p2p": (_i, [C.POINTER(_vp), C.POINTER(_vp), _i, _i, _i, _i, _sz, _sz, _i, _vp]),
    "aide_loss_sums": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _f, _f, _i, _f, _vp, _vp, _vp]),
    "aide_loss_blocks": (_i, [_i, _i]),
    "aide_loss_image_finalize": (_i, [_vp, _i, _i, _i, _f, _f, _f, _vp, _vp, _vp, _vp]),
    "aide_loss_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _f, _f, _i, _f, _vp, _vp]),
    "aide_pixel_loss_fwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _f, _f, _i, _i, _vp, _vp]),
    "aide_pixel_loss_bwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _f, _f, _i, _i, _vp, _vp, _vp]),
    "aide_softmax_mse_fwd": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp]),
    "aide_softmax_mse_bwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp]),
    "aide_maxpool_nchw_fwd": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "aide_maxpool_nchw_bwd": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "aide_pseudo_label": (_i, [C.POINTER(_vp), _i, _i, _i, _i, _f, _vp, _vp, _vp]),
    "aide_argmax_mask": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "aide_reverse_aug": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "aide_forward_aug": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "aide_coteach_select": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _f, _f, _f, _f, _f, _vp, _vp, _vp, _vp, _vp, _vp]),
    "aide_adam_amsgrad": (_i, [_vp, _vp, _vp, _vp, _vp, _sz, _f, _f, _f, _f, _i, _f, _vp]),
    "aide_coteach_select_ex": (_i, [_vp, _i, _i, _vp, _vp, _i, _i, _i, _i, _f, _vp, _f, _f, _f, _f, _vp, _vp, _vp, _vp,
                                    _vp, _vp]),
    "aide_adam_amsgrad_dev": (_i, [_vp, _vp, _vp, _vp, _vp, _sz, _f, _f, _f, _f, _vp, _vp, _f, _vp, _vp]),
}


class AideError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: the aide_b200 CUDA library has not been built. "
            "Run `python __graft_entry__.py` (or `make -C aide_b200/csrc`). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


def check(rc: int) -> None:
    if rc != 0:
        raise AideError(lib.aide_last_error().decode("utf-8", "replace"))


def call(name: str, *args) -> None:
    check(getattr(lib, name)(*args))
