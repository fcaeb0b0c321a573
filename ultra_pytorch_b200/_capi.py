"""ctypes binding of include/ultra_b200.h (the C-ABI shared library with the sm_100a kernels).

There is NO fallback: if the library cannot be loaded (or built from the in-tree sources with nvcc) importing
this module raises, and every product entry point fails loudly.
"""
import ctypes
import os

from . import build as _build

_c_float_p = ctypes.c_void_p   # raw device pointers (tensor.data_ptr())
_vp = ctypes.c_void_p
_sz = ctypes.c_size_t
_i = ctypes.c_int
_f = ctypes.c_float
_ip = ctypes.POINTER(ctypes.c_int)

# name -> (restype, argtypes); mirrors include/ultra_b200.h declaration by declaration
SIGNATURES = {
    "ub200_last_error": (ctypes.c_char_p, []),
    "ub200_abi_version": (_i, []),
    "ub200_launch_count": (ctypes.c_ulonglong, []),
    "ub200_feed_bytes": (_sz, [_i, _i, _i, _i]),
    "ub200_pack_feed_host": (_i, [_vp, _i, _i, _vp, _vp, _i, _i, _vp, _sz, _i]),
    "ub200_pack_ids_host": (_i, [_vp, _vp, _i, _i, _i, _vp, _sz]),
    "ub200_convert_f64_f32_host": (_i, [_vp, _vp, _sz, _i]),
    "ub200_stage_feed": (_i, [_vp, _i, _i, _vp, _vp, _i, _i, _vp, _sz, _vp, _i, _i, _vp]),
    "ub200_stage_feed_pipelined": (_i, [_vp, _i, _i, _vp, _vp, _i, _i, _vp, _sz, _vp, _i, _i, _vp, _vp, _vp, _vp]),
    "ub200_stage_ids_pipelined": (_i, [_vp, _vp, _i, _i, _i, _vp, _sz, _vp, _vp, _vp, _vp, _vp]),
    "ub200_event_record": (_i, [_vp, _vp]),
    "ub200_stage_timeline": (_i, [_vp]),
    "ub200_rank_metrics": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _ip, _i, ctypes.c_float, _vp, _vp, _vp]),
    "ub200_regression_em": (_i, [_vp, _vp, _i, _i, _vp, _vp, ctypes.c_ulonglong, ctypes.c_ulonglong, _vp, _vp, _vp, _sz, _vp]),
    "ub200_regem_update": (_i, [_vp, _vp, _i, ctypes.c_float, _vp]),
    "ub200_set_tc_mode": (_i, [_i]),
    "ub200_mlp_param_count": (_sz, [_i, _ip, _i]),
    "ub200_mlp_workspace_bytes": (_sz, [_i, _i, _i, _ip, _i, _i]),
    "ub200_mlp_forward": (_i, [_vp, _vp, _i, _i, _i, _ip, _i, _vp, _vp, _vp, _sz, _i, _vp]),
    "ub200_mlp_backward": (_i, [_vp, _vp, _i, _i, _i, _ip, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "ub200_mlp_forward_act": (_i, [_vp, _vp, _i, _i, _i, _ip, _i, _i, _vp, _vp, _vp, _sz, _i, _vp]),
    "ub200_mlp_backward_act": (_i, [_vp, _vp, _i, _i, _i, _ip, _i, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "ub200_loss_workspace_bytes": (_sz, [_i, _i]),
    "ub200_softmax_ce": (_i, [_vp, _vp, _i, _i, _i, _vp, _i, _vp, _vp, _vp, _sz, _vp]),
    "ub200_dla_loss": (_i, [_vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "ub200_pair_workspace_bytes": (_sz, [_i, _i]),
    "ub200_lambdarank": (_i, [_vp, _vp, _i, _i, _f, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "ub200_pairdebias": (_i, [_vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "ub200_prsrank": (_i, [_vp, _vp, _i, _i, _f, _vp, _i, _vp, _vp, _vp, _sz, _vp]),
    "ub200_em_update": (_i, [_vp, _vp, _vp, _i, _f, _f, _i, _vp]),
    "ub200_publish": (_i, [_vp, _i, _vp, _vp, _vp, _vp]),
    "ub200_opt_workspace_bytes": (_sz, [_sz]),
    "ub200_clip_update": (_i, [_vp, _vp, _vp, _sz, _vp, _f, _f, _f, _i, _vp, _vp, _sz, _vp]),
    "ub200_l2_term": (_i, [_vp, _vp, _sz, _f, _vp, _f, _vp, _vp]),
    "ub200_pl_sample": (_i, [_vp, _vp, _i, _i, _i, _f, ctypes.c_ulonglong, ctypes.c_ulonglong, _vp, _vp]),
    "ub200_click_batch": (_i, [_vp, _vp, _i, _i, _vp, _i, _vp, _i, _i, _i, _i, _i, _i, ctypes.c_ulonglong,
                               ctypes.c_ulonglong, _vp, _vp, _vp, _vp]),
    "ub200_click_batch_model": (_i, [_vp, _vp, _i, _i, _vp, _i, _vp, _i, _i, _i, _i, _i, _i, _i, ctypes.c_ulonglong,
                                     ctypes.c_ulonglong, _vp, _vp, _vp, _vp]),
    "ub200_peer_ctl_bytes": (_sz, []),
    "ub200_peer_flag_bytes": (_sz, [_i]),
    "ub200_peer_allreduce": (_i, [_vp, _vp, _i, _i, _vp, _sz, _vp, _vp]),
    "ub200_dp_inbox_bytes": (_sz, [_i, _sz]),
    "ub200_dp_flag_bytes": (_sz, [_i]),
    "ub200_dp_ctl_bytes": (_sz, []),
    "ub200_dp_reduce_update": (_i, [_vp, _sz, _vp, _vp, _i, _i, _vp, _vp, _sz, ctypes.c_longlong, _f, _f, _f, _i, _vp,
                                    _vp, _vp]),
    "ub200_dp_reduce_update_publish": (_i, [_vp, _sz, _vp, _vp, _i, _i, _vp, _vp, _sz, ctypes.c_longlong, _f, _f, _f, _i,
                                            _vp, _vp, _vp, ctypes.c_longlong, _i, _vp, _vp, _vp]),
}


class UltraB200Error(RuntimeError):
    pass


def _load():
    path = os.environ.get("UB200_LIB") or _build.LIBPATH     # UB200_LIB: debugging builds (e.g. timeline stamps)
    if not os.path.isfile(path):
        try:
            _build.build()
        except Exception as e:  # noqa: BLE001
            raise UltraB200Error(
                "libultra_b200.so is missing and could not be built with nvcc (%s). The CUDA extension is "
                "mandatory: there is no CPU / eager fallback. Run `python -m ultra_pytorch_b200.build`." % (e,))
    lib = ctypes.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError here means the .so is stale: rebuild
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()
LIB_PATH = _build.LIBPATH


def check(rc, what):
    if rc != 0:
        msg = lib.ub200_last_error()
        raise UltraB200Error("%s failed (code %d): %s" % (what, rc, msg.decode() if msg else "?"))


def int_array(values):
    arr = (ctypes.c_int * max(1, len(values)))(*values)
    return arr
