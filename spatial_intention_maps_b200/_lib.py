"""ctypes binding of ``libsimq.so`` (C-ABI declared in ``include/simq.h``).

There is no CPU or PyTorch fallback: if the library is missing or no B200 is present every entry
point of the package raises :class:`SimqError`.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('SIMQ_LIB_PATH') or os.path.join(_HERE, 'libsimq.so')      # override: A/B of two builds in one session

BACKEND_UMMA, BACKEND_FMA = 0, 1
PRECISION_PARITY, PRECISION_BF16 = 0, 1
SCHEDULE_SERIAL, SCHEDULE_LANES = 0, 1
X_NCHW, X_NHWC, X_NHWC_PLUS1 = 0, 1, 2
N_BN = 22
N_PARAM_TENSORS = 70


class SimqError(RuntimeError):
    pass


_lib = None

_c_ctx = C.c_void_p
_p = C.c_void_p
_PROTOS = {
    'simq_last_error': (C.c_char_p, []),
    'simq_version': (C.c_int, []),
    'simq_layout': (C.c_int, [C.c_int, C.c_int, _p, _p, _p, _p]),
    'simq_ctx_create': (C.c_int, [C.POINTER(_c_ctx), C.c_int, C.c_int, C.c_int, C.c_int]),
    'simq_ctx_destroy': (None, [_c_ctx]),
    'simq_set_backend': (C.c_int, [_c_ctx, C.c_int]),
    'simq_set_precision': (C.c_int, [_c_ctx, C.c_int]),
    'simq_set_schedule': (C.c_int, [_c_ctx, C.c_int]),
    'simq_set_backward_terms': (C.c_int, [_c_ctx, C.c_int, C.c_int, C.c_int]),
    'simq_workspace_bytes': (C.c_size_t, [_c_ctx]),
    'simq_fcn_forward': (C.c_int, [_c_ctx, _p, _p, _p, _p, C.c_int, C.c_int, C.c_int, C.c_int, _p, C.c_uint64, _p]),
    'simq_fcn_backward': (C.c_int, [_c_ctx, _p, _p, C.c_int, _p, C.c_int, _p, _p]),
    'simq_dqn_tail': (C.c_int, [_c_ctx, _p, _p, _p, _p, _p, _p, C.c_float, C.c_int, C.c_int, C.c_int, _p, _p, _p]),
    'simq_check_device_errors': (C.c_int, [_c_ctx, _p]),
    'simq_sgd_step': (C.c_int, [_c_ctx, _p, _p, _p, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int, _p, _p]),
    'simq_train_step': (C.c_int, [_c_ctx, _p, _p, _p, _p, _p, C.c_uint64, _p, _p, _p, _p, C.c_int, _p, _p, _p,
                                  C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float,
                                  C.c_int, C.c_int, C.c_int, _p, _p]),
    'simq_train_step_phase': (C.c_int, [_c_ctx, _p, _p, _p, _p, _p, C.c_uint64, _p, _p, _p, _p, C.c_int, _p, _p, _p,
                                        C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float,
                                        C.c_int, C.c_int, C.c_int, _p, C.c_int, _p]),
    'simq_set_next_state_event': (C.c_int, [_c_ctx, _p]),
    'simq_gather_rows': (C.c_int, [_p, _p, C.c_int, C.c_int64, _p, _p]),
    'simq_bce_tail': (C.c_int, [_c_ctx, _p, _p, C.c_int64, C.c_int64, _p, _p, _p]),
    'simq_intention_step': (C.c_int, [_c_ctx, _p, _p, _p, _p, _p, _p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float,
                                      C.c_int, C.c_int, _p, _p]),
    'simq_greedy_action': (C.c_int, [_c_ctx, _p, _p, _p, C.c_int, C.c_int, _p, _p, C.c_uint64, _p]),
    'simq_launch_count': (C.c_int64, [_c_ctx]),
    'simq_profile': (C.c_int, [C.c_int, _p, _p, _p]),
    'simq_profile_issued': (C.c_int, [_p]),
    'simq_debug_get': (C.c_int, [_c_ctx, C.c_int, C.c_int, C.c_int, _p, _p, _p]),
    'simq_debug_tensor_name': (C.c_char_p, [C.c_int]),
    'simq_test_conv': (C.c_int, [_c_ctx, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _p, _p, _p, _p, _p]),
}
EXPORTS = tuple(_PROTOS)


def lib():
    """The loaded library; raises SimqError when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SimqError(f'{LIB_PATH} not found: build it with __graft_entry__.build() '
                            '(spatial_intention_maps_b200/csrc/build.sh). There is no CPU fallback.')
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in _PROTOS.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        raise SimqError(f'{what} failed (rc={rc}): {lib().simq_last_error().decode()}')


def layout(Cin: int, A: int):
    """(n_params, n_bn, param_offsets[71], bn_offsets[23]) of the flat vectors for FCN(Cin, A)."""
    n_p, n_b = C.c_int64(), C.c_int64()
    po = (C.c_int64 * (N_PARAM_TENSORS + 1))()
    bo = (C.c_int64 * (N_BN + 1))()
    check(lib().simq_layout(Cin, A, C.byref(n_p), C.byref(n_b), po, bo), 'simq_layout')
    return n_p.value, n_b.value, list(po), list(bo)


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class Ctx:
    """Owns one simq_ctx (device workspace for one network at up to ``max_batch`` samples)."""

    def __init__(self, device_index: int, Cin: int, A: int, max_batch: int):
        self.handle = _c_ctx()
        self.max_batch = max_batch
        check(lib().simq_ctx_create(C.byref(self.handle), device_index, Cin, A, max_batch), 'simq_ctx_create')

    def close(self):
        if self.handle:
            lib().simq_ctx_destroy(self.handle)
            self.handle = _c_ctx()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def launches(self) -> int:
        return int(lib().simq_launch_count(self.handle))
