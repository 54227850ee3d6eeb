"""ctypes binding of libcasync_b200.so (C ABI: include/casync_b200.h).  No fallback: a missing library is an error."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcasync_b200.so")

F_BF16, F_FP32, F_OUT_U8_HWC = 0, 1, 2

# every symbol include/casync_b200.h declares: (name, restype, argtypes)
_P, _I, _SZ, _I64 = C.c_void_p, C.c_int, C.c_size_t, C.c_int64
SYMBOLS = [
    ("casync_version", C.c_char_p, []),
    ("casync_last_error", C.c_char_p, []),
    ("casync_weight_entry_count", _I, []),
    ("casync_weight_entry", _I, [_I, C.POINTER(C.c_char_p), C.POINTER(_SZ)]),
    ("casync_plan_create", _I, [_P, _P, _SZ, C.POINTER(_I64), _I, C.POINTER(_P)]),
    ("casync_plan_destroy", None, [_P]),
    ("casync_chunk_frames", _I, [_P]),
    ("casync_workspace_bytes", _SZ, [_P, _I]),
    ("casync_forward", _I, [_P, _P, _P, _P, _P, _I, C.c_uint, _P]),
    ("casync_prepare_inputs", _I, [_P, _P, _I, _P, _P, _P, _I, _P]),
    ("casync_stage_view", _I, [_P, _I, C.c_char_p, C.POINTER(_SZ), C.POINTER(_I64), C.POINTER(_I64), C.POINTER(_I64)]),
    ("casync_launches_per_forward", _I64, [_P, _I]),
    ("casync_graph_replays", _I64, [_P]),
    ("casync_ir_count", _I, []),
    ("casync_ir_info", _I, [_I, C.POINTER(C.c_char_p)] + [C.POINTER(_I)] * 5),
    ("casync_stage_scratch_bytes", _SZ, [_P, _I]),
    ("casync_ir_block", _I, [_P, _I, _P, _P, _P, _I, _P]),
    ("casync_audio_cnn", _I, [_P, _P, _P, _P, _I, _P]),
    ("casync_fusion_attention", _I, [_P, _P, _P, _P, _P, _I, _P]),
    ("casync_up_block", _I, [_P, _I, _P, _P, _P, _P, _I, _P]),
    ("casync_up_first", _I, [_P, _I, _P, _P, _P, _P, _I, _P]),
    ("casync_blend_paste", _I, [_P, _I, _I, _P, _I, _P, _P, _P, _I, _P]),
]



class LaunchRecord(C.Structure):
    _fields_ = [("name", C.c_char * 48), ("ms", C.c_float), ("flops", C.c_double), ("bytes", C.c_double)]


SYMBOLS.append(("casync_forward_profiled", _I,
                [_P, _P, _P, _P, _P, _I, C.c_uint, _P, C.POINTER(LaunchRecord), _I, C.POINTER(_I)]))

_lib = None


def load():
    """Load the CUDA extension or raise (the product path never degrades to a CPU implementation)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "calipsync_b200: CUDA extension %s is missing -- build it with "
            "`python -c 'import __graft_entry__ as g; g.build()'` (or `make -C calipsync_b200/csrc`). "
            "There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, res, args in SYMBOLS:
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        raise RuntimeError("calipsync_b200: %s failed (code %d): %s" % (what, rc, load().casync_last_error().decode()))


def weight_schema():
    lib = load()
    out = []
    for i in range(lib.casync_weight_entry_count()):
        name, nbytes = C.c_char_p(), C.c_size_t()
        check(lib.casync_weight_entry(i, C.byref(name), C.byref(nbytes)), "casync_weight_entry")
        out.append((name.value.decode(), nbytes.value))
    return out


def ir_table():
    lib = load()
    out = []
    for i in range(lib.casync_ir_count()):
        name = C.c_char_p()
        v = [C.c_int() for _ in range(5)]
        check(lib.casync_ir_info(i, C.byref(name), *[C.byref(a) for a in v]), "casync_ir_info")
        out.append(dict(index=i, name=name.value.decode(), cin=v[0].value, cout=v[1].value, h_in=v[2].value,
                        stride=v[3].value, residual=v[4].value))
    return out
