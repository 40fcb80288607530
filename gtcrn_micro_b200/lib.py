"""ctypes binding of libgtcrn_b200.so (the C ABI in include/gtcrn_b200.h).

There is deliberately NO fallback: if the shared library is missing or a call fails, a
`GtcrnB200Error` is raised.  PyTorch is only used by callers for device memory and streams; this
module passes raw pointers (`tensor.data_ptr()`) and `torch.cuda.current_stream().cuda_stream`.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GTCRNB_LIB") or os.path.join(HERE, "libgtcrn_b200.so")   # GTCRNB_LIB: A/B builds of the same ABI

# every symbol include/gtcrn_b200.h declares (tests check the .so exports all of them)
SYMBOLS = [
    "gtcrnb_abi_version", "gtcrnb_blob_floats", "gtcrnb_create", "gtcrnb_destroy", "gtcrnb_forward_spec",
    "gtcrnb_enhance", "gtcrnb_enhance_host", "gtcrnb_stft", "gtcrnb_istft", "gtcrnb_stream_state_bytes",
    "gtcrnb_stream_reset", "gtcrnb_stream_step_spec", "gtcrnb_stream_step_wav", "gtcrnb_stream_import_caches",
    "gtcrnb_stream_export_caches", "gtcrnb_debug_activation", "gtcrnb_launch_count", "gtcrnb_host_alloc",
    "gtcrnb_host_free", "gtcrnb_last_error", "gtcrnb_profile_enable", "gtcrnb_profile_read",
    "gtcrnb_profile_num_classes", "gtcrnb_profile_class_name", "gtcrnb_pair_metrics", "gtcrnb_host_alloc_wc",
    "gtcrnb_enhance_host_pcm16", "gtcrnb_pcm16_to_f32", "gtcrnb_f32_to_pcm16", "gtcrnb_stream_flush_wav",
    "gtcrnb_create_variant", "gtcrnb_blob_floats_variant", "gtcrnb_stream_state_bytes_variant", "gtcrnb_tc_mask",
]


class GtcrnB200Error(RuntimeError):
    pass


_lib = None
_lock = threading.Lock()


def load() -> C.CDLL:
    """Load the shared library (built by `python -m gtcrn_micro_b200.build`).  Raises if absent."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise GtcrnB200Error(f"{LIB_PATH} not found: build it with `python -m gtcrn_micro_b200.build` "
                                 "(there is no CPU / PyTorch fallback for this path)")
        lib = C.CDLL(LIB_PATH)
        p, i64, sz, vp, cp = C.POINTER(C.c_float), C.c_int64, C.c_size_t, C.c_void_p, C.c_char_p
        i64p = C.POINTER(C.c_int64)
        sig = {
            "gtcrnb_abi_version": (C.c_int, []),
            "gtcrnb_blob_floats": (sz, []),
            "gtcrnb_create": (C.c_int, [vp, sz, C.POINTER(vp)]),
            "gtcrnb_create_variant": (C.c_int, [vp, sz, C.c_int, C.POINTER(vp)]),
            "gtcrnb_blob_floats_variant": (sz, [C.c_int]),
            "gtcrnb_stream_state_bytes_variant": (sz, [C.c_int, i64]),
            "gtcrnb_destroy": (None, [vp]),
            "gtcrnb_forward_spec": (C.c_int, [vp, vp, i64p, i64, i64, vp, i64p, vp]),
            "gtcrnb_enhance": (C.c_int, [vp, vp, i64, i64, i64, vp, i64, vp]),
            "gtcrnb_enhance_host": (C.c_int, [vp, vp, i64, i64, vp]),
            "gtcrnb_enhance_host_pcm16": (C.c_int, [vp, vp, i64, i64, vp]),
            "gtcrnb_pcm16_to_f32": (C.c_int, [vp, vp, vp, i64, vp]),
            "gtcrnb_f32_to_pcm16": (C.c_int, [vp, vp, vp, i64, vp]),
            "gtcrnb_stft": (C.c_int, [vp, vp, i64, i64, i64, vp, vp]),
            "gtcrnb_istft": (C.c_int, [vp, vp, i64, i64, vp, i64, vp]),
            "gtcrnb_stream_state_bytes": (sz, [i64]),
            "gtcrnb_stream_reset": (C.c_int, [vp, vp, i64, vp]),
            "gtcrnb_stream_step_spec": (C.c_int, [vp, vp, vp, vp, i64, vp]),
            "gtcrnb_stream_step_wav": (C.c_int, [vp, vp, vp, vp, i64, vp]),
            "gtcrnb_stream_flush_wav": (C.c_int, [vp, vp, vp, i64, vp, i64, vp]),
            "gtcrnb_stream_import_caches": (C.c_int, [vp, vp, i64, vp, vp, C.POINTER(vp), vp]),
            "gtcrnb_stream_export_caches": (C.c_int, [vp, vp, i64, vp, vp, C.POINTER(vp), vp]),
            "gtcrnb_debug_activation": (C.c_int, [vp, cp, vp, sz, vp]),
            "gtcrnb_launch_count": (i64, [vp]),
            "gtcrnb_tc_mask": (C.c_int, [vp]),
            "gtcrnb_host_alloc": (vp, [sz]),
            "gtcrnb_host_alloc_wc": (vp, [sz]),
            "gtcrnb_host_free": (None, [vp]),
            "gtcrnb_last_error": (cp, []),
            "gtcrnb_profile_enable": (C.c_int, [vp, C.c_int]),
            "gtcrnb_profile_read": (C.c_int, [vp, C.c_int, C.POINTER(C.c_double), i64p, i64p]),
            "gtcrnb_profile_num_classes": (C.c_int, []),
            "gtcrnb_profile_class_name": (cp, [C.c_int]),
            "gtcrnb_pair_metrics": (C.c_int, [vp, vp, i64, vp, i64, i64, i64, vp, vp]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
        return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().gtcrnb_last_error()
        raise GtcrnB200Error(f"{what} failed (code {rc}): {msg.decode() if msg else ''}")


def strides4(values) -> "C.Array":
    return (C.c_int64 * 4)(*[int(v) for v in values])
