"""Bindings of the C ABI declared in include/nunet_b200.h (csrc/libnunet_b200.so).

`pyb()` is the binding the package uses: the thin pybind11 module csrc/_nunet_pybind (built by csrc/build.sh), one function
per extern "C" entry point, addresses passed as integers.  `lib()` is the same ABI bound with ctypes -- what a maintainer of
the reference would write without a compiler, used by the ABI tests to check every declared symbol.

There is deliberately no fallback: if the CUDA library or the pybind11 module is missing, or the device is not a B200, the
import / create call raises.  Nothing in this package computes the network on the CPU.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libnunet_b200.so")

NUNET_VARIANT_LSTM, NUNET_VARIANT_DDB, NUNET_VARIANT_LSTM_HYBRID = 0, 1, 2
NUNET_CTFA_CAUSAL_AVG32, NUNET_CTFA_FRAME_DIV32 = 0, 1
NUNET_DC_ZERO, NUNET_DC_EDGE = 0, 1

# every symbol include/nunet_b200.h declares (tests check the .so exports exactly these)
EXPORTS = [
    "nunet_last_error", "nunet_abi_version", "nunet_blob_validate", "nunet_create", "nunet_destroy", "nunet_num_frames",
    "nunet_forward_wav_dev", "nunet_forward_wav_host", "nunet_forward_mag_dev",
    "nunet_stream_reset", "nunet_stream_step_mag_dev", "nunet_stream_step_wav_dev", "nunet_stream_step_wav_host",
    "nunet_state_count", "nunet_state_name", "nunet_state_numel", "nunet_state_export", "nunet_state_import", "nunet_state_generation",
    "nunet_last_launch_count", "nunet_debug_read",
    "nunet_profile_enable", "nunet_profile_count", "nunet_profile_entry",
]


class NunetConfig(C.Structure):
    _fields_ = [("variant", C.c_int32), ("device", C.c_int32), ("max_frames", C.c_int32),
                ("max_streams", C.c_int32), ("ctfa_mode", C.c_int32), ("dc_mode", C.c_int32),
                ("stream_ctfa_history", C.c_int32), ("chunk_frames", C.c_int32)]


class NunetError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"nunet_b200 error {code}: {msg}")
        self.code = code


_lib = None
_pyb = None


def pyb():
    """The pybind11 layer over the C ABI (csrc/_nunet_pybind*.so)."""
    global _pyb
    if _pyb is not None:
        return _pyb
    import glob
    import importlib.util
    cands = glob.glob(os.path.join(_HERE, "csrc", "_nunet_pybind*.so"))
    if not cands or not os.path.exists(LIB_PATH):
        raise ImportError(f"{os.path.join(_HERE, 'csrc')}: libnunet_b200.so / _nunet_pybind*.so missing: build them with "
                          "`python -c 'import __graft_entry__ as g; g.build()'` (nvcc, sm_100a). There is no CPU fallback.")
    spec = importlib.util.spec_from_file_location("_nunet_pybind", cands[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    if mod.abi_version() != mod.ABI_VERSION:
        raise ImportError("nunet_b200: pybind11 module and libnunet_b200.so disagree on the ABI version")
    _pyb = mod
    return mod


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(nvcc, sm_100a). There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    fp, vp, i, ll = C.POINTER(C.c_float), C.c_void_p, C.c_int, C.c_longlong
    L.nunet_last_error.restype = C.c_char_p
    L.nunet_last_error.argtypes = []
    L.nunet_abi_version.restype = i
    L.nunet_blob_validate.restype = ll
    L.nunet_blob_validate.argtypes = [vp, C.c_size_t, i]
    L.nunet_create.restype = i
    L.nunet_create.argtypes = [C.POINTER(NunetConfig), vp, C.c_size_t, C.POINTER(vp)]
    L.nunet_destroy.restype = None
    L.nunet_destroy.argtypes = [vp]
    L.nunet_num_frames.restype = i
    L.nunet_num_frames.argtypes = [i]
    L.nunet_forward_wav_dev.restype = i
    L.nunet_forward_wav_dev.argtypes = [vp, vp, i, i, vp, vp, vp]
    L.nunet_forward_wav_host.restype = i
    L.nunet_forward_wav_host.argtypes = [vp, vp, i, i, vp, vp]
    L.nunet_forward_mag_dev.restype = i
    L.nunet_forward_mag_dev.argtypes = [vp, vp, i, i, vp, vp]
    L.nunet_stream_reset.restype = i
    L.nunet_stream_reset.argtypes = [vp, i, i, vp]
    L.nunet_stream_step_mag_dev.restype = i
    L.nunet_stream_step_mag_dev.argtypes = [vp, vp, i, vp, vp]
    L.nunet_stream_step_wav_dev.restype = i
    L.nunet_stream_step_wav_dev.argtypes = [vp, vp, i, vp, vp, vp]
    L.nunet_stream_step_wav_host.restype = i
    L.nunet_stream_step_wav_host.argtypes = [vp, vp, i, vp]
    L.nunet_state_count.restype = i
    L.nunet_state_count.argtypes = [vp]
    L.nunet_state_name.restype = i
    L.nunet_state_name.argtypes = [vp, i, C.c_char_p, i]
    L.nunet_state_numel.restype = i
    L.nunet_state_numel.argtypes = [vp, C.c_char_p]
    L.nunet_state_export.restype = i
    L.nunet_state_export.argtypes = [vp, i, C.c_char_p, vp]
    L.nunet_state_import.restype = i
    L.nunet_state_import.argtypes = [vp, i, C.c_char_p, vp]
    L.nunet_state_generation.restype = ll
    L.nunet_state_generation.argtypes = [vp]
    L.nunet_last_launch_count.restype = i
    L.nunet_last_launch_count.argtypes = [vp]
    L.nunet_profile_enable.restype = i
    L.nunet_profile_enable.argtypes = [vp, i]
    L.nunet_profile_count.restype = i
    L.nunet_profile_count.argtypes = [vp]
    L.nunet_profile_entry.restype = i
    L.nunet_profile_entry.argtypes = [vp, i, C.c_char_p, i, C.POINTER(C.c_float), C.POINTER(C.c_double)]
    L.nunet_debug_read.restype = ll
    L.nunet_debug_read.argtypes = [vp, C.c_char_p, vp, ll]
    _ = fp
    _lib = L
    return L


def check(rc: int) -> int:
    if rc < 0:
        raise NunetError(rc, pyb().last_error())
    return rc
