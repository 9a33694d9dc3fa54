"""Builds (g++) and loads the host emulation of the per-campaign device routines.

TEST SCAFFOLDING: see tests/hostemu/hostemu.cpp.  Never imported by the product package.
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "hostemu.cpp")
_SO = os.path.join(_HERE, "_hostemu.so")
_CSRC = os.path.join(os.path.dirname(os.path.dirname(_HERE)), "stpbenchmark_b200", "csrc")
_lib = None


def _stale():
    if not os.path.exists(_SO):
        return True
    t = os.path.getmtime(_SO)
    deps = [_SRC, os.path.join(_HERE, "hostemu_var.inc")] + [
        os.path.join(_CSRC, f) for f in os.listdir(_CSRC) if f.endswith(".cuh")]
    return any(os.path.getmtime(p) > t for p in deps if os.path.exists(p))


def load():
    global _lib
    if _lib is None:
        if _stale():
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++", _SRC, "-o", _SO,
                                   "-lm"])
        _lib = ctypes.CDLL(_SO)
        _lib.emu_log_h.restype = ctypes.c_double
        _lib.emu_log_h.argtypes = [ctypes.c_double]
        _lib.emu_log_h_mc.restype = ctypes.c_double
        _lib.emu_log_h_mc.argtypes = [ctypes.c_double]
        _lib.emu_log_ei.restype = ctypes.c_double
        _lib.emu_log_ei.argtypes = [ctypes.c_double] * 3
        _lib.emu_lbfgsb_new.restype = ctypes.c_void_p
        _lib.emu_set_model.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_double]
        _lib.emu_set_model.restype = None
        _lib.emu_acq_value.restype = ctypes.c_double
        _lib.emu_acq_value.argtypes = [ctypes.c_int] + [ctypes.c_double] * 4
        _lib.emu_log_ei_t.restype = ctypes.c_double
        _lib.emu_log_ei_t.argtypes = [ctypes.c_double] * 4
    return _lib
