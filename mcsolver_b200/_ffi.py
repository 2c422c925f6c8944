"""ctypes binding of libmcsolver_b200.so (include/mcsolver_b200.h).  No PyTorch, no fallback:
if the library is missing it is built with nvcc; if that fails, or no GPU is usable, the calls raise.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIBPATH = os.path.join(HERE, "libmcsolver_b200.so")
_lib = None


class McgError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("mcsolver_b200 error %d: %s" % (code, msg))
        self.code, self.msg = code, msg

    def __reduce__(self):
        # must survive pickling: the reference runs MCMainFunction inside multiprocessing.Pool workers (win.py:90-91) and
        # an exception that cannot be rebuilt in the parent wedges the pool instead of surfacing
        return (McgError, (self.code, self.msg))


class Tables(C.Structure):
    _fields_ = [("model", C.c_int32), ("N", C.c_int32), ("maxL", C.c_int32), ("S", C.c_void_p), ("D", C.c_void_p),
                ("nlink", C.c_void_p), ("J", C.c_void_p), ("nbr", C.c_void_p), ("nTri", C.c_int32), ("tri", C.c_void_p),
                ("nLat", C.c_int32), ("pairs", C.c_void_p), ("nG", C.c_int32), ("maxG", C.c_int32), ("groups", C.c_void_p),
                ("nR", C.c_int32), ("nC", C.c_int32), ("rOrb", C.c_void_p), ("rCl", C.c_void_p), ("rNbr", C.c_void_p),
                ("ignoreOffDiag", C.c_int32)]


class Bond(C.Structure):
    _fields_ = [("src", C.c_int32), ("tgt", C.c_int32), ("d", C.c_int32 * 3), ("J", C.c_double * 9)]


class LatticeDesc(C.Structure):
    _fields_ = [("model", C.c_int32), ("L", C.c_int32 * 3), ("norb", C.c_int32), ("S", C.c_void_p), ("D", C.c_void_p),
                ("nbond", C.c_int32), ("bonds", C.c_void_p), ("pair_s", C.c_int32), ("pair_t", C.c_int32),
                ("pair_d", C.c_int32 * 3), ("ncircuit", C.c_int32), ("circuits", C.c_void_p), ("ngroup", C.c_int32),
                ("group_mask", C.c_void_p), ("group_in_sc", C.c_int32), ("block_spin", C.c_int32)]


class Config(C.Structure):
    _fields_ = [("precision", C.c_int32), ("nReplica", C.c_int32), ("beta", C.c_void_p), ("field", C.c_void_p),
                ("seed", C.c_uint64), ("replica_offset", C.c_int32), ("device", C.c_int32)]


# name -> (restype, argtypes); also the list the CPU test checks against include/mcsolver_b200.h
_vp, _i, _i64, _d, _u64 = C.c_void_p, C.c_int, C.c_int64, C.c_double, C.c_uint64
SIGNATURES = {
    "mcg_last_error": (C.c_char_p, []),
    "mcg_version": (_i, []),
    "mcg_device_count": (_i, [_vp]),
    "mcg_create_tables": (_i, [C.POINTER(Tables), C.POINTER(Config), C.POINTER(_vp)]),
    "mcg_create_lattice": (_i, [C.POINTER(LatticeDesc), C.POINTER(Config), C.POINTER(_vp)]),
    "mcg_destroy": (_i, [_vp]),
    "mcg_jit_check": (_i, [C.POINTER(LatticeDesc), _i, _vp, _vp, _i]),
    "mcg_num_colours": (_i, [_vp, _vp]),
    "mcg_colour_order": (_i, [_vp, _vp]),
    "mcg_rng_layout": (_i, [_vp, _vp, _vp]),
    "mcg_set_params": (_i, [_vp, _vp, _vp]),
    "mcg_recycle": (_i, [_vp, _vp, _vp, _u64, _i]),
    "mcg_create_lattice_slab": (_i, [C.POINTER(LatticeDesc), C.POINTER(Config), _i, _i, _vp, C.POINTER(_vp)]),
    "mcg_slab_info": (_i, [_vp, _vp]),
    "mcg_slab_plan": (_i, [C.POINTER(LatticeDesc), _i, _i, _i, _vp]),
    "mcg_slab_sync": (_i, [_vp]),
    "mcg_init_spins": (_i, [_vp, _d]),
    "mcg_set_spins": (_i, [_vp, _i, _vp]),
    "mcg_get_spins": (_i, [_vp, _i, _vp]),
    "mcg_energy": (_i, [_vp, _i, _vp, _vp, _vp]),
    "mcg_metropolis_sweeps": (_i, [_vp, _i64, _d]),
    "mcg_timed_sweeps": (_i, [_vp, _i64, _d, _i, _vp]),
    "mcg_wolff_steps": (_i, [_vp, _i64]),
    "mcg_measure": (_i, [_vp]),
    "mcg_reset_measurements": (_i, [_vp]),
    "mcg_results": (_i, [_vp, _i, _vp, _vp]),
    "mcg_counters": (_i, [_vp, _i, _vp, _vp, _vp]),
    "mcg_wolff_frontier_steps": (_i, [_vp, _i, _vp]),
    "mcg_launch_count": (_i, [_vp, _vp]),
    "mcg_jit_launch_count": (_i, [_vp, _vp]),
    "mcg_jit_module_key": (_i, [_vp, _i, _vp]),
    "mcg_profile_passes": (_i, [_vp, _i]),
    "mcg_profile_read": (_i, [_vp, _vp, _vp]),
    "mcg_run": (_i, [_vp, _i, _i64, _i64, _i64, _i, _vp]),
    "mcg_run_on": (_i, [C.POINTER(Tables), _i, _i64, _i64, _i64, _d, _d, _i, _u64, _i, _vp, _vp, _vp]),
    "mcg_run_ising": (_i, [C.POINTER(Tables), _i, _i64, _i64, _i64, _d, _i, _u64, _i, _vp, _vp]),
    "mcg_pt_configure": (_i, [_vp, _i]),
    "mcg_pt_state": (_i, [_vp, _vp]),
    "mcg_pt_set_labels": (_i, [_vp, _vp, _vp, _vp]),
    "mcg_pt_decide": (_i, [_i, _vp, _vp, _vp, _vp, _vp, _i, _u64, _u64, _vp]),
    "mcg_comm_unique_id": (_i, [_vp, _i]),
    "mcg_pt_setup": (_i, [_vp, _i, _i, _vp, _i, _vp, _vp]),
    "mcg_pt_run": (_i, [_vp, _i64, _i64, _i, _vp]),
    "mcg_pt_stats": (_i, [_vp, _vp, _vp, _vp]),
    "mcg_pt_reduce": (_i, [_vp]),
    "mcg_pt_results": (_i, [_vp, _i, _vp, _vp]),
    "mcg_acc_get": (_i, [_vp, _i, _vp, _vp]),
    "mcg_acc_set": (_i, [_vp, _i, _vp]),
}


def lib():
    """Load (building first if needed) the engine library.  Raises if it cannot be had."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIBPATH):
            from . import build as _build
            _build.build()
        L = C.CDLL(LIBPATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)   # AttributeError if the library lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(code):
    if code != 0:
        raise McgError(code, lib().mcg_last_error().decode("utf-8", "replace"))


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None and a.size else None
