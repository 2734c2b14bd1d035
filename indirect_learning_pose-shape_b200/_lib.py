"""ctypes binding of libsmpl_b200.so (include/smpl_b200.h).

There is no CPU path: if the library is missing, or no CUDA device is usable, every entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libsmpl_b200.so")

ABI_VERSION = 3
OP_DECODE_FWD, OP_DECODE_BWD, OP_SILHOUETTE_FWD, OP_SILHOUETTE_BWD, OP_FULL_FWD, OP_FULL_BWD = 0, 1, 2, 3, 4, 5

_f32p = C.POINTER(C.c_float)
_i32p = C.POINTER(C.c_int32)


class HostModel(C.Structure):
    """struct SmplB200HostModel"""
    _fields_ = [("num_verts", C.c_int32), ("num_joints", C.c_int32), ("num_betas", C.c_int32),
                ("num_pose_basis", C.c_int32), ("num_reg_joints", C.c_int32),
                ("v_template", _f32p), ("shapedirs", _f32p), ("posedirs", _f32p), ("J_regressor", _f32p),
                ("lbs_weights", _f32p), ("parents", _i32p), ("joint_regressor", _f32p)]


class KernelStat(C.Structure):
    """struct SmplB200KernelStat"""
    _fields_ = [("name", C.c_char_p), ("launches", C.c_longlong), ("total_ms", C.c_double)]


# name -> (restype, argtypes): exactly the symbols include/smpl_b200.h declares
SIGNATURES = {
    "smpl_b200_abi_version": (C.c_int, []),
    "smpl_b200_last_error": (C.c_char_p, []),
    "smpl_b200_launch_count": (C.c_uint64, []),
    "smpl_b200_profile_enable": (C.c_int, [C.c_int]),
    "smpl_b200_profile_collect": (C.c_int, [C.POINTER(KernelStat), C.c_int, C.POINTER(C.c_int)]),
    "smpl_b200_model_create": (C.c_int, [C.POINTER(HostModel), C.c_int, C.POINTER(C.c_void_p)]),
    "smpl_b200_model_destroy": (None, [C.c_void_p]),
    "smpl_b200_model_num_verts": (C.c_int, [C.c_void_p]),
    "smpl_b200_model_lbs_width": (C.c_int, [C.c_void_p]),
    "smpl_b200_parts_create": (C.c_int, [C.c_int, C.c_int, _i32p, _i32p, C.c_int, C.POINTER(C.c_void_p)]),
    "smpl_b200_parts_destroy": (None, [C.c_void_p]),
    "smpl_b200_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]),
    "smpl_b200_decode_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]),
    "smpl_b200_decode_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "smpl_b200_full_state_bytes": (C.c_size_t, [C.c_void_p, C.c_int, C.c_int, C.c_int]),
    "smpl_b200_full_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "smpl_b200_full_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "smpl_b200_project_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "smpl_b200_project_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                        C.c_void_p, C.c_void_p]),
    "smpl_b200_mask_fwd": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "smpl_b200_seg_saved_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "smpl_b200_seg_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                    C.c_void_p, C.c_void_p]),
    "smpl_b200_seg_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                    C.c_int, C.c_void_p, C.c_void_p]),
    "smpl_b200_seg_loss_state_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "smpl_b200_seg_loss_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_float,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "smpl_b200_seg_loss_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                         C.c_int, C.c_void_p, C.c_void_p]),
    "smpl_b200_dense_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "smpl_b200_dense_fwd": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                      C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]),
    "smpl_b200_dense_bwd": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                      C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                      C.c_void_p, C.c_size_t, C.c_void_p]),
    "smpl_b200_axpy_cols": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_int, C.c_void_p,
                                      C.c_int, C.c_void_p]),
    "smpl_b200_silhouette_fwd": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t,
                                           C.c_void_p]),
    "smpl_b200_silhouette_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                           C.c_size_t, C.c_void_p]),
    "smpl_b200_focal_loss_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_float, C.c_void_p,
                                           C.c_int, C.c_void_p, C.c_void_p]),
    "smpl_b200_focal_loss_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_float,
                                           C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "smpl_b200_renderer_create": (C.c_int, [C.c_int, _i32p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "smpl_b200_renderer_destroy": (None, [C.c_void_p]),
    "smpl_b200_render_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int]),
    "smpl_b200_render": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                   C.c_int, _f32p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                   C.c_size_t, C.c_void_p]),
}

_lock = threading.Lock()
_lib = None


class SmplB200Error(RuntimeError):
    pass


def load(build_if_missing: bool = False) -> C.CDLL:
    """Load the CUDA library.  Raises SmplB200Error if it has not been built (there is no fallback)."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            if build_if_missing:
                from . import build_ext
                build_ext.build()
            else:
                raise SmplB200Error(
                    "%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                    "(nvcc, sm_100a).  This package has no CPU or PyTorch fallback." % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError here = header/library mismatch, keep it loud
            fn.restype = res
            fn.argtypes = args
        if lib.smpl_b200_abi_version() != ABI_VERSION:
            raise SmplB200Error("libsmpl_b200.so ABI %d != binding ABI %d; rebuild" %
                                (lib.smpl_b200_abi_version(), ABI_VERSION))
        _lib = lib
        return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().smpl_b200_last_error()
        raise SmplB200Error("%s failed (%d): %s" % (what or "smpl_b200 call", rc, (msg or b"").decode()))


def launch_count() -> int:
    return int(load().smpl_b200_launch_count())


def profile_enable(on: bool) -> None:
    """Bracket every kernel launch of the library with CUDA events on its stream (see smpl_b200_profile_enable)."""
    check(load().smpl_b200_profile_enable(1 if on else 0), "smpl_b200_profile_enable")


def profile_collect() -> dict:
    """{kernel name: (launches, total_ms)} since the last collect; synchronises the recorded events."""
    stats = (KernelStat * 32)()
    n = C.c_int(0)
    check(load().smpl_b200_profile_collect(stats, 32, C.byref(n)), "smpl_b200_profile_collect")
    return {stats[i].name.decode(): (int(stats[i].launches), float(stats[i].total_ms)) for i in range(min(n.value, 32))}


def _fp(a: np.ndarray):
    return a.ctypes.data_as(_f32p)


def make_host_model(hm) -> tuple:
    """Pack a smpl_io.SmplHostModel into the C struct.  Returns (struct, keepalive arrays)."""
    f32 = lambda a: np.ascontiguousarray(a, np.float32)  # noqa: E731
    arrs = dict(v_template=f32(hm.v_template), shapedirs=f32(hm.shapedirs), posedirs=f32(hm.posedirs),
                J_regressor=f32(hm.J_regressor), lbs_weights=f32(hm.lbs_weights),
                parents=np.ascontiguousarray(hm.parents, np.int32), joint_regressor=f32(hm.joint_regressor))
    V = arrs["v_template"].shape[0]
    s = HostModel(num_verts=V, num_joints=arrs["lbs_weights"].shape[1], num_betas=arrs["shapedirs"].shape[0],
                  num_pose_basis=arrs["posedirs"].shape[0], num_reg_joints=arrs["joint_regressor"].shape[1],
                  v_template=_fp(arrs["v_template"]), shapedirs=_fp(arrs["shapedirs"]), posedirs=_fp(arrs["posedirs"]),
                  J_regressor=_fp(arrs["J_regressor"]), lbs_weights=_fp(arrs["lbs_weights"]),
                  parents=arrs["parents"].ctypes.data_as(_i32p), joint_regressor=_fp(arrs["joint_regressor"]))
    return s, arrs
