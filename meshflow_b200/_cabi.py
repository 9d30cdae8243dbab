"""ctypes binding of ``libmeshflow_b200.so`` (the C ABI declared in ``include/meshflow_b200.h``).

There is no CPU fallback: if the shared library is missing or was not built for sm_100a the import
of the product path fails loudly.  Build it with ``python -c "import __graft_entry__ as g; g.build()"``
or ``bash meshflow_b200/csrc/build.sh``.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_int, c_int64, c_longlong, c_size_t, c_ulonglong, c_void_p, c_char_p

_LIB_NAME = "libmeshflow_b200.so"
# MESHFLOW_B200_LIB lets kernel-tuning scripts point at an alternative build of the same library
_LIB_PATH = os.environ.get("MESHFLOW_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), _LIB_NAME)


class MeshflowNativeError(RuntimeError):
    """A C-ABI call returned a negative MF_E_* code; the message is ``mf_last_error()``."""

    def __init__(self, code: int, message: str):
        super().__init__(f"meshflow_b200 native error {code}: {message}")
        self.code = code


# name -> (restype, argtypes); mirrors include/meshflow_b200.h one to one
_SIGNATURES = {
    "mf_version": (c_int, []),
    "mf_built_for_sm": (c_int, []),
    "mf_last_error": (c_char_p, []),
    "mf_vertex_motion_workspace_bytes": (c_size_t, [c_int64, c_int, c_int, c_int]),
    "mf_vertex_motion": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int,
                                 c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                 c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "mf_prefix_displacements": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int64, c_void_p]),
    "mf_jacobi_workspace_bytes": (c_size_t, [c_int, c_int64]),
    "mf_jacobi_solve": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int64, c_int64, c_int64, c_int, c_int,
                                c_int, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "mf_warp_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int]),
    "mf_warp_frames": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                               c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                               c_void_p]),
    "mf_warp_prepare": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p,
                                c_void_p, c_size_t, c_void_p]),
    "mf_warp_crop_bounds": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p,
                                    c_void_p, c_size_t, c_void_p]),
    "mf_warp_resize_frames": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                      c_void_p, c_void_p, c_void_p, c_size_t, c_void_p, c_size_t, c_void_p]),
    "mf_crop_resize_workspace_bytes": (c_size_t, [c_int, c_int]),
    "mf_crop_resize": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p,
                               c_void_p, c_size_t, c_void_p]),
    "mf_crop_combine": (c_int, [c_void_p, c_int, c_void_p, c_void_p]),
    "mf_crop_resize_device": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_size_t,
                                      c_void_p]),
    "mf_stability_ratios": (c_int, [c_void_p, c_int, c_int64, c_void_p, c_void_p]),
    "mf_debug_rcp_mismatches": (c_longlong, [c_longlong, c_ulonglong, c_void_p, c_void_p]),
    "mf_debug_force_generic_vertex_motion": (None, [c_int]),
}

_lib = None


def library_path() -> str:
    return _LIB_PATH


def exported_symbols():
    """Names the header declares (used by the CPU-only symbol test)."""
    return list(_SIGNATURES)


def load() -> ctypes.CDLL:
    """Open the shared library once and type every entry point.  Raises if it is not there."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise ImportError(
            f"{_LIB_PATH} not found: the CUDA extension is not built. Run "
            "`python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc); "
            "meshflow_b200 has no CPU fallback.")
    lib = ctypes.CDLL(_LIB_PATH)
    for name, (restype, argtypes) in _SIGNATURES.items():
        fn = getattr(lib, name)            # AttributeError if the build is stale
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(code: int) -> None:
    if code != 0:
        msg = load().mf_last_error()
        raise MeshflowNativeError(code, msg.decode("utf-8", "replace") if msg else "")
