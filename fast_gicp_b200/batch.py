"""ctypes binding of lib/libvgicp_batch_b200.so (include/vgicp_batch_b200.h): many registrations in one C call -- a pool of vgicp
handles on one device, worker threads pulling pairs from a shared counter; each pair is one vgicp_register (the body of the
reference's benchmark loop, src/align.cpp:72-81).  Requires the built libraries (no CPU fallback)."""
import ctypes as C
import os

import numpy as np

from . import core as _core

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libvgicp_batch_b200.so")
EXPORTED_SYMBOLS = ["vgicp_batch_create", "vgicp_batch_destroy", "vgicp_batch_last_error", "vgicp_batch_num_streams", "vgicp_batch_configure", "vgicp_batch_register"]
_lib = None


def load_library():
    global _lib
    if _lib is None:
        _core.load_library()  # libvgicp_b200.so first (the batch library links against it)
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run `python build_native.py`")
        L = C.CDLL(LIB_PATH)
        vp = C.c_void_p
        L.vgicp_batch_create.argtypes = [C.c_int, C.c_int, C.POINTER(vp)]
        L.vgicp_batch_destroy.argtypes = [vp]
        L.vgicp_batch_destroy.restype = None
        L.vgicp_batch_last_error.argtypes = [vp]
        L.vgicp_batch_last_error.restype = C.c_char_p
        L.vgicp_batch_num_streams.argtypes = [vp]
        L.vgicp_batch_configure.argtypes = [vp, C.c_double, C.c_int, C.c_double]
        L.vgicp_batch_register.argtypes = [vp, C.c_size_t, C.POINTER(vp), C.POINTER(C.c_size_t), C.POINTER(vp), C.POINTER(C.c_size_t), C.c_size_t, C.c_int, C.c_int, C.c_int,
                                           C.POINTER(C.c_double), C.c_void_p, C.POINTER(_core.AlignResult), C.POINTER(vp)]
        _lib = L
    return _lib


class BatchRegistration:
    """A pool of `n_streams` registration contexts on one device."""

    def __init__(self, device=0, n_streams=16):
        self._lib = load_library()
        self._b = C.c_void_p()
        rc = self._lib.vgicp_batch_create(int(device), int(n_streams), C.byref(self._b))
        if rc:
            self._b = None
            raise _core.VgicpError(rc, "vgicp_batch_create failed")

    def close(self):
        if getattr(self, "_b", None):
            self._lib.vgicp_batch_destroy(self._b)
            self._b = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc:
            raise _core.VgicpError(rc, self._lib.vgicp_batch_last_error(self._b).decode())

    def configure(self, resolution=1.0, method="DIRECT1", radius=-1.0):
        self._check(self._lib.vgicp_batch_configure(self._b, float(resolution), int(_core.NEIGHBOR_SEARCH[method] if isinstance(method, str) else method), float(radius)))

    def prepare(self, targets, sources, aligned_out=None):
        """Build the pointer tables of a batch once (arrays must stay alive and C-contiguous float32; aligned_out: optional list of
        (n_i, 3) float32 host buffers, e.g. pinned).  Returns a callable running the batch: call() -> (first status, results)."""
        n = len(targets)
        assert len(sources) == n and (aligned_out is None or len(aligned_out) == n)
        for a in list(targets) + list(sources) + list(aligned_out or []):
            assert a.dtype == np.float32 and a.flags.c_contiguous and a.ndim == 2 and a.shape[1] == 3
        vp = C.c_void_p
        tp = (vp * n)(*[a.ctypes.data for a in targets])
        sp = (vp * n)(*[a.ctypes.data for a in sources])
        nt = (C.c_size_t * n)(*[len(a) for a in targets])
        ns = (C.c_size_t * n)(*[len(a) for a in sources])
        ap = (vp * n)(*[a.ctypes.data for a in aligned_out]) if aligned_out is not None else None
        res = (_core.AlignResult * n)()
        keep = (targets, sources, aligned_out)

        def call(k=20, reg=_core.REG_PLANE):
            _ = keep
            self._check(self._lib.vgicp_batch_register(self._b, n, tp, nt, sp, ns, 12, 0, int(k), int(reg), None, None, res, ap))
            return res

        return call

    def register(self, targets, sources, k=20, reg=_core.REG_PLANE, guesses=None, params=None, want_aligned=False):
        """targets / sources: sequences of (n_i, 3) float32 C-contiguous host arrays.  Returns (list of 4x4 poses, list of
        AlignResult, list of aligned clouds or None)."""
        n = len(targets)
        assert len(sources) == n
        t = [np.ascontiguousarray(a, dtype=np.float32) for a in targets]
        s = [np.ascontiguousarray(a, dtype=np.float32) for a in sources]
        vp = C.c_void_p
        tp = (vp * n)(*[a.ctypes.data for a in t])
        sp = (vp * n)(*[a.ctypes.data for a in s])
        nt = (C.c_size_t * n)(*[len(a) for a in t])
        ns = (C.c_size_t * n)(*[len(a) for a in s])
        res = (_core.AlignResult * n)()
        g = None
        if guesses is not None:
            flat = np.concatenate([np.asarray(_core.pose_to_c(G), dtype=np.float64).ravel() for G in guesses])
            g = flat.ctypes.data_as(C.POINTER(C.c_double))
        aligned, ap = None, None
        if want_aligned:
            aligned = [np.empty((len(a), 3), dtype=np.float32) for a in s]
            ap = (vp * n)(*[a.ctypes.data for a in aligned])
        self._check(self._lib.vgicp_batch_register(self._b, n, tp, nt, sp, ns, 12, 0, int(k), int(reg), g, C.byref(params) if params is not None else None, res, ap))
        return [_core.pose_from_c(r.T) for r in res], list(res), aligned
