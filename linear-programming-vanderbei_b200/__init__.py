"""vbkkt -- Python host side over the C ABI of libvbkkt.so (include/vbkkt.h).

The directory name carries a hyphen, so load it with `importlib` (see `__graft_entry__.py` /
`tests/conftest.py`, which register it as module ``vbkkt``).  The functions mirror the reference's
plugin interface for the hot path -- ``ldltfac`` / ``forwardbackward`` (src/ipo/ldlt.h),
``smx`` / ``atnum`` / ``dotprod`` / ``maxv`` (src/common/linalg.h), ``solver`` for METHOD=hsd and
METHOD=intpt -- same argument meaning, numpy arrays in place of raw pointers.

There is no CPU implementation behind this module: if the CUDA library is missing the import of the
native part raises, and on a box without a GPU every numeric call terminates the process with the
library's "no CUDA device" message (only the host-side symbolic analysis works without a GPU).
"""
from __future__ import annotations

import ctypes as C
import importlib.util
import os
import sys
from pathlib import Path

import numpy as np

PKG_DIR = Path(__file__).resolve().parent
ROOT = PKG_DIR.parent
LIB_PATH = PKG_DIR / "libvbkkt.so"

MODE_STRICT = 0
MODE_FAST = 1

_ip = C.POINTER(C.c_int)
_dp = C.POINTER(C.c_double)


class Profile(C.Structure):
    _fields_ = [("total_s", C.c_double), ("setup_s", C.c_double), ("factor_s", C.c_double),
                ("solve_s", C.c_double), ("factor_calls", C.c_longlong), ("solve_calls", C.c_longlong),
                ("rawsolve_calls", C.c_longlong), ("kernel_launches", C.c_longlong),
                ("refine_passes", C.c_longlong), ("iterations", C.c_int), ("N", C.c_int),
                ("lnz", C.c_longlong), ("narth", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


def _i(a):
    return a.ctypes.data_as(_ip)


def _d(a):
    return a.ctypes.data_as(_dp)


def _ai(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _ad(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _declare(lib):
    lib.vbk_version.restype = C.c_char_p
    lib.vbk_device_count.restype = C.c_int
    lib.vbk_get_mode.restype = C.c_int
    lib.vbk_set_mode.argtypes = [C.c_int]
    lib.vbk_set_device.argtypes = [C.c_int]
    lib.ldltfac.argtypes = [C.c_int, C.c_int, _ip, _ip, _dp, _dp, _dp, _ip, _ip, _dp, C.c_int]
    lib.ldltfac.restype = None
    lib.forwardbackward.argtypes = [_dp, _dp, _dp, _dp]
    lib.forwardbackward.restype = None
    lib.inv_clo.restype = None
    lib.dotprod.argtypes = [_dp, _dp, C.c_int]
    lib.dotprod.restype = C.c_double
    lib.maxv.argtypes = [_dp, C.c_int]
    lib.maxv.restype = C.c_double
    lib.smx.argtypes = [C.c_int, C.c_int, _dp, _ip, _ip, _dp, _dp]
    lib.smx.restype = None
    lib.atnum.argtypes = [C.c_int, C.c_int, _ip, _ip, _dp, _ip, _ip, _dp]
    lib.atnum.restype = None
    lib.vbk_solve_lp.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _ip, _ip, _dp, _dp,
                                 _dp, C.c_double, _dp, _dp, C.POINTER(Profile)]
    lib.vbk_solve_lp.restype = C.c_int
    lib.vbk_kkt_create.argtypes = [C.c_int, C.c_int]
    lib.vbk_kkt_create.restype = C.c_void_p
    lib.vbk_kkt_destroy.argtypes = [C.c_void_p]
    lib.vbk_kkt_analyze.argtypes = [C.c_void_p, C.c_int, C.c_int, _ip, _ip, _dp, _ip, _ip, _dp]
    lib.vbk_kkt_factor.argtypes = [C.c_void_p, _dp, _dp]
    lib.vbk_kkt_solve.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp]
    lib.vbk_kkt_solve.restype = C.c_int
    lib.vbk_kkt_factor_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.vbk_kkt_solve_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.vbk_kkt_solve_dev.restype = C.c_int
    lib.vbk_kkt_solve2_dev.argtypes = [C.c_void_p] * 7
    lib.vbk_kkt_solve2_dev.restype = C.c_int
    lib.vbk_kkt_solve2.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp, _dp, _dp]
    lib.vbk_kkt_solve2.restype = C.c_int
    lib.vbk_kkt_last_passes2.argtypes = [C.c_void_p, C.c_int]
    lib.vbk_kkt_last_passes2.restype = C.c_int
    lib.vbk_kkt_rawsolve.argtypes = [C.c_void_p, _dp]
    lib.vbk_kkt_rawsolve.restype = C.c_int
    lib.vbk_kkt_sync.argtypes = [C.c_void_p]
    lib.vbk_kkt_stream.argtypes = [C.c_void_p]
    lib.vbk_kkt_stream.restype = C.c_void_p
    for name, rt in [("dim", C.c_int), ("lnz", C.c_longlong), ("denwin", C.c_int), ("pdf", C.c_int), ("window", C.c_int),
                     ("narth", C.c_double), ("nlevels", C.c_int), ("nsupernodes", C.c_int),
                     ("perm", _ip), ("iperm", _ip), ("kAAt", _ip), ("iAAt", _ip),
                     ("epsdiag", C.c_double), ("ndep", C.c_int), ("last_passes", C.c_int),
                     ("launches", C.c_longlong)]:
        fn = getattr(lib, "vbk_kkt_" + name)
        fn.argtypes = [C.c_void_p]
        fn.restype = rt
    lib.vbk_kkt_get_factor.argtypes = [C.c_void_p, _dp, _dp, _ip]
    lib.vbk_kkt_last_factor_kernel_ms.argtypes = [C.c_void_p]
    lib.vbk_kkt_last_factor_kernel_ms.restype = C.c_float
    lib.vbk_set_iteration_limit.argtypes = [C.c_int]
    lib.vbk_measure_fp64_tflops.argtypes = [C.c_int]
    lib.vbk_measure_fp64_tflops.restype = C.c_double
    lib.vbk_measure_hbm_gbs.argtypes = [C.c_int]
    lib.vbk_measure_hbm_gbs.restype = C.c_double
    lib.vbk_capture.argtypes = [C.c_int] + [_dp] * 6
    return lib


_LIBS = {}


def _submodule(name):
    """batch / rowblock / workloads live beside this file; the directory name has a hyphen, so they are
    loaded by path and registered as ``<this module>.<name>``."""
    full = f"{__name__}.{name}"
    if full in sys.modules:
        return sys.modules[full]
    spec = importlib.util.spec_from_file_location(full, PKG_DIR / f"{name}.py")
    mod = importlib.util.module_from_spec(spec)
    sys.modules[full] = mod
    spec.loader.exec_module(mod)
    return mod


def __getattr__(name):
    if name in ("batch", "rowblock", "workloads"):
        return _submodule(name)
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")


def load(path=None):
    """dlopen the native library (default: the in-tree nvcc build).  Raises if it is missing --
    there is deliberately no fallback."""
    path = Path(path) if path else LIB_PATH
    key = str(path)
    if key not in _LIBS:
        if not path.exists():
            raise ImportError(f"{path} not built: run `python __graft_entry__.py build` "
                              "(nvcc, sm_100a); vbkkt has no CPU implementation")
        _LIBS[key] = _declare(C.CDLL(key))
    return _LIBS[key]


def device_count(lib=None) -> int:
    return int((lib or load()).vbk_device_count())


# ------------------------------------------------------------------------------------------------
# reference plugin interface (B1), numpy in / numpy out
# ------------------------------------------------------------------------------------------------
def dotprod(x, y, lib=None):
    """reference linalg.c:17-25"""
    x, y = _ad(x), _ad(y)
    return float((lib or load()).dotprod(_d(x), _d(y), len(x)))


def maxv(x, lib=None):
    """reference linalg.c:108-116"""
    x = _ad(x)
    return float((lib or load()).maxv(_d(x), len(x)))


def smx(m, n, a, ka, ia, x, lib=None):
    """y = A x for CSC (a, ka, ia); reference linalg.c:62-70"""
    a, ka, ia, x = _ad(a), _ai(ka), _ai(ia), _ad(x)
    y = np.zeros(m, dtype=np.float64)
    (lib or load()).smx(m, n, _d(a), _i(ka), _i(ia), _d(x), _d(y))
    return y


def atnum(m, n, ka, ia, a, lib=None):
    """CSC transpose; reference linalg.c:75-103.  Returns (kat, iat, at)."""
    a, ka, ia = _ad(a), _ai(ka), _ai(ia)
    nz = int(ka[n])
    kat = np.zeros(m + 1, dtype=np.int32)
    iat = np.zeros(max(nz, 1), dtype=np.int32)
    at = np.zeros(max(nz, 1), dtype=np.float64)
    (lib or load()).atnum(m, n, _i(ka), _i(ia), _d(a), _i(kat), _i(iat), _d(at))
    return kat, iat[:nz], at[:nz]


class KKT:
    """Handle-based factor object: symbolic once, numeric factor + solves on the GPU.

    Arguments are ldlt-space, as in the reference's ``ldltfac`` (src/ipo/ldlt.h:1-13): ``A`` is an
    m x n CSC matrix, ``At`` its transpose, ``dn`` (length n) feeds the negative block and ``dm``
    (length m) the positive block of K.  ``device=-1`` gives a host-only handle for the symbolic phase.
    """

    def __init__(self, device=0, mode=MODE_STRICT, lib=None):
        self.lib = lib or load()
        self.h = self.lib.vbk_kkt_create(device, mode)
        self._keep = None

    def close(self):
        if self.h:
            self.lib.vbk_kkt_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def analyze(self, m, n, kA, iA, A, kAt, iAt, At):
        arrs = (_ai(kA), _ai(iA), _ad(A), _ai(kAt), _ai(iAt), _ad(At))
        self._keep = arrs
        self.m, self.n = m, n
        self.lib.vbk_kkt_analyze(self.h, m, n, _i(arrs[0]), _i(arrs[1]), _d(arrs[2]),
                                 _i(arrs[3]), _i(arrs[4]), _d(arrs[5]))
        return self

    # symbolic results
    @property
    def dim(self): return int(self.lib.vbk_kkt_dim(self.h))
    @property
    def lnz(self): return int(self.lib.vbk_kkt_lnz(self.h))
    @property
    def denwin(self): return int(self.lib.vbk_kkt_denwin(self.h))
    @property
    def pdf(self): return int(self.lib.vbk_kkt_pdf(self.h))
    @property
    def window(self): return int(self.lib.vbk_kkt_window(self.h))
    @property
    def narth(self): return float(self.lib.vbk_kkt_narth(self.h))
    @property
    def nlevels(self): return int(self.lib.vbk_kkt_nlevels(self.h))
    @property
    def nsupernodes(self): return int(self.lib.vbk_kkt_nsupernodes(self.h))

    def _iarr(self, fn, count):
        p = fn(self.h)
        return np.ctypeslib.as_array(p, (max(count, 1),))[:count].copy()

    @property
    def perm(self): return self._iarr(self.lib.vbk_kkt_perm, self.dim)
    @property
    def iperm(self): return self._iarr(self.lib.vbk_kkt_iperm, self.dim)
    @property
    def kAAt(self): return self._iarr(self.lib.vbk_kkt_kAAt, self.dim + 1)
    @property
    def iAAt(self): return self._iarr(self.lib.vbk_kkt_iAAt, self.lnz)

    # numeric
    def factor(self, dn, dm):
        dn, dm = _ad(dn), _ad(dm)
        self.lib.vbk_kkt_factor(self.h, _d(dn), _d(dm))

    def solve(self, Dn, Dm, dx, dy):
        """Returns (dx_out, dy_out, consistent); dx has length n, dy length m (ldlt-space)."""
        Dn, Dm = _ad(Dn), _ad(Dm)
        dx, dy = _ad(dx).copy(), _ad(dy).copy()
        ok = self.lib.vbk_kkt_solve(self.h, _d(Dn), _d(Dm), _d(dx), _d(dy))
        return dx, dy, int(ok)

    def solve2(self, Dn, Dm, dx0, dy0, dx1, dy1):
        """Two right-hand sides on the same factor (the two forwardbackward calls of an hsd iteration, hsd.c:223,228)
        in one pair of sweeps per refinement pass.  Returns ((dx0, dy0, consistent0), (dx1, dy1, consistent1))."""
        Dn, Dm = _ad(Dn), _ad(Dm)
        dx0, dy0, dx1, dy1 = (_ad(v).copy() for v in (dx0, dy0, dx1, dy1))
        ok = int(self.lib.vbk_kkt_solve2(self.h, _d(Dn), _d(Dm), _d(dx0), _d(dy0), _d(dx1), _d(dy1)))
        return (dx0, dy0, ok & 1), (dx1, dy1, (ok >> 1) & 1)

    def rawsolve(self, zperm):
        z = _ad(zperm).copy()
        self.lib.vbk_kkt_rawsolve(self.h, _d(z))
        return z

    def get_factor(self):
        L = np.zeros(max(self.lnz, 1), dtype=np.float64)
        diag = np.zeros(self.dim, dtype=np.float64)
        mark = np.zeros(self.dim, dtype=np.int32)
        self.lib.vbk_kkt_get_factor(self.h, _d(L), _d(diag), _i(mark))
        return L[: self.lnz], diag, mark

    @property
    def epsdiag(self): return float(self.lib.vbk_kkt_epsdiag(self.h))
    @property
    def ndep(self): return int(self.lib.vbk_kkt_ndep(self.h))
    @property
    def last_passes(self): return int(self.lib.vbk_kkt_last_passes(self.h))
    @property
    def launches(self): return int(self.lib.vbk_kkt_launches(self.h))

    # device-pointer interface (raw addresses, e.g. torch tensor .data_ptr())
    def factor_dev(self, dn_ptr, dm_ptr):
        self.lib.vbk_kkt_factor_dev(self.h, dn_ptr, dm_ptr)

    def solve_dev(self, Dn_ptr, Dm_ptr, dx_ptr, dy_ptr):
        return int(self.lib.vbk_kkt_solve_dev(self.h, Dn_ptr, Dm_ptr, dx_ptr, dy_ptr))

    def solve2_dev(self, Dn_ptr, Dm_ptr, dx0_ptr, dy0_ptr, dx1_ptr, dy1_ptr):
        """both systems of an hsd iteration (hsd.c:223,228) in one pair of sweeps per refinement pass"""
        return int(self.lib.vbk_kkt_solve2_dev(self.h, Dn_ptr, Dm_ptr, dx0_ptr, dy0_ptr, dx1_ptr, dy1_ptr))

    def last_passes2(self, rhs):
        return int(self.lib.vbk_kkt_last_passes2(self.h, rhs))

    def sync(self):
        self.lib.vbk_kkt_sync(self.h)

    @property
    def stream(self):
        return self.lib.vbk_kkt_stream(self.h)


def capture_iterate(method, m, n, nz, iA, kA, A, b, c, f, it, device=0, mode=MODE_STRICT, lib=None):
    """Run the device-resident METHOD up to iteration `it` and return the KKT-step inputs and outputs
    of that iteration: (E[m], D[n], rhs_y[m], rhs_x[n], sol_y[m], sol_x[n]).  Used by the bench to get a
    realistic interior-point iterate as input for the timed KKT steps."""
    lib = lib or load()
    bufs = [np.zeros(m), np.zeros(n), np.zeros(m), np.zeros(n), np.zeros(m), np.zeros(n)]
    lib.vbk_capture(it, *[_d(v) for v in bufs])
    lib.vbk_set_iteration_limit(it + 1)
    try:
        solve_lp(method, m, n, nz, iA, kA, A, b, c, f, device=device, mode=mode, lib=lib)
    finally:
        lib.vbk_set_iteration_limit(0)
        lib.vbk_capture(-1, None, None, None, None, None, None)
    return bufs


def solve_lp(method, m, n, nz, iA, kA, A, b, c, f=0.0, device=0, mode=MODE_STRICT, profile=False, lib=None):
    """METHOD plugin on explicit device/mode: ``method`` is "hsd" (reference src/ipo/hsd.c:27) or
    "intpt" (src/ipo/intpt.c:33) or "hsdls" (src/ipo/hsdls.c:37).  Prints the reference's iteration log to stdout.
    Returns (status, x[n], y[m], profile-dict-or-None)."""
    lib = lib or load()
    iA, kA, A, b, c = _ai(iA), _ai(kA), _ad(A), _ad(b), _ad(c)
    x = np.zeros(n + m, dtype=np.float64)
    y = np.zeros(n + m, dtype=np.float64)
    prof = Profile() if profile else None
    st = lib.vbk_solve_lp({"hsd": 0, "intpt": 1, "hsdls": 2}[method], device, mode, m, n, nz, _i(iA), _i(kA), _d(A),
                          _d(b), _d(c), float(f), _d(x), _d(y), C.byref(prof) if profile else None)
    return int(st), x[:n].copy(), y[:m].copy(), (prof.as_dict() if profile else None)
