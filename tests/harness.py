"""Shared test plumbing: solver-space LP fixtures and C-ABI call helpers.

Everything the reference's METHOD plugin sees is the argument list of
``solver(m,n,nz,iA,kA,A,b,c,f,x,y,w,z)`` (reference src/common/solve.c:24-26,237), so a
fixture is exactly those arrays.  The same helper drives three implementations through the
same signature: the compiled reference (oracle/_ref/libref_*.so), our C restatement
(oracle/libkkt_oracle.so) and the product (libvbkkt.so) -- only tests may load the first two.
"""
from __future__ import annotations

import ctypes as C
import os
import sys
import tempfile
from dataclasses import dataclass
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
GOLDEN = ROOT / "tests" / "golden"
ORACLE_DIR = ROOT / "oracle"
REF_DIR = ORACLE_DIR / "_ref"

c_int_p = C.POINTER(C.c_int)
c_double_p = C.POINTER(C.c_double)
_libc = C.CDLL(None)
_libc.malloc.restype = C.c_void_p
_libc.malloc.argtypes = [C.c_size_t]
_libc.free.argtypes = [C.c_void_p]
_libc.fflush.argtypes = [C.c_void_p]


@dataclass
class LPData:
    """Solver-space LP:  max c'x + f  s.t.  Ax <= b, x >= 0  (CSC A)."""
    name: str
    m: int
    n: int
    nz: int
    kA: np.ndarray   # int32 [n+1]
    iA: np.ndarray   # int32 [nz]
    A: np.ndarray    # float64 [nz]
    b: np.ndarray    # float64 [m]
    c: np.ndarray    # float64 [n]
    f: float
    extra: dict


def read_dump(path, name="lp") -> LPData:
    """Parse the binary written by oracle/ref_dump_solver.c."""
    raw = Path(path).read_bytes()
    m, n, nz = np.frombuffer(raw, dtype="<i4", count=3, offset=0)
    m, n, nz = int(m), int(n), int(nz)
    off = 12
    f = float(np.frombuffer(raw, dtype="<f8", count=1, offset=off)[0]); off += 8
    kA = np.frombuffer(raw, dtype="<i4", count=n + 1, offset=off).copy(); off += 4 * (n + 1)
    iA = np.frombuffer(raw, dtype="<i4", count=nz, offset=off).copy(); off += 4 * nz
    A = np.frombuffer(raw, dtype="<f8", count=nz, offset=off).copy(); off += 8 * nz
    b = np.frombuffer(raw, dtype="<f8", count=m, offset=off).copy(); off += 8 * m
    c = np.frombuffer(raw, dtype="<f8", count=n, offset=off).copy(); off += 8 * n
    assert off == len(raw), (off, len(raw))
    return LPData(name, m, n, nz, kA, iA, A, b, c, f, {})


def load_fixture(name: str) -> LPData:
    z = np.load(GOLDEN / "netlib" / f"{name}.npz", allow_pickle=False)
    extra = {k: z[k] for k in z.files if k not in ("kA", "iA", "A", "b", "c", "dims", "f")}
    m, n, nz = (int(v) for v in z["dims"])
    return LPData(name, m, n, nz, z["kA"].astype(np.int32), z["iA"].astype(np.int32),
                  z["A"].astype(np.float64), z["b"].astype(np.float64),
                  z["c"].astype(np.float64), float(z["f"]), extra)


def fixture_names():
    d = GOLDEN / "netlib"
    return sorted(p.stem for p in d.glob("*.npz")) if d.exists() else []


def transpose_csc(m, n, kA, iA, A):
    """numpy counting-sort transpose with the entry order of reference atnum
    (src/common/linalg.c:75-103): inside each output column, input columns ascend."""
    nz = int(kA[n])
    cols = np.repeat(np.arange(n, dtype=np.int32), np.diff(kA[: n + 1]))
    order = np.argsort(iA[:nz], kind="stable")
    kAt = np.zeros(m + 1, dtype=np.int32)
    np.cumsum(np.bincount(iA[:nz], minlength=m), out=kAt[1:])
    return kAt, cols[order].astype(np.int32), A[:nz][order].copy()


class capture_stdout:
    """Redirect the process-level fd 1 (C printf) into a temp file."""

    def __init__(self):
        self.text = ""

    def __enter__(self):
        sys.stdout.flush()
        _libc.fflush(None)
        self._tmp = tempfile.TemporaryFile(mode="w+b")
        self._saved = os.dup(1)
        os.dup2(self._tmp.fileno(), 1)
        return self

    def __exit__(self, *exc):
        _libc.fflush(None)
        os.dup2(self._saved, 1)
        os.close(self._saved)
        self._tmp.seek(0)
        self.text = self._tmp.read().decode("latin-1")
        self._tmp.close()
        return False


def ptr_i(a):
    return a.ctypes.data_as(c_int_p)


def ptr_d(a):
    return a.ctypes.data_as(c_double_p)


def call_solver(fn, lp: LPData, capture=True):
    """Call a ``solver``-signature C function on fresh copies of the LP arrays.

    Mirrors the allocation contract of solvelp (reference src/common/solve.c:180-197): A/iA
    have nz+m slots, kA n+m+1, b and c n+m, x and y n+m (calloc), w[m], z[n] malloc-owned
    because the plugin frees them (src/ipo/hsd.c:290-291).  Returns (status, log, x, y).
    """
    m, n, nz = lp.m, lp.n, lp.nz
    kA = np.zeros(n + m + 1, dtype=np.int32); kA[: n + 1] = lp.kA
    iA = np.zeros(nz + m, dtype=np.int32); iA[:nz] = lp.iA
    A = np.zeros(nz + m, dtype=np.float64); A[:nz] = lp.A
    b = np.zeros(n + m, dtype=np.float64); b[:m] = lp.b
    c = np.zeros(n + m, dtype=np.float64); c[:n] = lp.c
    x = np.zeros(n + m, dtype=np.float64)
    y = np.zeros(n + m, dtype=np.float64)
    w = _libc.malloc(8 * max(m, 1))
    z = _libc.malloc(8 * max(n, 1))
    C.memset(w, 0, 8 * max(m, 1))
    C.memset(z, 0, 8 * max(n, 1))
    fn.restype = C.c_int
    fn.argtypes = [C.c_int, C.c_int, C.c_int, c_int_p, c_int_p, c_double_p, c_double_p,
                   c_double_p, C.c_double, c_double_p, c_double_p, C.c_void_p, C.c_void_p]
    if capture:
        with capture_stdout() as cap:
            status = fn(m, n, nz, ptr_i(iA), ptr_i(kA), ptr_d(A), ptr_d(b), ptr_d(c), lp.f,
                        ptr_d(x), ptr_d(y), w, z)
        log = cap.text
    else:
        status = fn(m, n, nz, ptr_i(iA), ptr_i(kA), ptr_d(A), ptr_d(b), ptr_d(c), lp.f,
                    ptr_d(x), ptr_d(y), w, z)
        log = ""
    return int(status), log, x[:n].copy(), y[:m].copy()


def load_ref(method="hsd"):
    """The compiled, unmodified reference (oracle/_ref, built by oracle/Makefile)."""
    path = REF_DIR / f"libref_{method}.so"
    if not path.exists():
        return None
    return C.CDLL(str(path))


def iteration_lines(log: str):
    """The per-iteration lines of a METHOD log (everything after the dashed banner rule)."""
    out, seen = [], False
    for line in log.splitlines():
        if line.startswith("- - - -"):
            seen = True
            continue
        if seen and line[:9].strip().isdigit():
            out.append(line.rstrip())
    return out


# ------------------------------------------------------------------------------------------------
# oracle restatement helpers (oracle/libkkt_oracle.so) -- tests only
# ------------------------------------------------------------------------------------------------
def declare_oracle(lib):
    lib.kko_dotprod.restype = C.c_double
    lib.kko_dotprod.argtypes = [c_double_p, c_double_p, C.c_int]
    lib.kko_maxv.restype = C.c_double
    lib.kko_maxv.argtypes = [c_double_p, C.c_int]
    lib.kko_smx.argtypes = [C.c_int, C.c_int, c_double_p, c_int_p, c_int_p, c_double_p, c_double_p]
    lib.kko_atnum.argtypes = [C.c_int, C.c_int, c_int_p, c_int_p, c_double_p, c_int_p, c_int_p, c_double_p]
    lib.kko_create.restype = C.c_void_p
    lib.kko_destroy.argtypes = [C.c_void_p]
    lib.kko_ldltfac.argtypes = [C.c_void_p, C.c_int, C.c_int, c_int_p, c_int_p, c_double_p, c_double_p,
                                c_double_p, c_int_p, c_int_p, c_double_p]
    lib.kko_forwardbackward.argtypes = [C.c_void_p] + [c_double_p] * 4
    lib.kko_forwardbackward.restype = C.c_int
    lib.kko_rawsolve.argtypes = [C.c_void_p, c_double_p]
    lib.kko_rawsolve.restype = C.c_int
    for name, rt in [("dim", C.c_int), ("denwin", C.c_int), ("pdf", C.c_int), ("ndep", C.c_int),
                     ("epsdiag", C.c_double), ("last_passes", C.c_int), ("perm", c_int_p),
                     ("iperm", c_int_p), ("kAAt", c_int_p), ("iAAt", c_int_p), ("AAt", c_double_p),
                     ("diag", c_double_p), ("mark", c_int_p)]:
        fn = getattr(lib, "kko_" + name)
        fn.restype = rt
        fn.argtypes = [C.c_void_p]
    lib.kko_capture.argtypes = [C.c_int] + [c_double_p] * 6
    lib.kko_last_timing.argtypes = [c_double_p, c_double_p, c_int_p, c_int_p, c_int_p]
    return lib


class OracleFactor:
    """Handle on the oracle's factor object, ldlt-space arguments (mirrors vbkkt.KKT)."""

    def __init__(self, lib, m, n, kA, iA, A, kAt, iAt, At):
        self.lib, self.m, self.n = lib, m, n
        self._keep = [np.ascontiguousarray(v) for v in (kA, iA, A, kAt, iAt, At)]
        self.h = lib.kko_create()

    def factor(self, dn, dm):
        kA, iA, A, kAt, iAt, At = self._keep
        dn = np.ascontiguousarray(dn, dtype=np.float64)
        dm = np.ascontiguousarray(dm, dtype=np.float64)
        self.lib.kko_ldltfac(self.h, self.m, self.n, ptr_i(kA), ptr_i(iA), ptr_d(A), ptr_d(dn), ptr_d(dm),
                             ptr_i(kAt), ptr_i(iAt), ptr_d(At))

    def solve(self, Dn, Dm, dx, dy):
        Dn = np.ascontiguousarray(Dn, dtype=np.float64)
        Dm = np.ascontiguousarray(Dm, dtype=np.float64)
        dx, dy = np.array(dx, dtype=np.float64), np.array(dy, dtype=np.float64)
        ok = self.lib.kko_forwardbackward(self.h, ptr_d(Dn), ptr_d(Dm), ptr_d(dx), ptr_d(dy))
        return dx, dy, int(ok)

    def rawsolve(self, z):
        z = np.array(z, dtype=np.float64)
        self.lib.kko_rawsolve(self.h, ptr_d(z))
        return z

    @property
    def dim(self): return self.m + self.n
    @property
    def lnz(self): return int(self._arr(self.lib.kko_kAAt, self.dim + 1)[self.dim])
    def _arr(self, fn, count):
        return np.ctypeslib.as_array(fn(self.h), (max(count, 1),))[:count].copy()
    @property
    def perm(self): return self._arr(self.lib.kko_perm, self.dim)
    @property
    def kAAt(self): return self._arr(self.lib.kko_kAAt, self.dim + 1)
    @property
    def iAAt(self): return self._arr(self.lib.kko_iAAt, self.lnz)
    @property
    def L(self): return self._arr(self.lib.kko_AAt, self.lnz)
    @property
    def diag(self): return self._arr(self.lib.kko_diag, self.dim)
    @property
    def mark(self): return self._arr(self.lib.kko_mark, self.dim)
    @property
    def ndep(self): return int(self.lib.kko_ndep(self.h))
    @property
    def passes(self): return int(self.lib.kko_last_passes(self.h))

    def close(self):
        if self.h:
            self.lib.kko_destroy(self.h)
            self.h = None


def capture_step(oracle, lp: LPData, method: str, it: int):
    """Run the oracle METHOD and return (E, D, rhs_y, rhs_x, sol_y, sol_x) of iteration `it`."""
    m, n = lp.m, lp.n
    bufs = [np.zeros(m), np.zeros(n), np.zeros(m), np.zeros(n), np.zeros(m), np.zeros(n)]
    oracle.kko_capture(it, *[ptr_d(v) for v in bufs])
    call_solver(getattr(oracle, "kko_solver_" + method), lp)
    oracle.kko_capture(-1, None, None, None, None, None, None)
    return bufs


def solve_via(vbkkt, lib, lp: LPData, method: str, mode=0, profile=False):
    """Product METHOD plugin on a fixture; returns (status, log, x, y, profile)."""
    with capture_stdout() as cap:
        st, x, y, prof = vbkkt.solve_lp(method, lp.m, lp.n, lp.nz, lp.iA, lp.kA, lp.A, lp.b, lp.c, lp.f,
                                        mode=mode, profile=profile, lib=lib)
    return st, cap.text, x, y, prof


def kkt_for(vbkkt, lib, lp: LPData, device=0, mode=0):
    """Factor object for a solver-space LP with the argument swap of hsd.c:218."""
    kAt, iAt, At = transpose_csc(lp.m, lp.n, lp.kA, lp.iA, lp.A)
    k = vbkkt.KKT(device=device, mode=mode, lib=lib)
    k.analyze(lp.n, lp.m, kAt, iAt, At, lp.kA, lp.iA, lp.A)
    return k


def oracle_factor_for(oracle, lp: LPData):
    kAt, iAt, At = transpose_csc(lp.m, lp.n, lp.kA, lp.iA, lp.A)
    return OracleFactor(oracle, lp.n, lp.m, kAt, iAt, At, lp.kA, lp.iA, lp.A)
