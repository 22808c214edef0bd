"""Parity checks shared by tests/test_emu.py (host emulation of the kernels, tiny inputs, CPU) and
tests/test_gpu.py (the real thing through the C ABI on a B200).  `lib` is the library under test,
`oracle` the plain-C restatement (checker)."""
import numpy as np

import harness as H


def check_linalg(vbkkt, lib, oracle, sizes, seed=1):
    rng = np.random.default_rng(seed)
    for n in sizes:
        x = rng.standard_normal(n) * 10.0 ** rng.integers(-8, 8, n)
        y = rng.standard_normal(n)
        # bit-exact: strict mode replays the reference's left-to-right sum (linalg.c:22)
        assert vbkkt.dotprod(x, y, lib=lib) == oracle.kko_dotprod(H.ptr_d(x), H.ptr_d(y), n), n
        assert vbkkt.maxv(x, lib=lib) == oracle.kko_maxv(H.ptr_d(x), n), n
    z = np.array([0.0, -0.0, 0.0])
    assert vbkkt.maxv(z, lib=lib) == 0.0


def check_transpose_and_smx(vbkkt, lib, oracle, lp, seed=2):
    rng = np.random.default_rng(seed)
    kat, iat, at = vbkkt.atnum(lp.m, lp.n, lp.kA, lp.iA, lp.A, lib=lib)
    k2, i2, a2 = H.transpose_csc(lp.m, lp.n, lp.kA, lp.iA, lp.A)
    assert np.array_equal(kat, k2) and np.array_equal(iat, i2) and np.array_equal(at, a2)
    x = rng.standard_normal(lp.n)
    y = vbkkt.smx(lp.m, lp.n, lp.A, lp.kA, lp.iA, x, lib=lib)
    y2 = np.zeros(lp.m)
    oracle.kko_smx(lp.m, lp.n, H.ptr_d(lp.A), H.ptr_i(lp.kA), H.ptr_i(lp.iA), H.ptr_d(x), H.ptr_d(y2))
    assert np.array_equal(y, y2)
    # transpose of the transpose is the original (size-independent property)
    kb, ib, ab = vbkkt.atnum(lp.n, lp.m, kat, iat, at, lib=lib)
    assert np.array_equal(kb, lp.kA[: lp.n + 1]) and np.array_equal(ab, lp.A[: lp.nz])


def check_ragged_transpose(vbkkt, lib):
    # empty column, empty row, a dense row
    kA = np.array([0, 2, 2, 3, 5], dtype=np.int32)
    iA = np.array([0, 3, 1, 0, 3], dtype=np.int32)
    A = np.array([1.0, 2.0, 3.0, 4.0, 5.0])
    kat, iat, at = vbkkt.atnum(5, 4, kA, iA, A, lib=lib)
    k2, i2, a2 = H.transpose_csc(5, 4, kA, iA, A)
    assert np.array_equal(kat, k2) and np.array_equal(iat, i2) and np.array_equal(at, a2)
    y = vbkkt.smx(5, 4, A, kA, iA, np.array([1.0, 10.0, 100.0, 1000.0]), lib=lib)
    assert np.array_equal(y, np.array([4001.0, 300.0, 0.0, 5002.0, 0.0]))


def check_kkt_step(vbkkt, lib, oracle, lp, method, it, refine_rhs=True):
    """One KKT step on inputs captured from iteration `it` of the oracle's run: the numeric factor
    (L, diag, mark, ndep) and the refined solution must be BIT-EQUAL to the oracle's."""
    E, D, ry, rx, sy, sx = H.capture_step(oracle, lp, method, it)
    F = H.oracle_factor_for(oracle, lp)
    K = H.kkt_for(vbkkt, lib, lp)
    try:
        F.factor(E, D)
        K.factor(E, D)
        L, d, mk = K.get_factor()
        assert np.array_equal(mk, F.mark)
        assert np.array_equal(d, F.diag)
        assert np.array_equal(L, F.L)
        assert K.ndep == F.ndep
        ndep1 = F.ndep
        oy, ox, _ = F.solve(E, D, ry, rx)
        gy, gx, _ = K.solve(E, D, ry, rx)
        oy_passes = F.passes
        assert K.last_passes == F.passes
        assert np.array_equal(gy, oy) and np.array_equal(gx, ox)
        z = np.random.default_rng(it).standard_normal(lp.m + lp.n)
        assert np.array_equal(K.rawsolve(z), F.rawsolve(z))
        # both systems of an hsd iteration (hsd.c:223,228) in one pair of sweeps: each right-hand side must come out
        # exactly as from its own forwardbackward call, with its own number of refinement passes -- in either slot
        rng = np.random.default_rng(1000 + it)
        qy, qx = -lp.b * (1 + 1e-3 * rng.standard_normal(lp.m)), -lp.c * (1 + 1e-3 * rng.standard_normal(lp.n))
        o2y, o2x, _ = F.solve(E, D, qy, qx)
        p2 = F.passes
        for (a, b) in (((ry, rx), (qy, qx)), ((qy, qx), (ry, rx))):
            r0, r1 = K.solve2(E, D, a[0], a[1], b[0], b[1])
            want0, want1 = ((oy, ox), (o2y, o2x)) if a[0] is ry else ((o2y, o2x), (oy, ox))
            assert np.array_equal(r0[0], want0[0]) and np.array_equal(r0[1], want0[1])
            assert np.array_equal(r1[0], want1[0]) and np.array_equal(r1[1], want1[1])
            passes = (K.last_passes2(0), K.last_passes2(1))
            assert passes == ((oy_passes, p2) if a[0] is ry else (p2, oy_passes))
        # a second factorisation on the same handles: state carried across calls (epsdiag
        # escalation ldlt.c:293-306, mark reset ldlt.c:280) must stay in lock step
        F.factor(E * 0.5, D * 2.0)
        K.factor(E * 0.5, D * 2.0)
        L, d, mk = K.get_factor()
        assert np.array_equal(mk, F.mark) and np.array_equal(d, F.diag) and np.array_equal(L, F.L)
        assert K.epsdiag == float(oracle.kko_epsdiag(F.h))
        return dict(ndep=ndep1, passes=oy_passes, lnz=K.lnz)
    finally:
        F.close()
        K.close()


def _rel(a, b):
    return float(np.max(np.abs(a - b)) / max(float(np.max(np.abs(b))), 1e-300))


def check_kkt_step_fast(vbkkt, lib, oracle, lp, method, it, tol=1e-6):
    """FAST mode (re-associated sums, dense-window factorisation) on a captured iterate WITHOUT
    dependent pivots: the factor and the refined solution agree with the oracle to rounding, and the
    residual of K z = rhs is as small as the reference's own refinement criterion asks."""
    import scipy.sparse as sp
    E, D, ry, rx, sy, sx = H.capture_step(oracle, lp, method, it)
    F = H.oracle_factor_for(oracle, lp)
    K = H.kkt_for(vbkkt, lib, lp, mode=vbkkt.MODE_FAST)
    try:
        F.factor(E, D)
        K.factor(E, D)
        assert F.ndep == 0, "pick an iterate without dependent pivots for the tolerance check"
        L, d, mk = K.get_factor()
        assert np.array_equal(mk, F.mark)
        assert _rel(d, F.diag) < tol and _rel(L, F.L) < tol
        oy, ox, _ = F.solve(E, D, ry, rx)
        gy, gx, _ = K.solve(E, D, ry, rx)
        assert _rel(gy, oy) < tol and _rel(gx, ox) < tol
        A = sp.csc_matrix((lp.A, lp.iA, lp.kA), shape=(lp.m, lp.n))
        r1 = -E * gy + A @ gx - ry            # [-E A; A^T D][y;x] = [ry;rx]  
        r2 = A.T @ gy + D * gx - rx
        scale = max(np.abs(ry).max(), np.abs(rx).max()) + 1
        assert max(np.abs(r1).max(), np.abs(r2).max()) <= 1e-8 * scale
        z = np.random.default_rng(it).standard_normal(lp.m + lp.n)
        assert _rel(K.rawsolve(z), F.rawsolve(z)) < 1e-5
        return dict(rel_y=_rel(gy, oy), rel_x=_rel(gx, ox))
    finally:
        F.close()
        K.close()


def check_full_solve_fast(vbkkt, lib, lp, method="hsd", iter_slack=1):
    """FAST mode end to end on a problem of the robust list: the north_star tolerances
    (status, iteration count +-1, objective 1e-8 relative, infeasibilities 1e-7)."""
    st, log, x, y, _ = H.solve_via(vbkkt, lib, lp, method, mode=vbkkt.MODE_FAST)
    north_star_tolerances(lp, method, x, y, st, log, iter_slack=iter_slack)
    return st


def check_full_solve(vbkkt, lib, lp, method, want_bits=True):
    """The device-resident METHOD plugin against the golden fixture: same status, byte-identical
    iteration log (the reference's golden log for hsd), bit-equal x and y."""
    st, log, x, y, _ = H.solve_via(vbkkt, lib, lp, method)
    assert st == int(lp.extra[method + "_status"])
    exp = str(lp.extra[method + "_log"])
    if log != exp:
        a, b = log.splitlines(), exp.splitlines()
        for i, (u, v) in enumerate(zip(a, b)):
            assert u == v, f"{lp.name} {method}: first differing log line {i}:\n got {u!r}\n exp {v!r}"
        assert len(a) == len(b), f"{lp.name} {method}: {len(a)} lines, expected {len(b)}"
    if want_bits:
        assert np.array_equal(x, lp.extra[method + "_x"])
        assert np.array_equal(y, lp.extra[method + "_y"])
    return st


def check_hsdls(vbkkt, lib, lp):
    """METHOD = hsdls (reference src/ipo/hsdls.c:37-336) against the committed output of the compiled reference
    (tests/golden/hsdls/, made by tests/golden/make_hsdls.py): same status, byte-identical log, bit-equal x and y."""
    z = np.load(H.GOLDEN / "hsdls" / f"{lp.name}.npz")
    st, log, x, y, _ = H.solve_via(vbkkt, lib, lp, "hsdls")
    assert st == int(z["status"])
    exp = str(z["log"])
    a, b = log.splitlines(), exp.splitlines()
    for i, (u, v) in enumerate(zip(a, b)):
        assert u == v, f"{lp.name} hsdls: first differing log line {i}:\n got {u!r}\n exp {v!r}"
    assert len(a) == len(b), f"{lp.name} hsdls: {len(a)} lines, expected {len(b)}"
    assert np.array_equal(x, z["x"]) and np.array_equal(y, z["y"])
    return st


def north_star_tolerances(lp, method, x, y, status, log, iter_slack=1):
    """The tolerance form of parity (BASELINE.json north_star): objective 1e-8 relative,
    infeasibilities 1e-7, iteration count +-1, same status."""
    xr, yr = lp.extra[method + "_x"], lp.extra[method + "_y"]
    assert status == int(lp.extra[method + "_status"])
    obj, obj_r = float(lp.c @ x), float(lp.c @ xr)
    assert abs(obj - obj_r) <= 1e-8 * max(1.0, abs(obj_r))
    dobj, dobj_r = float(lp.b @ y), float(lp.b @ yr)
    assert abs(dobj - dobj_r) <= 1e-8 * max(1.0, abs(dobj_r))
    n_it = len(H.iteration_lines(log))
    n_ref = len(H.iteration_lines(str(lp.extra[method + "_log"])))
    assert abs(n_it - n_ref) <= iter_slack, (n_it, n_ref)
    import scipy.sparse as sp
    A = sp.csc_matrix((lp.A, lp.iA, lp.kA), shape=(lp.m, lp.n))
    pinf = np.linalg.norm(np.maximum(A @ x - lp.b, 0.0)), np.linalg.norm(np.maximum(A @ xr - lp.b, 0.0))
    dinf = np.linalg.norm(np.maximum(lp.c - A.T @ y, 0.0)), np.linalg.norm(np.maximum(lp.c - A.T @ yr, 0.0))
    assert abs(pinf[0] - pinf[1]) <= 1e-7 * max(1.0, np.linalg.norm(lp.b))
    assert abs(dinf[0] - dinf[1]) <= 1e-7 * max(1.0, np.linalg.norm(lp.c))


def check_fast_sparse_columns(vbkkt, lib, oracle, lp, method, it):
    """Fast mode keeps the reference's arithmetic for the sparse columns j < T: L, diag and mark of those
    columns equal the oracle's factor bit for bit (level-scheduled kernels of vbk_sparse_level.cuh, or the task
    kernel -- whichever the handle runs)."""
    E, D, *_ = H.capture_step(oracle, lp, method, it)
    F = H.oracle_factor_for(oracle, lp)
    K = H.kkt_for(vbkkt, lib, lp, mode=vbkkt.MODE_FAST)
    try:
        F.factor(E, D)
        out = None
        for _ in range(3):                   # the first two factorisations time the two paths, the third runs the choice
            K.factor(E, D)
            L, d, mk = K.get_factor()
            T = K.dim - K.window
            nsp = int(K.kAAt[T])
            assert np.array_equal(L[:nsp], F.L[:nsp])
            assert np.array_equal(d[:T], F.diag[:T]) and np.array_equal(mk[:T], F.mark[:T])
            out = dict(T=T, nsp=nsp, ndep=F.ndep)
        return out
    finally:
        F.close()
        K.close()


def check_b1_speculation(lib, oracle, lp, rounds=4):
    """The reference's call sequence on the B1 symbols -- ldltfac, forwardbackward(f), forwardbackward(g) with g = (-b, -c)
    every round (hsd.c:218-228) -- against the oracle, call by call, bit for bit; round 2 asks for a different g than the one
    the library remembered, so a stale speculative answer would show."""
    import ctypes as C
    kAt, iAt, At = H.transpose_csc(lp.m, lp.n, lp.kA, lp.iA, lp.A)
    lib.ldltfac.argtypes = [C.c_int, C.c_int, H.c_int_p, H.c_int_p, H.c_double_p, H.c_double_p, H.c_double_p,
                            H.c_int_p, H.c_int_p, H.c_double_p, C.c_int]
    lib.ldltfac.restype = None
    lib.forwardbackward.argtypes = [H.c_double_p] * 4
    lib.forwardbackward.restype = None
    lib.inv_clo.restype = None
    lib.inv_clo()
    F = H.oracle_factor_for(oracle, lp)
    rng = np.random.default_rng(5)
    gy0, gx0 = -lp.b.copy(), -lp.c.copy()
    try:
        for it in range(rounds):
            E = 10.0 ** rng.uniform(-2, 2, lp.m)
            D = 10.0 ** rng.uniform(-2, 2, lp.n)
            lib.ldltfac(lp.n, lp.m, H.ptr_i(kAt), H.ptr_i(iAt), H.ptr_d(At), H.ptr_d(E), H.ptr_d(D),
                        H.ptr_i(lp.kA), H.ptr_i(lp.iA), H.ptr_d(lp.A), 0)
            F.factor(E, D)
            fy, fx = rng.standard_normal(lp.m), rng.standard_normal(lp.n)
            gy, gx = (gy0.copy(), gx0.copy()) if it != 2 else (gy0 * 1.5, gx0 * 0.5)
            oy, ox, _ = F.solve(E, D, fy, fx)
            o2y, o2x, _ = F.solve(E, D, gy, gx)
            lib.forwardbackward(H.ptr_d(E), H.ptr_d(D), H.ptr_d(fy), H.ptr_d(fx))
            lib.forwardbackward(H.ptr_d(E), H.ptr_d(D), H.ptr_d(gy), H.ptr_d(gx))
            assert np.array_equal(fy, oy) and np.array_equal(fx, ox), it
            assert np.array_equal(gy, o2y) and np.array_equal(gx, o2x), it
    finally:
        F.close()
        lib.inv_clo()

