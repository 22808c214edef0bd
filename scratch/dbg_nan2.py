import sys, os, ctypes as C, numpy as np, time
sys.path.insert(0, "tests")
import harness as H, parity as P, conftest
vb = conftest._load_pkg(); lib = vb.load()
lp = vb.workloads.random_sparse_lp(0, 2000, 4000)
last = int(sys.argv[1])
Kf = H.kkt_for(vb, lib, lp, mode=vb.MODE_FAST)
def stats(v):
    v = np.asarray(v); f = np.isfinite(v)
    return "finite=%s nan=%d inf=%d absmax=%.3e" % (bool(f.all()), int(np.isnan(v).sum()), int(np.isinf(v).sum()), np.abs(v[f]).max() if f.any() else 0)
for it in range(0, last + 1):
    with H.capture_stdout():
        E, D, ry, rx, sy, sx = vb.capture_iterate("hsd", lp.m, lp.n, lp.nz, lp.iA, lp.kA, lp.A, lp.b, lp.c, lp.f, it, mode=vb.MODE_FAST, lib=lib)
    Kf.factor(E, D)
    if it >= last - 6:
        Lf, df, mf = Kf.get_factor()
        print("iter", it, "epsdiag %.1e ndep %d" % (Kf.epsdiag, Kf.ndep), "diag", stats(df), "L", stats(Lf), "| in-solver sol", stats(sy), stats(sx))
        T = None
        gy, gx, _ = Kf.solve(E, D, ry, rx); print("    replay solve", stats(gy), stats(gx), "passes", Kf.last_passes)
        gb, gc, _ = Kf.solve(E, D, -lp.b, -lp.c); print("    replay solve2", stats(gb), stats(gc), "passes", Kf.last_passes, flush=True)
        z = np.ones(lp.m + lp.n); zz = Kf.rawsolve(z); print("    rawsolve(ones)", stats(zz))
