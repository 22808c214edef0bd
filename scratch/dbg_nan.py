import sys, os, ctypes as C, numpy as np, time
sys.path.insert(0, "tests")
import harness as H, parity as P, conftest
vb = conftest._load_pkg(); lib = vb.load()
m, n = 2000, 4000
lp = vb.workloads.random_sparse_lp(0, m, n)
its = [int(v) for v in sys.argv[1:]]
oracle = H.declare_oracle(C.CDLL("oracle/libkkt_oracle.so"))
F = H.oracle_factor_for(oracle, lp)
Kf = H.kkt_for(vb, lib, lp, mode=vb.MODE_FAST)
Ks = H.kkt_for(vb, lib, lp, mode=vb.MODE_STRICT)
def stats(v): 
    v = np.asarray(v); return "finite=%s absmax=%.3e" % (bool(np.isfinite(v).all()), np.nanmax(np.abs(v)) if v.size else 0)
for it in its:
    with H.capture_stdout():
        E, D, ry, rx, sy, sx = vb.capture_iterate("hsd", lp.m, lp.n, lp.nz, lp.iA, lp.kA, lp.A, lp.b, lp.c, lp.f, it, mode=vb.MODE_FAST, lib=lib)
    print("iter", it, "E", stats(E), "min %.2e" % E.min(), "D", stats(D), "min %.2e" % D.min(), "rhs", stats(ry), stats(rx), "sol", stats(sy), stats(sx), flush=True)
    if not (np.isfinite(E).all() and np.isfinite(D).all()): continue
    Kf.factor(E, D); Lf, df, mf = Kf.get_factor()
    Ks.factor(E, D); Ls, ds, ms = Ks.get_factor()
    F.factor(E, D)
    print("   ndep fast %d strict %d oracle %d | marks0 %d %d | diag fast %s | strict %s | L fast %s strict %s" % (Kf.ndep, Ks.ndep, F.ndep, (mf==0).sum(), (ms==0).sum(), stats(df), stats(ds), stats(Lf), stats(Ls)))
    bad = np.where(~np.isfinite(df))[0]
    print("   nonfinite diag idx", bad[:10], "rel diff diag %.2e" % P._rel(np.nan_to_num(df), ds))
    gy, gx, _ = Kf.solve(E, D, ry, rx); print("   fast solve", stats(gy), stats(gx), "passes", Kf.last_passes)
    oy, ox, _ = Ks.solve(E, D, ry, rx); print("   strict solve", stats(oy), stats(ox), "passes", Ks.last_passes, "rel fast vs strict %.2e %.2e" % (P._rel(np.nan_to_num(gy), oy), P._rel(np.nan_to_num(gx), ox)))
