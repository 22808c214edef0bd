import sys, os, time, numpy as np
sys.path.insert(0, "tests")
import harness as H, conftest
vb = conftest._load_pkg(); lib = vb.load()
seeds = [int(v) for v in sys.argv[1:]] or [0, 1, 2, 3]
for sd in seeds:
    lp = vb.workloads.random_sparse_lp(sd, 2000, 4000)
    with H.capture_stdout() as cap:
        st, x, y, prof = vb.solve_lp("hsd", lp.m, lp.n, lp.nz, lp.iA, lp.kA, lp.A, lp.b, lp.c, lp.f, mode=vb.MODE_FAST, profile=True)
    lines = H.iteration_lines(cap.text)
    print(os.environ.get("TAG", ""), "seed", sd, "status", st, "lines", len(lines), "obj %.9e dual %.9e" % (lp.c @ x, lp.b @ y), "last:", lines[-1][9:] if lines else "", flush=True)
