import sys, os, ctypes as C, importlib.util, subprocess, numpy as np, time
sys.path.insert(0, "tests")
import harness as H, parity as P, conftest
vb = conftest._load_pkg()
lib = vb.load(os.environ.get("VBK_LIB"))
m, n = int(sys.argv[1]), int(sys.argv[2])
mode = vb.MODE_FAST if sys.argv[3] == "fast" else vb.MODE_STRICT
itn = int(sys.argv[4])
lp = vb.workloads.random_sparse_lp(0, m, n)
lib.vbk_set_iteration_limit(itn)
t0 = time.time()
with H.capture_stdout() as cap:
    st, x, y, prof = vb.solve_lp("hsd", lp.m, lp.n, lp.nz, lp.iA, lp.kA, lp.A, lp.b, lp.c, lp.f, mode=mode, profile=True)
lines = H.iteration_lines(cap.text)
print(os.environ.get("TAG", ""), "status", st, "lines", len(lines), "t %.2f" % (time.time() - t0), "factor ms %.2f solve ms %.2f" % (1e3 * prof["factor_s"] / max(1, prof["factor_calls"]), 1e3 * prof["solve_s"] / max(1, prof["solve_calls"])), "passes/solve %.2f" % (prof["refine_passes"] / max(1, prof["solve_calls"])))
for l in lines[:3] + lines[-3:]: print("   ", l)
