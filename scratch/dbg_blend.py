import sys, ctypes as C, importlib.util, subprocess, numpy as np
sys.path.insert(0, "tests")
import harness as H, parity as P
spec = importlib.util.spec_from_file_location("vbkkt", "linear-programming-vanderbei_b200/__init__.py")
vb = importlib.util.module_from_spec(spec); spec.loader.exec_module(vb)
import os
lib = vb.load(os.environ.get("VBK_LIB"))
subprocess.run(["make", "-C", "oracle", "restatement"], check=True, stdout=subprocess.DEVNULL)
oracle = H.declare_oracle(C.CDLL("oracle/libkkt_oracle.so"))
name = sys.argv[1]
lp = H.load_fixture(name)
nit = len(H.iteration_lines(str(lp.extra["hsd_log"])))
F = H.oracle_factor_for(oracle, lp)
K = H.kkt_for(vb, lib, lp, mode=vb.MODE_FAST)
for it in range(max(0, nit - 14), nit - 1):
    with H.capture_stdout():
        E, D, ry, rx, sy, sx = H.capture_step(oracle, lp, "hsd", it)
    F.factor(E, D); K.factor(E, D)
    L, d, mk = K.get_factor()
    oy, ox, _ = F.solve(E, D, ry, rx)
    gy, gx, _ = K.solve(E, D, ry, rx)
    print(it, "ndep", F.ndep, K.ndep, "mark_eq", bool(np.array_equal(mk, F.mark)), "nmark0", int((F.mark == 0).sum()), int((mk == 0).sum()),
          "rel d %.2e L %.2e" % (P._rel(d, F.diag), P._rel(L, F.L)), "sol y %.2e x %.2e" % (P._rel(gy, oy), P._rel(gx, ox)),
          "passes", F.passes, K.last_passes, flush=True)
