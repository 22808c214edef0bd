"""Full hsd solves of the synthetic multicommodity LPs of BASELINE config 3 on the GPU (fast mode; strict mode refuses
LPs whose slice-task block table would not fit): status, iterations, seconds, primal/dual objective and infeasibilities.

    python profiles/mcf_solve.py 32:25 50:40
"""
import importlib.util
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "tests"))
import harness as H  # noqa: E402

os.environ.setdefault("VBK_SYM_CACHE", str(ROOT / "tests" / "golden" / "symcache"))


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


vb = _load("vbkkt", ROOT / "linear-programming-vanderbei_b200" / "__init__.py")
vbw = _load("vbkkt_workloads", ROOT / "linear-programming-vanderbei_b200" / "workloads.py")
lib = vb.load()
for arg in sys.argv[1:]:
    R, K = (int(v) for v in arg.split(":"))
    lp = vbw.multicommodity_lp(R, K)
    t0 = time.perf_counter()
    st, log, x, y, prof = H.solve_via(vb, lib, lp, "hsd", mode=vb.MODE_FAST, profile=True)
    dt = time.perf_counter() - t0
    import scipy.sparse as sp
    A = sp.csc_matrix((lp.A, lp.iA, lp.kA), shape=(lp.m, lp.n))
    pinf = float(np.linalg.norm(np.maximum(A @ x - lp.b, 0.0)) / (1 + np.linalg.norm(lp.b)))
    dinf = float(np.linalg.norm(np.maximum(lp.c - A.T @ y, 0.0)) / (1 + np.linalg.norm(lp.c)))
    pobj, dobj = float(lp.c @ x), float(lp.b @ y)
    print(json.dumps(dict(lp=f"mcf:{R}:{K}", m=lp.m, n=lp.n, nz=lp.nz, status=int(st), iterations=len(H.iteration_lines(log)),
                          seconds=round(dt, 2), setup_s=round(prof["setup_s"], 2), factor_ms=round(1e3 * prof["factor_s"] / max(prof["factor_calls"], 1), 2),
                          solve_ms=round(1e3 * prof["solve_s"] / max(prof["solve_calls"], 1), 2), primal_obj=pobj, dual_obj=dobj,
                          rel_gap=abs(pobj - dobj) / (1 + abs(pobj)), primal_infeas_rel=pinf, dual_infeas_rel=dinf,
                          x_min=float(x.min()), last_log_line=log.strip().splitlines()[-1][:120])), flush=True)
