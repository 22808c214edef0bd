"""Orderings of the synthetic multicommodity LPs of BASELINE config 3, computed once on the host and kept as
$VBK_SYM_CACHE files (8 bytes per row/column of K): the explicit-fill minimum-degree ordering of the reference takes
minutes at these sizes (460 s for R=50/K=40) and depends on the pattern only.

    python tests/golden/make_symcache.py 32:25 50:40
"""
import importlib.util
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT / "tests"))
import harness as H  # noqa: E402

os.environ["VBK_SYM_CACHE"] = str(ROOT / "tests" / "golden" / "symcache")


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
    Kk = H.kkt_for(vb, lib, lp, device=-1)
    print(f"mcf:{R}:{K} m {lp.m} n {lp.n} nz {lp.nz}: analyze {time.perf_counter() - t0:.1f} s, N {Kk.dim}, Lnz {Kk.lnz}, "
          f"window {Kk.window}, narth {Kk.narth:.4g}", flush=True)
    Kk.close()
