"""Timing of the dense tail of one STRICT factorisation from the per-column event trace ($VBK_PROF=1,
csrc/vbk_strict_factor.cuh): for the columns near the end of the elimination order (each one's youngest child is the
column before it) the time from the child's col_done to the column's own claim, first consumed ring stage, end of the
chains, publication of the slices and col_done, and the chain's rate in ns per group of 32 contributors.

    VBK_PROF=1 python profiles/strict_tail_trace.py dfl001

What it showed in round 2 (dfl001): the column's last slice is claimed ~150 us BEFORE its child is final, but the first
stage is consumed only 2.4 us AFTER it -- the reference's accumulation order starts with the youngest child's product
(DESIGN.md 5.1), so the whole chain (70-100 groups at ~370 ns) sits behind the hand-off."""
import ctypes as C
import importlib.util
import os
import sys

import numpy as np

sys.path.insert(0, "tests")
import harness as H  # noqa: E402

os.environ.setdefault("VBK_PROF", "1")
spec = importlib.util.spec_from_file_location("vbkkt", "linear-programming-vanderbei_b200/__init__.py")
vb = importlib.util.module_from_spec(spec)
spec.loader.exec_module(vb)
lib = vb.load(os.environ.get("VBK_LIB"))
lib.vbk_kkt_trace.argtypes = [C.c_void_p, C.POINTER(C.c_longlong)]
name = sys.argv[1]
lp = H.load_fixture(name)
z = np.load(H.GOLDEN / "iterates" / f"{name}_it20.npz")
K = H.kkt_for(vb, lib, lp)
K.factor(z["E"], z["D"])
K.factor(z["E"], z["D"])
N = K.dim
tr = np.zeros((N, 8), dtype=np.int64)
lib.vbk_kkt_trace(K.h, tr.ctypes.data_as(C.POINTER(C.c_longlong)))
rows = []
for i in range(max(1, N - 1500), N - 200):
    tc = tr[i - 1, 4]
    rows.append((i, tr[i, 6], tr[i, 5], (tr[i, 0] - tc) / 1e3, (tr[i, 1] - tc) / 1e3, (tr[i, 2] - tc) / 1e3, (tr[i, 3] - tc) / 1e3,
                 (tr[i, 4] - tc) / 1e3, (tr[i, 2] - tr[i, 1]) / max(tr[i, 5], 1)))
a = np.array(rows, dtype=np.float64)
print(f"{name}: columns {int(a[0, 0])}..{int(a[-1, 0])}, slices per column (median) {np.median(a[:, 1]):.0f}, "
      f"contributor groups per column (median) {np.median(a[:, 2]):.0f}")
for k, nm in enumerate(["claim - child done [us]", "first stage consumed - child done [us]", "chains done - child done [us]",
                        "slices published - child done [us]", "col_done - child done [us]", "chain rate [ns per group]"]):
    v = a[:, 3 + k]
    print(f"{nm:42s} p10 {np.percentile(v, 10):9.2f}  median {np.median(v):9.2f}  p90 {np.percentile(v, 90):9.2f}")
K.close()
