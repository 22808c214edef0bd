"""Critical-path breakdown of one STRICT factorisation from the per-column event trace ($VBK_PROF=1,
csrc/vbk_strict_factor.cuh): walks from the last column down the elimination tree, always to the child
that finished last, and sums per phase where the time of that path went.

    VBK_PROF=1 python profiles/strict_trace.py pilot87 [dfl001 ...]
"""
import ctypes as C
import importlib.util
import json
import os
import sys

import numpy as np

sys.path.insert(0, "tests")
import harness as H  # noqa: E402

os.environ.setdefault("VBK_PROF", "1")
spec = importlib.util.spec_from_file_location("vbkkt", "linear-programming-vanderbei_b200/__init__.py")
vb = importlib.util.module_from_spec(spec)
spec.loader.exec_module(vb)
lib = vb.load()
lib.vbk_kkt_trace.argtypes = [C.c_void_p, C.POINTER(C.c_longlong)]

for name in sys.argv[1:]:
    lp = H.load_fixture(name)
    z = np.load(H.GOLDEN / "iterates" / f"{name}_it20.npz")
    K = H.kkt_for(vb, lib, lp)
    K.factor(z["E"], z["D"])
    K.factor(z["E"], z["D"])
    ms = float(lib.vbk_kkt_last_factor_kernel_ms(K.h))
    N = K.dim
    tr = np.zeros((N, 8), dtype=np.int64)
    lib.vbk_kkt_trace(K.h, tr.ctypes.data_as(C.POINTER(C.c_longlong)))
    kL, iL = K.kAAt, K.iAAt
    parent = np.full(N, -1, dtype=np.int64)
    nz = kL[1:] > kL[:-1]
    parent[nz] = iL[kL[:-1][nz]]
    # last-finishing child of every column
    last_child = np.full(N, -1, dtype=np.int64)
    best = np.zeros(N, dtype=np.int64)
    for j in range(N):
        p = parent[j]
        if p >= 0 and tr[j, 4] > best[p]:
            best[p] = tr[j, 4]
            last_child[p] = j
    # walk down from the column that finished last
    i = int(np.argmax(tr[:, 4]))
    t_end = tr[i, 4]
    acc = dict(levels=0, handoff_first_stage=0, chains=0, slices_publish=0, pivot_divide_done=0, links=0, groups=0)
    while i >= 0:
        c = last_child[i]
        t_child = tr[c, 4] if c >= 0 else tr[i, 0]
        first = max(tr[i, 1], t_child)
        acc["levels"] += 1
        acc["handoff_first_stage"] += max(0, tr[i, 1] - t_child) if tr[i, 5] > 0 else 0
        acc["chains"] += max(0, tr[i, 2] - first) if tr[i, 5] > 0 else 0
        acc["slices_publish"] += max(0, tr[i, 3] - max(tr[i, 2], t_child))
        acc["pivot_divide_done"] += max(0, tr[i, 4] - tr[i, 3])
        acc["links"] += int(K.kAAt[0] * 0)  # placeholder keeps the dict shape
        acc["groups"] += int(tr[i, 5])
        if c < 0:
            t_start = tr[i, 0]
        i = int(c)
    span_us = (t_end - t_start) / 1e3
    lv = acc["levels"]
    out = dict(name=name, factor_ms=round(ms, 3), path_levels=lv, path_span_ms=round(span_us / 1e3, 3), path_groups=acc["groups"],
               us_per_level={k: round(acc[k] / 1e3 / lv, 3) for k in ("handoff_first_stage", "chains", "slices_publish", "pivot_divide_done")},
               ms_total={k: round(acc[k] / 1e6, 3) for k in ("handoff_first_stage", "chains", "slices_publish", "pivot_divide_done")},
               ns_per_group_chain=round(acc["chains"] / max(acc["groups"], 1), 1))
    print(json.dumps(out), flush=True)
    K.close()
