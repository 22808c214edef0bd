import ctypes as C, importlib.util, json, os, sys
import numpy as np
sys.path.insert(0, "tests")
import harness as H
os.environ.setdefault("VBK_PROF", "1")
spec = importlib.util.spec_from_file_location("vbkkt", "linear-programming-vanderbei_b200/__init__.py")
vb = importlib.util.module_from_spec(spec); spec.loader.exec_module(vb)
lib = vb.load(os.environ.get("VBK_LIB"))
lib.vbk_kkt_trace.argtypes = [C.c_void_p, C.POINTER(C.c_longlong)]
name = sys.argv[1]
lp = H.load_fixture(name)
z = np.load(H.GOLDEN / "iterates" / f"{name}_it20.npz")
K = H.kkt_for(vb, lib, lp)
K.factor(z["E"], z["D"]); K.factor(z["E"], z["D"])
N = K.dim
tr = np.zeros((N, 8), dtype=np.int64)
lib.vbk_kkt_trace(K.h, tr.ctypes.data_as(C.POINTER(C.c_longlong)))
# dense tail: consecutive columns; child of i is i-1
rows = []
for i in range(N - 1500, N - 200):
    tc = tr[i - 1, 4]
    rows.append((i, tr[i, 6], tr[i, 5], (tr[i, 0] - tc) / 1e3, (tr[i, 1] - tc) / 1e3, (tr[i, 2] - tc) / 1e3, (tr[i, 3] - tc) / 1e3, (tr[i, 4] - tc) / 1e3,
                 (tr[i, 2] - tr[i, 1]) / max(tr[i, 5], 1), tr[i, 6] / max(tr[i, 5], 1), tr[i, 7] / max(tr[i, 5], 1), 0))
a = np.array(rows)
print("cols", len(a), "nslices med", np.median(a[:, 1]), "ngroups med", np.median(a[:, 2]))
for k, nm in enumerate(["claim-childdone", "firststage-childdone", "chainsdone-childdone", "published-childdone", "done-childdone", "ns/group", "wait cyc/group", "add cyc/group", "-"]):
    v = a[:, 3 + k]
    print(f"{nm:24s} p10 {np.percentile(v,10):9.2f} med {np.median(v):9.2f} p90 {np.percentile(v,90):9.2f}")
for r in a[600:620]:
    print(" ".join(f"{x:9.2f}" for x in r))
