import sys, time, importlib.util, numpy as np
sys.path.insert(0, "tests")
import harness as H
spec = importlib.util.spec_from_file_location("vbkkt", "linear-programming-vanderbei_b200/__init__.py")
vb = importlib.util.module_from_spec(spec); sys.modules["vbkkt"]=vb; spec.loader.exec_module(vb)
spec2 = importlib.util.spec_from_file_location("vbkkt.workloads", "linear-programming-vanderbei_b200/workloads.py")
wl = importlib.util.module_from_spec(spec2); sys.modules["vbkkt.workloads"]=wl; spec2.loader.exec_module(wl)
lib = vb.load()
def analyze(lp):
    t0=time.time()
    kAt, iAt, At = H.transpose_csc(lp.m, lp.n, lp.kA, lp.iA, lp.A)
    k = vb.KKT(device=-1, lib=lib)
    k.analyze(lp.n, lp.m, kAt, iAt, At, lp.kA, lp.iA, lp.A)
    print(lp.name, "m,n,nz", lp.m, lp.n, lp.nz, "N", k.dim, "Lnz", k.lnz, "narth %.3e" % k.narth, "denwin", k.dim-k.denwin, "levels", k.nlevels, "sn", k.nsupernodes, "t %.2fs" % (time.time()-t0), flush=True)
    k.close()
what = sys.argv[1]
if what == "rand":
    for m,n in [(200,400),(500,1000),(1000,2000),(2000,4000)]:
        analyze(wl.random_sparse_lp(0, m, n))
else:
    for R,K in [(4,3),(8,5),(12,8),(16,10),(20,12)]:
        analyze(wl.multicommodity_lp(R,K))
