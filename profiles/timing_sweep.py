import sys, json, time, importlib.util, numpy as np
sys.path.insert(0, "tests")
import harness as H
spec = importlib.util.spec_from_file_location("vbkkt", "linear-programming-vanderbei_b200/__init__.py")
vb = importlib.util.module_from_spec(spec); spec.loader.exec_module(vb)
lib = vb.load()
idx = {r["name"]: r for r in json.load(open("tests/golden/netlib/index.json")) if "m" in r}
for name in sys.argv[1:]:
    lp = H.load_fixture(name)
    t0 = time.time()
    st, log, x, y, prof = H.solve_via(vb, lib, lp, "hsd", profile=True)
    dt = time.time() - t0
    ok = log == str(lp.extra["hsd_log"]) and np.array_equal(x, lp.extra["hsd_x"])
    it = max(prof["iterations"], 1)
    print(json.dumps(dict(name=name, N=prof["N"], lnz=prof["lnz"], narth=prof["narth"], iters=prof["iterations"], parity=bool(ok),
        gpu_total_s=round(prof["total_s"],3), setup_s=round(prof["setup_s"],3), factor_s=round(prof["factor_s"],3), solve_s=round(prof["solve_s"],3),
        factor_ms_per_call=round(1e3*prof["factor_s"]/max(prof["factor_calls"],1),3), solve_ms_per_call=round(1e3*prof["solve_s"]/max(prof["solve_calls"],1),3),
        passes_per_solve=round(prof["refine_passes"]/max(prof["solve_calls"],1),2), launches=prof["kernel_launches"],
        factor_gflops=round(prof["narth"]*prof["factor_calls"]/max(prof["factor_s"],1e-9)/1e9,2),
        ref_cpu_s=idx[name]["seconds"])), flush=True)
