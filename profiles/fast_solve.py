import sys, time, importlib.util, numpy as np
sys.path.insert(0, "tests")
import harness as H
spec = importlib.util.spec_from_file_location("vbkkt", "linear-programming-vanderbei_b200/__init__.py")
vb = importlib.util.module_from_spec(spec); spec.loader.exec_module(vb)
lib = vb.load(sys.argv[1])
for name in sys.argv[2:]:
    lp = H.load_fixture(name)
    for meth in ("hsd",):
        t0 = time.time()
        st, log, x, y, prof = H.solve_via(vb, lib, lp, meth, mode=vb.MODE_FAST, profile=True)
        a = H.iteration_lines(log); b = H.iteration_lines(str(lp.extra[meth + "_log"]))
        obj = float(lp.c @ x); obj_r = float(lp.c @ lp.extra[meth + "_x"])
        print(f"{name} {meth} fast: status {st}/{int(lp.extra[meth+'_status'])} lines {len(a)}/{len(b)} obj relerr {abs(obj-obj_r)/max(1,abs(obj_r)):.2e} "
              f"total {prof['total_s']:.2f}s factor {1e3*prof['factor_s']/max(prof['factor_calls'],1):.2f} ms/call solve {1e3*prof['solve_s']/max(prof['solve_calls'],1):.2f} ms/call", flush=True)
        if len(a) and a[-1] != b[-1]: print("   last lines:\n   ", a[-1], "\n   ", b[-1])
