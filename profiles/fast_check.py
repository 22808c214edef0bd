"""Strict vs fast mode on the committed hsd iterates: time of one KKT step and agreement with the
reference's solution stored in the fixture."""
import importlib.util, json, sys, time
import numpy as np
sys.path.insert(0, "tests")
import harness as H
spec = importlib.util.spec_from_file_location("vbkkt", "linear-programming-vanderbei_b200/__init__.py")
vb = importlib.util.module_from_spec(spec); spec.loader.exec_module(vb)
lib = vb.load()
rel = lambda a, b: float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))
for name in sys.argv[1:]:
    lp = H.load_fixture(name)
    z = np.load(H.GOLDEN / "iterates" / f"{name}_it20.npz")
    out = {"name": name}
    for mode, tag in ((vb.MODE_STRICT, "strict"), (vb.MODE_FAST, "fast")):
        K = H.kkt_for(vb, lib, lp, mode=mode)
        for _ in range(2):
            K.factor(z["E"], z["D"])
        t0 = time.perf_counter(); K.factor(z["E"], z["D"]); tf = time.perf_counter() - t0
        kms = lib.vbk_kkt_last_factor_kernel_ms(K.h)
        t0 = time.perf_counter(); sy, sx, _ = K.solve(z["E"], z["D"], z["rhs_y"], z["rhs_x"]); ts = time.perf_counter() - t0
        out[tag] = {"factor_call_ms": round(tf * 1e3, 3), "factor_kernels_ms": round(kms, 3), "solve_call_ms": round(ts * 1e3, 3),
                    "passes": K.last_passes, "ndep": K.ndep, "sol_relerr_y": rel(sy, z["sol_y"]), "sol_relerr_x": rel(sx, z["sol_x"])}
        K.close()
    print(json.dumps(out), flush=True)
