"""Per-phase cycle breakdown of the tiled strict factor kernel ($VBK_PROF=1) on one KKT step."""
import ctypes as C, importlib.util, json, os, sys, time
import numpy as np
os.environ["VBK_PROF"] = "1"
sys.path.insert(0, "tests")
import harness as H
spec = importlib.util.spec_from_file_location("vbkkt", "linear-programming-vanderbei_b200/__init__.py")
vb = importlib.util.module_from_spec(spec); spec.loader.exec_module(vb)
lib = vb.load()
lib.vbk_kkt_phase_profile.argtypes = [C.c_void_p, C.POINTER(C.c_ulonglong)]
names = ["claim+init", "wait", "stage", "scan", "scatter", "accumulate", "pivot", "write"]
for name in sys.argv[1:]:
    lp = H.load_fixture(name)
    z = np.load(H.GOLDEN / "iterates" / f"{name}_it20.npz")
    K = H.kkt_for(vb, lib, lp)
    for rep in range(3):
        K.factor(z["E"], z["D"])
    buf = (C.c_ulonglong * 8)()
    lib.vbk_kkt_phase_profile(K.h, buf)
    t0 = time.perf_counter(); K.factor(z["E"], z["D"]); dt = time.perf_counter() - t0
    ms = lib.vbk_kkt_last_factor_kernel_ms(K.h)
    lib.vbk_kkt_phase_profile(K.h, buf)
    tot = sum(buf) or 1
    ry, rx = z["rhs_y"], z["rhs_x"]
    t0 = time.perf_counter(); K.solve(z["E"], z["D"], ry, rx); ds = time.perf_counter() - t0
    print(json.dumps({"name": name, "cap": os.environ.get("VBK_WHOLE_CAP", "512"), "factor_kernel_ms": round(ms, 3),
                      "factor_call_ms": round(dt * 1e3, 3), "solve_call_ms": round(ds * 1e3, 3), "passes": K.last_passes,
                      "phase_share": {n: round(b / tot, 3) for n, b in zip(names, buf)}}), flush=True)
    K.close()
