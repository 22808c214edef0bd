"""Per-call timing of the STRICT KKT step on committed iterates (tests/golden/iterates): numeric factor
(CUDA events around the factor kernel), forwardbackward through host buffers (wall clock around the
C-ABI call), bit-equality of the solution with the fixture.  Usage:

    python profiles/strict_step.py [--reps R] pilot87 dfl001 ...

Environment variables of the library (VBK_PIPE_WARPS, VBK_PIPE_STAGES, VBK_FACTOR=tiled, ...) are read
when the handle is analysed, so `--env K=V,K=V/K=V` runs the same problems under several settings."""
import importlib.util
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, "tests")
import harness as H  # noqa: E402

spec = importlib.util.spec_from_file_location("vbkkt", "linear-programming-vanderbei_b200/__init__.py")
vb = importlib.util.module_from_spec(spec)
spec.loader.exec_module(vb)
lib = vb.load(os.environ.get("VBK_LIB"))


def run(name, reps, tag):
    lp = H.load_fixture(name)
    z = np.load(H.GOLDEN / "iterates" / f"{name}_it20.npz")
    t0 = time.perf_counter()
    K = H.kkt_for(vb, lib, lp)
    t_sym = time.perf_counter() - t0
    fac, fb = [], []
    ok = True
    for _ in range(reps):
        K.factor(z["E"], z["D"])
        fac.append(float(lib.vbk_kkt_last_factor_kernel_ms(K.h)))
        t0 = time.perf_counter()
        sy, sx, _ = K.solve(z["E"], z["D"], z["rhs_y"], z["rhs_x"])
        fb.append(1e3 * (time.perf_counter() - t0))
        ok = ok and bool(np.array_equal(sy, z["sol_y"]) and np.array_equal(sx, z["sol_x"]))
    prof = None
    if os.environ.get("VBK_PROF"):
        import ctypes as C
        buf = (C.c_ulonglong * 16)()
        lib.vbk_kkt_phase_profile.argtypes = [C.c_void_p, C.POINTER(C.c_ulonglong)]
        lib.vbk_kkt_phase_profile(K.h, buf)
        v = [int(x) for x in buf]
        names = ["cons_wait", "cons_add", "prod_slot_wait", "prod_static_zero", "prod_col_wait", "prod_fence_lij", "prod_products",
                 "prod_publish", "setup", "epi_pub", "owner_wait", "pivot", "divide", "done_fence", "groups", "tasks"]
        prof = dict(zip(names, v))
        g = max(v[14], 1)
        prof["per_group"] = {k: round(prof[k] / g, 1) for k in names[2:8]}
        prof["cons_add_per_group"] = round(v[1] / g * (1.0), 1)
    out = dict(name=name, prof=prof, env=tag, N=K.dim, lnz=K.lnz, narth=K.narth, levels=K.nlevels, analyze_s=round(t_sym, 3),
               factor_ms=[round(v, 3) for v in fac], forwardbackward_ms=[round(v, 3) for v in fb],
               passes=K.last_passes, ns_per_link=round(1e6 * min(fac) / max(K.lnz, 1), 2), bit_equal=ok)
    K.close()
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    args = sys.argv[1:]
    reps, envs = 3, [""]
    while args and args[0].startswith("--"):
        if args[0] == "--reps":
            reps = int(args[1]); args = args[2:]
        elif args[0] == "--env":
            envs = args[1].split("/"); args = args[2:]
        else:
            raise SystemExit("unknown option " + args[0])
    for e in envs:
        saved = {}
        for kv in filter(None, e.split(",")):
            k, v = kv.split("=", 1)
            saved[k] = os.environ.get(k)
            os.environ[k] = v
        for name in args:
            run(name, reps, e)
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
