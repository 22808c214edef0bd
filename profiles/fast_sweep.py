"""Fast-mode full solves over the committed netlib fixtures: per problem, does fast mode stay inside the
north_star tolerances (status, iterations +-1, objective 1e-8) -- SURVEY.md 8d config 2.
    python profiles/fast_sweep.py [names...]      (default: every fixture)"""
import json, sys, time
import numpy as np
sys.path.insert(0, "tests")
import harness as H, conftest
vb = conftest._load_pkg()
lib = vb.load()
names = sys.argv[1:] or H.fixture_names()
rows = []
for name in names:
    lp = H.load_fixture(name)
    if "hsd_log" not in lp.extra:
        continue
    t0 = time.time()
    st, log, x, y, prof = H.solve_via(vb, lib, lp, "hsd", mode=vb.MODE_FAST, profile=True)
    a, b = H.iteration_lines(log), H.iteration_lines(str(lp.extra["hsd_log"]))
    obj, obj_r = float(lp.c @ x), float(lp.c @ lp.extra["hsd_x"])
    rel = abs(obj - obj_r) / max(1.0, abs(obj_r))
    ok = st == int(lp.extra["hsd_status"]) and abs(len(a) - len(b)) <= 1 and rel <= 1e-8
    rows.append(dict(name=name, N=lp.m + lp.n, status=st, ref_status=int(lp.extra["hsd_status"]), lines=len(a), ref_lines=len(b),
                     obj_rel=rel, in_tol=bool(ok), total_s=round(prof["total_s"], 3),
                     factor_ms=round(1e3 * prof["factor_s"] / max(prof["factor_calls"], 1), 3),
                     solve_ms=round(1e3 * prof["solve_s"] / max(prof["solve_calls"], 1), 3)))
    print(json.dumps(rows[-1]), flush=True)
n_ok = sum(r["in_tol"] for r in rows)
print(json.dumps(dict(summary=True, problems=len(rows), in_tolerance=n_ok, same_status=sum(r["status"] == r["ref_status"] for r in rows),
                      nan_or_fail=[r["name"] for r in rows if not np.isfinite(r["obj_rel"])],
                      out=[r["name"] for r in rows if not r["in_tol"]])))
