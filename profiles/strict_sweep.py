"""Strict-mode full solves over every committed netlib fixture (BASELINE config 2): byte-identical golden log,
bit-equal x and y, solve time per problem.     python profiles/strict_sweep.py [names...]"""
import json, sys, time
import numpy as np
sys.path.insert(0, "tests")
import harness as H, conftest
vb = conftest._load_pkg()
lib = vb.load()
idx = {r["name"]: r for r in json.load(open("tests/golden/netlib/index.json")) if "name" in r}
names = sys.argv[1:] or H.fixture_names()
rows = []
for name in names:
    lp = H.load_fixture(name)
    if "hsd_log" not in lp.extra:
        continue
    t0 = time.time()
    st, log, x, y, prof = H.solve_via(vb, lib, lp, "hsd", mode=vb.MODE_STRICT, profile=True)
    ok = bool(log == str(lp.extra["hsd_log"]) and st == int(lp.extra["hsd_status"])
              and np.array_equal(x, lp.extra["hsd_x"]) and np.array_equal(y, lp.extra["hsd_y"]))
    rows.append(dict(name=name, N=lp.m + lp.n, lnz=prof["lnz"], status=st, iterations=prof["iterations"], parity=ok,
                     gpu_s=round(prof["total_s"], 3), setup_s=round(prof["setup_s"], 3),
                     factor_ms=round(1e3 * prof["factor_s"] / max(prof["factor_calls"], 1), 3),
                     solve_ms=round(1e3 * prof["solve_s"] / max(prof["solve_calls"], 1), 3),
                     ref_cpu_s=idx.get(name, {}).get("seconds")))
    print(json.dumps(rows[-1]), flush=True)
print(json.dumps(dict(summary=True, problems=len(rows), parity_ok=sum(r["parity"] for r in rows),
                      failed=[r["name"] for r in rows if not r["parity"]], gpu_total_s=round(sum(r["gpu_s"] for r in rows), 1),
                      ref_cpu_total_s=round(sum(r["ref_cpu_s"] or 0 for r in rows), 1))))
