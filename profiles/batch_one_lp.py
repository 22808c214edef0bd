"""One random sparse LP of the batch workload (m=2000, n=4000, seed 0) through vbk_solve_batch, fast mode, one stream.
Run under `ncu --metrics gpu__time_duration.sum --clock-control none` for the per-kernel launch list of a small LP
(profiles/r02_launches_batch_one_lp.csv; summary by profiles/launch_summary in r02_summary.md)."""
import importlib.util, pathlib, sys, time
ROOT = pathlib.Path(__file__).resolve().parent.parent
spec = importlib.util.spec_from_file_location("vbkkt", ROOT / "linear-programming-vanderbei_b200" / "__init__.py",
                                              submodule_search_locations=[str(ROOT / "linear-programming-vanderbei_b200")])
vb = importlib.util.module_from_spec(spec); sys.modules["vbkkt"] = vb; spec.loader.exec_module(vb)
lib = vb.load()
lp = vb.workloads.random_sparse_lp(0, 2000, 4000)
t0 = time.perf_counter()
r = vb.batch.solve_local(lib, [lp], "hsd", 0, vb.MODE_FAST, 1)[0]
print({"status": r["status"], "iterations": r["iterations"], "seconds": time.perf_counter() - t0,
       "launches": int(lib.vbk_batch_launches())})
