nproc; grep -m1 "model name" /proc/cpuinfo
( time python bench.py --workload batch --batch-per-gpu 8 --steps 1 --streams 4 ) > gpurun_out/s29_a.json 2> gpurun_out/s29_a.err; tail -4 gpurun_out/s29_a.err; cut -c1-200 gpurun_out/s29_a.json
( time python bench.py --workload batch --batch-per-gpu 16 --steps 1 --streams 8 ) > gpurun_out/s29_b.json 2> gpurun_out/s29_b.err; tail -4 gpurun_out/s29_b.err; cut -c1-200 gpurun_out/s29_b.json
VBK_BATCH_PROFILE=1 python - <<'PY'
import sys, time, importlib.util
sys.path.insert(0,'tests')
import harness as H, conftest
vb = conftest._load_pkg(); lib = vb.load()
lp = vb.workloads.random_sparse_lp(0, 2000, 4000)
t0=time.time(); kAt,iAt,At = H.transpose_csc(lp.m, lp.n, lp.kA, lp.iA, lp.A); k = vb.KKT(device=0, mode=vb.MODE_FAST, lib=lib); k.analyze(lp.n, lp.m, kAt, iAt, At, lp.kA, lp.iA, lp.A); t1=time.time()
print("analyze s", t1-t0, "N", k.dim, "Lnz", k.lnz, "window", k.window, "narth %.2e"%k.narth)
t0=time.time(); st, log, x, y, prof = H.solve_via(vb, lib, lp, "hsd", mode=vb.MODE_FAST, profile=True); print("solve total s", time.time()-t0, prof)
PY
