TAG="fast full" python scratch/dbg_rand.py 2000 4000 fast 200 | cut -c1-150 | head -12
python - <<'PY'
import sys, os
sys.path.insert(0, "tests")
import harness as H, conftest
vb = conftest._load_pkg(); lib = vb.load()
lp = vb.workloads.random_sparse_lp(0, 2000, 4000)
with H.capture_stdout() as cap:
    st, x, y, prof = vb.solve_lp("hsd", lp.m, lp.n, lp.nz, lp.iA, lp.kA, lp.A, lp.b, lp.c, lp.f, mode=vb.MODE_FAST, profile=True)
lines = H.iteration_lines(cap.text)
for l in lines[20:80]: print(l)
PY
TAG="strict full" python scratch/dbg_rand.py 2000 4000 strict 200 | head -12
