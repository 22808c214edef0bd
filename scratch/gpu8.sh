TAG="base" python scratch/rand_sweep.py 0 1 2 3
TAG="tol64+static1.5e-8" VBK_PIVOT_TOL_ULPS=64 VBK_PIVOT_STATIC=1.5e-8 python scratch/rand_sweep.py 0 1 2 3
TAG="tol1+static1.5e-8" VBK_PIVOT_STATIC=1.5e-8 python scratch/rand_sweep.py 0 1 2 3
TAG="tol64+static1e-10" VBK_PIVOT_TOL_ULPS=64 VBK_PIVOT_STATIC=1e-10 python scratch/rand_sweep.py 0 1 2 3
VBK_PIVOT_TOL_ULPS=64 VBK_PIVOT_STATIC=1.5e-8 python profiles/fast_sweep.py > gpurun_out/sweep_t64_s15e-9.jsonl 2>/dev/null; tail -1 gpurun_out/sweep_t64_s15e-9.jsonl
VBK_PIVOT_STATIC=1.5e-8 python profiles/fast_sweep.py > gpurun_out/sweep_t1_s15e-9.jsonl 2>/dev/null; tail -1 gpurun_out/sweep_t1_s15e-9.jsonl
