for t in 1 8 64 1024 65536; do
TAG="tol $t" VBK_PIVOT_TOL_ULPS=$t python scratch/dbg_rand.py 2000 4000 fast 200 | cut -c1-150 | head -1
done
TAG="tol 64 rho1" VBK_WINDOW_RHO=1 VBK_PIVOT_TOL_ULPS=64 python scratch/dbg_rand.py 2000 4000 fast 200 | head -1
TAG="tol 1 rho1" VBK_WINDOW_RHO=1 python scratch/dbg_rand.py 2000 4000 fast 200 | head -1
python profiles/fast_sweep.py > gpurun_out/sweep_tol1.jsonl 2> gpurun_out/sweep_tol1.err
tail -1 gpurun_out/sweep_tol1.jsonl
VBK_PIVOT_TOL_ULPS=64 python profiles/fast_sweep.py > gpurun_out/sweep_tol64.jsonl 2> gpurun_out/sweep_tol64.err
tail -1 gpurun_out/sweep_tol64.jsonl
