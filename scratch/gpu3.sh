for sz in "500 1000" "1000 2000" "2000 4000"; do
TAG="fast $sz" python scratch/dbg_rand.py $sz fast 8
done
TAG="strict 2000" python scratch/dbg_rand.py 2000 4000 strict 8
TAG="rho1" VBK_WINDOW_RHO=1 python scratch/dbg_rand.py 2000 4000 fast 8
TAG="schur light" VBK_SCHUR=light python scratch/dbg_rand.py 2000 4000 fast 8
TAG="schur heavy" VBK_SCHUR=heavy python scratch/dbg_rand.py 2000 4000 fast 8
TAG="dense v1" VBK_DENSE=v1 python scratch/dbg_rand.py 2000 4000 fast 8
TAG="wsolve v1" VBK_WSOLVE=v1 python scratch/dbg_rand.py 2000 4000 fast 8
