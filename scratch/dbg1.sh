set -x
L=linear-programming-vanderbei_b200/libvbkkt.so
python profiles/fast_solve.py $L blend adlittle sc205 2>&1 | grep -v "^ *[0-9]* " | tail -8
VBK_WSOLVE=v1 python profiles/fast_solve.py $L blend 2>&1 | grep fast
VBK_DENSE=v1 python profiles/fast_solve.py $L blend 2>&1 | grep fast
VBK_DENSE=v1 VBK_WSOLVE=v1 python profiles/fast_solve.py $L blend 2>&1 | grep fast
VBK_SCHUR=light python profiles/fast_solve.py $L blend 2>&1 | grep fast
VBK_SCHUR=heavy python profiles/fast_solve.py $L blend 2>&1 | grep fast
VBK_WINDOW_RHO=1 python profiles/fast_solve.py $L blend 2>&1 | grep fast
for i in 1 2 3; do python profiles/fast_solve.py $L blend 2>&1 | grep fast; done
