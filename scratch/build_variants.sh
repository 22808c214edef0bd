#!/bin/bash
# builds libvbkkt variants with different producer batch sizes into scratch/var/<FB>_<QB>/
cd /root/repo/linear-programming-vanderbei_b200/csrc
for cfg in "$@"; do
  fb=${cfg%_*}; qb=${cfg#*_}
  mkdir -p /root/repo/scratch/var/$cfg
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC,-ffp-contract=off,-fvisibility=default -Xlinker -Bsymbolic -shared -cudart static -DVBK_PIPE_FB=$fb -DVBK_PIPE_QB=$qb -I . -o /root/repo/scratch/var/$cfg/libvbkkt.so vbk_symbolic.cpp vbk_kkt.cu vbk_kkt_fast.cu vbk_linalg.cu vbk_solver.cu vbk_batch.cu vbk_rowblock.cu vbk_capi.cu &
done
wait
