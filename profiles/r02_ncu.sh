#!/bin/bash
# Round-2 evidence run (one GPU): the bench line, then -- only after it exited 0 -- the ncu launch list of the same
# command and one `--set full` capture each of the strict factor kernel and of the two-right-hand-side sweeps.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash profiles/r02_ncu.sh'
set -u
mkdir -p gpurun_out
BENCH="python bench.py --steps 2 --warmup 3 --no-solve-time --no-strict --no-cpu-baseline"
$BENCH > gpurun_out/r02_bench_short.json 2> gpurun_out/r02_bench_short.err || { echo "bench failed"; tail -5 gpurun_out/r02_bench_short.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_strict.csv $BENCH > gpurun_out/r02_ncu_launches.log 2>&1
python profiles/strict_step.py --reps 1 dfl001 > /dev/null 2>&1 || { echo "strict_step failed"; exit 1; }
ncu --set full --clock-control none --import-source on --kernel-name regex:k_factor_pipe --launch-skip 1 --launch-count 1 -f -o gpurun_out/r02_factor_pipe \
    python profiles/strict_step.py --reps 2 dfl001 > gpurun_out/r02_ncu_pipe.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name 'regex:k_bwd_pipe|k_fwd_flags' --launch-skip 2 --launch-count 2 -f -o gpurun_out/r02_sweeps2 \
    $BENCH > gpurun_out/r02_ncu_sweeps.log 2>&1
ls -la gpurun_out/r02_*
