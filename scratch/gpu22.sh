set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv,noheader
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=12 ) > gpurun_out/s9_pytest_gpu.log 2>&1; tail -25 gpurun_out/s9_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/s9_bench_dfl001.json 2> gpurun_out/s9_bench.err; tail -3 gpurun_out/s9_bench.err; cat gpurun_out/s9_bench_dfl001.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/s9_bench_ref.json 2>/dev/null; cat gpurun_out/s9_bench_ref.json
python bench.py --workload pilot87 --no-cpu-baseline > gpurun_out/s9_bench_pilot87.json 2>/dev/null; cat gpurun_out/s9_bench_pilot87.json
# launch list of the bench command (fast mode only, no side measurements)
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r01_ncu_launches_bench_dfl001_fast.csv python bench.py --steps 2 --warmup 3 --no-strict --no-cpu-baseline > gpurun_out/s9_ncu1.log 2>&1
python profiles/summarize_launches.py gpurun_out/r01_ncu_launches_bench_dfl001_fast.csv 20
# per-launch DRAM traffic
VBK_LOOKAHEAD=0 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r01_dram_launches_dfl001_fast.csv python profiles/fast_one.py dfl001 > /dev/null 2>&1
# full capture of representative launches of the heavy kernels
VBK_LOOKAHEAD=0 ncu --set full --import-source on --clock-control none -k regex:'k_dense_update_k|k_panel_diag|k_panel_rows' --launch-skip 6 -c 3 -f -o gpurun_out/r01_full_panel_dfl001 python profiles/fast_one.py dfl001 > /dev/null 2>&1
ncu -i gpurun_out/r01_full_panel_dfl001.ncu-rep --page raw --csv > gpurun_out/r01_full_panel_dfl001_raw.csv 2>/dev/null
ncu -i gpurun_out/r01_full_panel_dfl001.ncu-rep --page details --csv > gpurun_out/r01_full_panel_dfl001_details.csv 2>/dev/null
VBK_PROF=1 VBK_LOOKAHEAD=0 python profiles/fast_one.py dfl001 2>&1 | grep -i "profile" | tail -4
find gpurun_out -name "*.ncu-rep" -size +30M -delete
du -sh gpurun_out
