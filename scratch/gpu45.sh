set -x
mkdir -p gpurun_out/ev
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv,noheader > gpurun_out/ev/gpu.txt
./build/ubench > gpurun_out/ev/r01_ubench.txt 2>&1
./build/ubench2 >> gpurun_out/ev/r01_ubench.txt 2>&1
python bench.py > gpurun_out/ev/r01_bench_dfl001.json 2> gpurun_out/ev/bench.err; tail -2 gpurun_out/ev/bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/ev/r01_bench_dfl001_reference.json 2>/dev/null
python bench.py --workload pilot87 --cpu-budget 5 > gpurun_out/ev/r01_bench_pilot87.json 2>/dev/null
python bench.py --workload mcf --steps 5 > gpurun_out/ev/r01_bench_mcf.json 2> gpurun_out/ev/bench_mcf.err; tail -2 gpurun_out/ev/bench_mcf.err
# launch list of the bench command (fast mode only, no side measurements)
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/ev/r01_ncu_launches_bench_dfl001_fast.csv python bench.py --steps 2 --warmup 3 --no-strict --no-cpu-baseline > gpurun_out/ev/ncu1.log 2>&1
python profiles/summarize_launches.py gpurun_out/ev/r01_ncu_launches_bench_dfl001_fast.csv 24 > gpurun_out/ev/r01_ncu_launches_bench_dfl001_fast_summary.txt
# per-launch DRAM traffic (two factorisations + two solves)
VBK_LOOKAHEAD=0 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 3000 --csv --log-file gpurun_out/ev/r01_dram_launches_dfl001_fast.csv python profiles/fast_one.py dfl001 > /dev/null 2>&1
VBK_LOOKAHEAD=0 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 3000 --csv --log-file gpurun_out/ev/r01_dram_launches_mcf_fast.csv python profiles/fast_one.py mcf > /dev/null 2>&1
python profiles/summarize_launches.py gpurun_out/ev/r01_dram_launches_mcf_fast.csv 16 > gpurun_out/ev/r01_launches_mcf_fast_summary.txt 2>&1
# full sections: the heavy kernels, a few launches each
VBK_LOOKAHEAD=0 ncu --set full --import-source on --clock-control none -k regex:'k_dense_update_m' --launch-skip 0 -c 1 -f -o gpurun_out/ev/r01_full_update_m_mcf python profiles/fast_one.py mcf > /dev/null 2>&1
VBK_LOOKAHEAD=0 ncu --set full --import-source on --clock-control none -k regex:'k_panel_diag|k_panel_rows|k_window_tri3|k_schur_window2|k_sparse_level_heavy' --launch-skip 40 -c 6 -f -o gpurun_out/ev/r01_full_chain_dfl001 python profiles/fast_one.py dfl001 > /dev/null 2>&1
for f in r01_full_update_m_mcf r01_full_chain_dfl001; do
ncu -i gpurun_out/ev/$f.ncu-rep --page raw --csv > gpurun_out/ev/${f}_raw.csv 2>/dev/null
ncu -i gpurun_out/ev/$f.ncu-rep --page details --csv > gpurun_out/ev/${f}_details.csv 2>/dev/null
done
VBK_PROF=1 VBK_LOOKAHEAD=0 python profiles/fast_one.py dfl001 2>&1 | grep -i "profile" | tail -1 > gpurun_out/ev/r01_panel_cycles_dfl001.txt
python bench.py --workload batch --batch-per-gpu 8 --steps 1 > gpurun_out/ev/r01_bench_batch_1gpu.json 2> gpurun_out/ev/batch.err; tail -2 gpurun_out/ev/batch.err
find gpurun_out -name "*.ncu-rep" -size +25M -delete
du -sh gpurun_out/ev
