set -x
timeout 600 python -m pytest tests/test_gpu.py -m gpu -x -q -k "fast_mode or batch" 2>&1 | tail -3
python bench.py --no-strict --no-cpu-baseline > gpurun_out/s7_bench_dfl001.json 2> gpurun_out/s7_bench.err
python -c "
import json; d=json.load(open('gpurun_out/s7_bench_dfl001.json')); print('LOOKAHEAD dfl001 ms/step', d['ms_per_step'], 'factor ms', d['roofline']['kernel_ms'], 'value', d['value'], d['parity'])"
VBK_LOOKAHEAD=0 python bench.py --no-strict --no-cpu-baseline > gpurun_out/s7_bench_dfl001_nola.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/s7_bench_dfl001_nola.json')); print('NO-LOOKAHEAD dfl001 ms/step', d['ms_per_step'], 'factor ms', d['roofline']['kernel_ms'])"
python bench.py --workload pilot87 --no-strict --no-cpu-baseline > gpurun_out/s7_bench_pilot87.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/s7_bench_pilot87.json')); print('pilot87 ms/step', d['ms_per_step'], 'factor ms', d['roofline']['kernel_ms'], d['parity'])"
VBK_PROF=1 VBK_LOOKAHEAD=0 python profiles/fast_one.py dfl001 2>&1 | grep "panel profile" | tail -1
# per-launch DRAM traffic + duration of every kernel of 2 factorisations + 2 solves (cheap metrics-only pass)
VBK_LOOKAHEAD=0 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r01_dram_launches_dfl001_fast.csv python profiles/fast_one.py dfl001 > /dev/null 2>&1
# full sections for a few representative launches of the heavy kernels (kept small)
VBK_LOOKAHEAD=0 ncu --set full --import-source on --clock-control none -k regex:'k_dense_update_k|k_panel_diag|k_panel_rows' --launch-skip 6 -c 3 -f -o gpurun_out/r01_full_panel_dfl001 python profiles/fast_one.py dfl001 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
ncu -i gpurun_out/r01_full_panel_dfl001.ncu-rep --page raw --csv > gpurun_out/r01_full_panel_dfl001_raw.csv 2>/dev/null
ncu -i gpurun_out/r01_full_panel_dfl001.ncu-rep --page details --csv > gpurun_out/r01_full_panel_dfl001_details.csv 2>/dev/null
find gpurun_out -name "*.ncu-rep" -size +20M -delete
du -sh gpurun_out
