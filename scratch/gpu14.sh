timeout 600 python -m pytest tests/test_gpu.py -m gpu -x -q -k "fast_mode or batch or strict_mode_big" 2>&1 | tail -3
python bench.py --no-strict --no-cpu-baseline > gpurun_out/s4_bench_dfl001.json 2> gpurun_out/s4_bench.err
python -c "
import json; d=json.load(open('gpurun_out/s4_bench_dfl001.json')); print('dfl001 ms/step', d['ms_per_step'], 'factor ms', d['roofline']['kernel_ms'], 'value', d['value'], d['parity'])"
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 3000 --csv --log-file gpurun_out/s4_launches_dfl001_fast_warm.csv python profiles/fast_one.py dfl001 > gpurun_out/s4_ncu.log 2>&1
python profiles/summarize_launches.py gpurun_out/s4_launches_dfl001_fast_warm.csv 10 k_panel_diag k_panel_rows k_dense_update_k
