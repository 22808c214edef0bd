timeout 900 python -m pytest tests/test_gpu.py -m gpu -x -q -k "fast_mode or batch or residual" 2>&1 | tail -3
python bench.py --no-strict --no-cpu-baseline > gpurun_out/s21_bench_dfl001.json 2> gpurun_out/s21_bench.err; tail -2 gpurun_out/s21_bench.err
python -c "
import json; d=json.load(open('gpurun_out/s21_bench_dfl001.json')); print('dfl001 ms/step', d['ms_per_step'], 'factor ms', d['roofline']['kernel_ms'], 'value', d['value'], d['parity'])"
python bench.py --workload mcf --no-cpu-baseline --steps 5 > gpurun_out/s21_bench_mcf.json 2> gpurun_out/s21_bench_mcf.err; tail -4 gpurun_out/s21_bench_mcf.err
python -c "
import json; d=json.load(open('gpurun_out/s21_bench_mcf.json')); print('mcf ms/step', d['ms_per_step'], 'factor ms', d['roofline']['kernel_ms'], 'GFLOP/s', d['value'], d['roofline']['frac'], d['parity'])"
VBK_LOOKAHEAD=0 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_dense_update_m' -c 70 --csv --log-file gpurun_out/s21_upd_mcf.csv python profiles/fast_one.py mcf > gpurun_out/s21_ncu.log 2>&1
python profiles/summarize_launches.py gpurun_out/s21_upd_mcf.csv 3 "void k_dense_update_m<64>"
