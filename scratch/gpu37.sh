timeout 900 python -m pytest tests/test_gpu.py -m gpu -x -q -k "fast_mode or batch or residual" 2>&1 | tail -3
python bench.py --no-strict --no-cpu-baseline > gpurun_out/s19_bench_dfl001.json 2> gpurun_out/s19_bench.err; tail -2 gpurun_out/s19_bench.err
python -c "
import json; d=json.load(open('gpurun_out/s19_bench_dfl001.json')); print('dfl001 ms/step', d['ms_per_step'], 'factor ms', d['roofline']['kernel_ms'], 'value', d['value'], d['parity'])"
for w in mcf:26:16 mcf; do
python bench.py --workload $w --no-cpu-baseline --steps 5 > gpurun_out/s19_bench_$w.json 2> gpurun_out/s19_bench_$w.err; tail -4 gpurun_out/s19_bench_$w.err
python -c "
import json; d=json.load(open('gpurun_out/s19_bench_$w.json')); print('$w ms/step', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], 'factor ms', d['roofline']['kernel_ms'], 'GFLOP/s', d['value'], d['roofline']['bound'], d['roofline']['frac'], d['parity'])"
done
VBK_LOOKAHEAD=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/s19_launches_mcf.csv python profiles/fast_one.py mcf > gpurun_out/s19_ncu.log 2>&1
python profiles/summarize_launches.py gpurun_out/s19_launches_mcf.csv 14
VBK_LOOKAHEAD=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/s19_launches_dfl001.csv python profiles/fast_one.py dfl001 > gpurun_out/s19_ncu.log 2>&1
python profiles/summarize_launches.py gpurun_out/s19_launches_dfl001.csv 14 k_sparse_level k_sparse_level_heavy
