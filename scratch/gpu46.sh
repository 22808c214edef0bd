timeout 900 python -m pytest tests/test_gpu.py -m gpu -x -q -k "fast_mode or batch or residual" 2>&1 | tail -3
for w in dfl001 pilot87; do
python bench.py --workload $w --no-strict --no-cpu-baseline > gpurun_out/s26_bench_$w.json 2> gpurun_out/s26_bench.err; tail -2 gpurun_out/s26_bench.err
python -c "
import json; d=json.load(open('gpurun_out/s26_bench_$w.json')); print('$w ms/step', d['ms_per_step'], 'factor ms', d['roofline']['kernel_ms'], 'value', d['value'], d['parity'])"
done
VBK_ROWS=dfma python bench.py --no-strict --no-cpu-baseline > gpurun_out/s26_bench_dfma.json 2> gpurun_out/s26_bench.err; tail -2 gpurun_out/s26_bench.err
python -c "
import json; d=json.load(open('gpurun_out/s26_bench_dfma.json')); print('rows=dfma dfl001 ms/step', d['ms_per_step'], 'factor ms', d['roofline']['kernel_ms'], 'value', d['value'], d['parity'])"
VBK_PROF=1 VBK_LOOKAHEAD=0 python profiles/fast_one.py dfl001 2>&1 | grep -i "profile" | tail -1
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/s26_launches_dfl001.csv python profiles/fast_one.py dfl001 > gpurun_out/s26_ncu.log 2>&1
python profiles/summarize_launches.py gpurun_out/s26_launches_dfl001.csv 8
