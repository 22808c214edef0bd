( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 ) > gpurun_out/s25_pytest_gpu.log 2>&1; tail -16 gpurun_out/s25_pytest_gpu.log
for w in dfl001 pilot87; do
python bench.py --workload $w --no-strict --no-cpu-baseline > gpurun_out/s25_bench_$w.json 2> gpurun_out/s25_bench.err; tail -2 gpurun_out/s25_bench.err
python -c "
import json; d=json.load(open('gpurun_out/s25_bench_$w.json')); print('$w ms/step', d['ms_per_step'], 'factor ms', d['roofline']['kernel_ms'], 'value', d['value'], d['parity'])"
done
