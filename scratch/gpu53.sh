timeout 900 python -m pytest tests/test_gpu.py -m gpu -x -q -k "fast_mode or batch or residual" 2>&1 | tail -3
for w in dfl001 pilot87; do
python bench.py --workload $w --no-strict --no-cpu-baseline > gpurun_out/s31_bench_$w.json 2> gpurun_out/s31_bench.err; tail -2 gpurun_out/s31_bench.err
python -c "
import json; d=json.load(open('gpurun_out/s31_bench_$w.json')); print('$w ms/step', d['ms_per_step'], 'factor ms', d['roofline']['kernel_ms'], 'value', d['value'], d['parity'])"
done
VBK_SPLIT=0 python bench.py --no-strict --no-cpu-baseline > gpurun_out/s31_bench_nosplit.json 2> gpurun_out/s31_bench.err; tail -2 gpurun_out/s31_bench.err
python -c "
import json; d=json.load(open('gpurun_out/s31_bench_nosplit.json')); print('split=0 dfl001 ms/step', d['ms_per_step'], 'factor ms', d['roofline']['kernel_ms'], 'value', d['value'], d['parity'])"
python bench.py --workload mcf --no-cpu-baseline --steps 5 > gpurun_out/s31_bench_mcf.json 2> gpurun_out/s31_bench_mcf.err; tail -2 gpurun_out/s31_bench_mcf.err
python -c "
import json; d=json.load(open('gpurun_out/s31_bench_mcf.json')); print('mcf ms/step', d['ms_per_step'], 'factor ms', d['roofline']['kernel_ms'], 'GFLOP/s', d['value'], d['roofline']['frac'], d['parity'])"
