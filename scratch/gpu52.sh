( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/s30_pytest_gpu.log 2>&1; tail -6 gpurun_out/s30_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for w in dfl001 pilot87; do
python bench.py --workload $w --no-strict --no-cpu-baseline > gpurun_out/s30_bench_$w.json 2> gpurun_out/s30_bench.err; tail -2 gpurun_out/s30_bench.err
python -c "
import json; d=json.load(open('gpurun_out/s30_bench_$w.json')); print('$w ms/step', d['ms_per_step'], 'factor ms', d['roofline']['kernel_ms'], 'value', d['value'], d['parity'])"
done
python bench.py --workload mcf --no-cpu-baseline --steps 5 > gpurun_out/s30_bench_mcf.json 2> gpurun_out/s30_bench_mcf.err; tail -2 gpurun_out/s30_bench_mcf.err
python -c "
import json; d=json.load(open('gpurun_out/s30_bench_mcf.json')); print('mcf ms/step', d['ms_per_step'], 'factor ms', d['roofline']['kernel_ms'], 'GFLOP/s', d['value'], d['roofline']['frac'], d['parity'])"
VBK_PROF=1 VBK_LOOKAHEAD=0 python profiles/fast_one.py dfl001 2>&1 | grep -i "profile" | tail -1
