mkdir -p gpurun_out/ev4
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/ev4/r01_pytest_gpu.log 2>&1; tail -4 gpurun_out/ev4/r01_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/ev4/r01_bench_dfl001.json 2> gpurun_out/ev4/bench.err; tail -2 gpurun_out/ev4/bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/ev4/r01_bench_dfl001_reference.json 2>/dev/null
for f in gpurun_out/ev4/*.json; do python -c "
import json,sys; d=json.load(open('$f')); print('$f', d['value'], d['unit'], 'ms/step', d['ms_per_step'], 'e2e', d.get('e2e',{}).get('value'), d.get('roofline',{}).get('frac'))"; done
