( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r01_pytest_gpu_final.log 2>&1; tail -4 gpurun_out/r01_pytest_gpu_final.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 5 --no-strict --cpu-budget 4 > gpurun_out/r01_bench_final_quick.json 2> gpurun_out/final.err; tail -1 gpurun_out/final.err; python -c "
import json; d=json.load(open('gpurun_out/r01_bench_final_quick.json')); print(d['value'], d['ms_per_step'], d['e2e'], d['roofline']['frac'], d['gpu_launches'], d['clocks'])"
