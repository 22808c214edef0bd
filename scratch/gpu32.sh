./build/ubench2
python bench.py --no-strict --no-cpu-baseline > gpurun_out/s15_bench_dfl001.json 2> gpurun_out/s15_bench.err
python -c "
import json; d=json.load(open('gpurun_out/s15_bench_dfl001.json')); print('dfl001 ms/step', d['ms_per_step'], 'factor ms', d['roofline']['kernel_ms'], 'value', d['value'], d['parity'])"
VBK_PROF=1 VBK_LOOKAHEAD=0 python profiles/fast_one.py dfl001 2>&1 | grep -i "profile" | tail -1
