timeout 900 python -m pytest tests/test_gpu.py -m gpu -x -q -k "fast_mode or batch or residual" 2>&1 | tail -3
python bench.py --no-strict --no-cpu-baseline > gpurun_out/s27_bench_dfl001.json 2> gpurun_out/s27_bench.err; tail -2 gpurun_out/s27_bench.err
python -c "
import json; d=json.load(open('gpurun_out/s27_bench_dfl001.json')); print('dfl001 ms/step', d['ms_per_step'], 'factor ms', d['roofline']['kernel_ms'], 'value', d['value'], d['parity'], d['roofline']['traffic'])"
VBK_PROF=1 VBK_LOOKAHEAD=0 python profiles/fast_one.py dfl001 2>&1 | grep -i "profile" | tail -1
