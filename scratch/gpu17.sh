VBK_PROF=1 python profiles/fast_one.py dfl001 2>&1 | grep "panel profile" | tail -1
python bench.py --no-strict --no-cpu-baseline > gpurun_out/s5_bench_dfl001.json 2> gpurun_out/s5_bench.err
python -c "
import json; d=json.load(open('gpurun_out/s5_bench_dfl001.json')); print('dfl001 ms/step', d['ms_per_step'], 'factor ms', d['roofline']['kernel_ms'], 'value', d['value'], d['parity'])"
timeout 600 python -m pytest tests/test_gpu.py -m gpu -x -q -k "fast_mode or batch" 2>&1 | tail -3
