timeout 600 python -m pytest tests/test_gpu.py -m gpu -x -q -k "fast_mode or batch" 2>&1 | tail -2
python bench.py --no-strict --no-cpu-baseline > gpurun_out/s8_bench_dfl001.json 2> gpurun_out/s8_bench.err
python -c "
import json; d=json.load(open('gpurun_out/s8_bench_dfl001.json')); print('LOOKAHEAD dfl001 ms/step', d['ms_per_step'], 'factor ms', d['roofline']['kernel_ms'], 'value', d['value'], d['parity'])"
VBK_LOOKAHEAD=0 python bench.py --no-strict --no-cpu-baseline > gpurun_out/s8_bench_dfl001_nola.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/s8_bench_dfl001_nola.json')); print('NO-LOOKAHEAD dfl001 ms/step', d['ms_per_step'], 'factor ms', d['roofline']['kernel_ms'])"
python bench.py --workload pilot87 --no-strict --no-cpu-baseline > gpurun_out/s8_bench_pilot87.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/s8_bench_pilot87.json')); print('pilot87 ms/step', d['ms_per_step'], 'factor ms', d['roofline']['kernel_ms'], d['parity'])"
VBK_PROF=1 VBK_LOOKAHEAD=0 python profiles/fast_one.py dfl001 2>&1 | grep "panel profile" | tail -1
