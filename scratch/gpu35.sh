timeout 600 python -m pytest tests/test_gpu.py -m gpu -x -q -k "fast_mode or batch or residual" 2>&1 | tail -3
python bench.py --no-strict --no-cpu-baseline > gpurun_out/s17_bench_dfl001.json 2> gpurun_out/s17_bench.err; tail -2 gpurun_out/s17_bench.err
python -c "
import json; d=json.load(open('gpurun_out/s17_bench_dfl001.json')); print('dfl001 ms/step', d['ms_per_step'], 'factor ms', d['roofline']['kernel_ms'], 'value', d['value'], d['parity'])"
VBK_WSOLVE=v2 python bench.py --no-strict --no-cpu-baseline > gpurun_out/s17_bench_dfl001_v2.json 2> gpurun_out/s17_bench.err; tail -2 gpurun_out/s17_bench.err
python -c "
import json; d=json.load(open('gpurun_out/s17_bench_dfl001_v2.json')); print('old tri: dfl001 ms/step', d['ms_per_step'], 'factor ms', d['roofline']['kernel_ms'], 'value', d['value'], d['parity'])"
VBK_LOOKAHEAD=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/s17_launches.csv python profiles/fast_one.py dfl001 > /dev/null 2>&1
python profiles/summarize_launches.py gpurun_out/s17_launches.csv 12
