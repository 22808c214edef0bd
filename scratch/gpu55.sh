timeout 900 python -m pytest tests/test_gpu.py -m gpu -x -q -k "fast_mode or batch or residual or multicommodity" 2>&1 | tail -3
for w in dfl001 pilot87; do
python bench.py --workload $w --no-strict --no-cpu-baseline > gpurun_out/s32_bench_$w.json 2> gpurun_out/s32_bench.err; tail -2 gpurun_out/s32_bench.err
python -c "
import json; d=json.load(open('gpurun_out/s32_bench_$w.json')); print('$w ms/step', d['ms_per_step'], 'factor ms', d['roofline']['kernel_ms'], 'value', d['value'], d['parity'])"
done
python bench.py --workload mcf --no-cpu-baseline --steps 5 > gpurun_out/s32_bench_mcf.json 2> gpurun_out/s32_bench_mcf.err; tail -2 gpurun_out/s32_bench_mcf.err
python -c "
import json; d=json.load(open('gpurun_out/s32_bench_mcf.json')); print('mcf ms/step', d['ms_per_step'], 'factor ms', d['roofline']['kernel_ms'], 'GFLOP/s', d['value'], d['roofline']['frac'], d['parity'])"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_window_gather|k_fwd_flags|k_bwd_flags|k_window_tri3' -c 40 --csv --log-file gpurun_out/s32_solve.csv python profiles/fast_one.py dfl001 > /dev/null 2>&1
python profiles/summarize_launches.py gpurun_out/s32_solve.csv 6
