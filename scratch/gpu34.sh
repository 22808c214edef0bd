python bench.py --no-strict --no-cpu-baseline --steps 5 > gpurun_out/s16_bench_dfl001.json 2> gpurun_out/s16_bench.err; tail -2 gpurun_out/s16_bench.err
python -c "
import json; d=json.load(open('gpurun_out/s16_bench_dfl001.json')); print('dfl001 ms/step', d['ms_per_step'], 'factor ms', d['roofline']['kernel_ms'], 'value', d['value'], d['roofline']['bound'], d['roofline']['frac'], d['parity'])"
for w in mcf:26:16 mcf; do
( time python bench.py --workload $w --no-cpu-baseline --steps 5 ) > gpurun_out/s16_bench_$w.json 2> gpurun_out/s16_bench_$w.err; tail -4 gpurun_out/s16_bench_$w.err
python -c "
import json; d=json.load(open('gpurun_out/s16_bench_$w.json')); print('$w', d['config']['workload']); print(' ms/step', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], 'factor ms', d['roofline']['kernel_ms'], 'GFLOP/s', d['value'], d['roofline']['bound'], d['roofline']['frac'], d['parity'], d['symbolic'])"
done
VBK_LOOKAHEAD=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/s16_launches_mcf.csv python profiles/fast_one.py mcf > gpurun_out/s16_ncu.log 2>&1
python profiles/summarize_launches.py gpurun_out/s16_launches_mcf.csv 12
