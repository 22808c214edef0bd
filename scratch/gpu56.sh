mkdir -p gpurun_out/ev3
python bench.py > gpurun_out/ev3/r01_bench_dfl001.json 2> gpurun_out/ev3/bench.err; tail -2 gpurun_out/ev3/bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/ev3/r01_bench_dfl001_reference.json 2>/dev/null
python bench.py --workload pilot87 --cpu-budget 5 > gpurun_out/ev3/r01_bench_pilot87.json 2>/dev/null
python bench.py --workload mcf --steps 5 --no-cpu-baseline > gpurun_out/ev3/r01_bench_mcf_nocpu.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/ev3/r01_ncu_launches_bench_dfl001_fast.csv python bench.py --steps 2 --warmup 3 --no-strict --no-cpu-baseline > gpurun_out/ev3/ncu1.log 2>&1
python profiles/summarize_launches.py gpurun_out/ev3/r01_ncu_launches_bench_dfl001_fast.csv 24 > gpurun_out/ev3/r01_ncu_launches_bench_dfl001_fast_summary.txt
cat gpurun_out/ev3/r01_ncu_launches_bench_dfl001_fast_summary.txt | head -16
VBK_PROF=1 VBK_LOOKAHEAD=0 python profiles/fast_one.py dfl001 2>&1 | grep -i "profile" | tail -1 > gpurun_out/ev3/r01_panel_cycles_dfl001.txt
for f in gpurun_out/ev3/*.json; do python -c "
import json,sys; d=json.load(open('$f')); print('$f', d['value'], d['unit'], 'ms/step', d['ms_per_step'], 'e2e', d.get('e2e',{}).get('value'))"; done
