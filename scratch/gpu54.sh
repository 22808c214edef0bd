mkdir -p gpurun_out/ev2
python bench.py > gpurun_out/ev2/r01_bench_dfl001.json 2> gpurun_out/ev2/bench.err; tail -2 gpurun_out/ev2/bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/ev2/r01_bench_dfl001_reference.json 2>/dev/null
python bench.py --workload pilot87 --cpu-budget 5 > gpurun_out/ev2/r01_bench_pilot87.json 2>/dev/null
python bench.py --workload batch --steps 1 > gpurun_out/ev2/r01_bench_batch_1gpu.json 2> gpurun_out/ev2/batch.err; tail -2 gpurun_out/ev2/batch.err
python bench.py --workload rowblock --steps 10 > gpurun_out/ev2/r01_bench_rowblock_1gpu.json 2> gpurun_out/ev2/rowblock.err; tail -2 gpurun_out/ev2/rowblock.err
for f in gpurun_out/ev2/*.json; do python -c "
import json,sys; d=json.load(open('$f')); print('$f', d['value'], d['unit'], 'ms/step', d['ms_per_step'], d.get('parity'))"; done
