timeout 600 python -m pytest tests/test_gpu.py -m gpu -x -q -k "fast_mode or batch or residual" 2>&1 | tail -2
for u in p k; do
VBK_UPDATE=$u python bench.py --no-strict --no-cpu-baseline > gpurun_out/s10_bench_dfl001_$u.json 2> gpurun_out/s10_bench.err
python -c "
import json; d=json.load(open('gpurun_out/s10_bench_dfl001_$u.json')); print('update=$u dfl001 ms/step', d['ms_per_step'], 'factor ms', d['roofline']['kernel_ms'], 'value', d['value'], d['parity'])"
done
VBK_LOOKAHEAD=0 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_dense_update' -c 80 --csv --log-file gpurun_out/s10_upd_launches.csv python profiles/fast_one.py dfl001 > /dev/null 2>&1
python profiles/summarize_launches.py gpurun_out/s10_upd_launches.csv 5 vbk::k_dense_update_p
