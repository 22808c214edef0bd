for rho in 0.1 0.06 0.03; do
VBK_WINDOW_RHO=$rho python bench.py --no-strict --no-cpu-baseline --steps 5 > gpurun_out/s23_dfl001_$rho.json 2> gpurun_out/s23.err; tail -1 gpurun_out/s23.err
python -c "
import json; d=json.load(open('gpurun_out/s23_dfl001_$rho.json')); print('rho $rho dfl001 ms/step', round(d['ms_per_step'],3), 'factor ms', round(d['roofline']['kernel_ms'],3), d['parity']['max_rel_err'])"
VBK_WINDOW_RHO=$rho python bench.py --workload pilot87 --no-strict --no-cpu-baseline --steps 5 > gpurun_out/s23_pilot87_$rho.json 2> gpurun_out/s23.err; tail -1 gpurun_out/s23.err
python -c "
import json; d=json.load(open('gpurun_out/s23_pilot87_$rho.json')); print('rho $rho pilot87 ms/step', round(d['ms_per_step'],3), 'factor ms', round(d['roofline']['kernel_ms'],3), d['parity']['max_rel_err'])"
done
for rho in 0.1 0.06; do
VBK_WINDOW_RHO=$rho python bench.py --workload mcf:26:16 --no-cpu-baseline --steps 5 > gpurun_out/s23_mcf_$rho.json 2> gpurun_out/s23.err; tail -1 gpurun_out/s23.err
python -c "
import json; d=json.load(open('gpurun_out/s23_mcf_$rho.json')); print('rho $rho mcf:26:16 ms/step', round(d['ms_per_step'],3), 'factor ms', round(d['roofline']['kernel_ms'],3), d['parity'])"
done
VBK_WINDOW_RHO=0.1 python bench.py --workload mcf --no-cpu-baseline --steps 5 > gpurun_out/s23_mcf32_0.1.json 2> gpurun_out/s23.err; tail -1 gpurun_out/s23.err
python -c "
import json; d=json.load(open('gpurun_out/s23_mcf32_0.1.json')); print('rho 0.1 mcf R32 ms/step', round(d['ms_per_step'],3), 'factor ms', round(d['roofline']['kernel_ms'],3), d['value'], d['parity'])"
python bench.py --workload pilot87 --no-strict --no-cpu-baseline --steps 5 > gpurun_out/s23_pilot87_def.json 2> gpurun_out/s23.err
python -c "
import json; d=json.load(open('gpurun_out/s23_pilot87_def.json')); print('rho default pilot87 ms/step', round(d['ms_per_step'],3), 'factor ms', round(d['roofline']['kernel_ms'],3), d['parity']['max_rel_err'])"
