for rho in 0.15 0.35 0.5 0.7; do
VBK_WINDOW_RHO=$rho python bench.py --no-strict --no-cpu-baseline --steps 5 > gpurun_out/s22_dfl001_$rho.json 2> gpurun_out/s22.err; tail -1 gpurun_out/s22.err
python -c "
import json; d=json.load(open('gpurun_out/s22_dfl001_$rho.json')); print('rho $rho dfl001 ms/step', round(d['ms_per_step'],3), 'factor ms', round(d['roofline']['kernel_ms'],3), d['parity']['max_rel_err'])"
done
for rho in 0.15 0.4 0.6; do
VBK_WINDOW_RHO=$rho python bench.py --workload mcf:26:16 --no-cpu-baseline --steps 5 > gpurun_out/s22_mcf_$rho.json 2> gpurun_out/s22.err; tail -1 gpurun_out/s22.err
python -c "
import json; d=json.load(open('gpurun_out/s22_mcf_$rho.json')); print('rho $rho mcf:26:16 ms/step', round(d['ms_per_step'],3), 'factor ms', round(d['roofline']['kernel_ms'],3), d['parity'])"
done
