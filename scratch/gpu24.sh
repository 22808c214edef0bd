./build/ubench
VBK_LOOKAHEAD=0 ncu --set full --import-source on --clock-control none -k regex:'k_dense_update_p' --launch-skip 0 -c 1 -f -o gpurun_out/r01_full_update_p python profiles/fast_one.py dfl001 > /dev/null 2>&1
ncu -i gpurun_out/r01_full_update_p.ncu-rep --page raw --csv > gpurun_out/r01_full_update_p_raw.csv 2>/dev/null
ncu -i gpurun_out/r01_full_update_p.ncu-rep --page details --csv > gpurun_out/r01_full_update_p_details.csv 2>/dev/null
ncu -i gpurun_out/r01_full_update_p.ncu-rep --page source --csv > gpurun_out/r01_full_update_p_source.csv 2>/dev/null
ls -la gpurun_out/*.ncu-rep
