VBK_UPDATE=m128 VBK_LOOKAHEAD=0 ncu --set full --import-source on --clock-control none -k regex:'k_dense_update_m' --launch-skip 0 -c 1 -f -o gpurun_out/r01_full_update_m128 python profiles/fast_one.py dfl001 > /dev/null 2>&1
ncu -i gpurun_out/r01_full_update_m128.ncu-rep --page raw --csv > gpurun_out/r01_full_update_m128_raw.csv 2>/dev/null
ncu -i gpurun_out/r01_full_update_m128.ncu-rep --page details --csv > gpurun_out/r01_full_update_m128_details.csv 2>/dev/null
ncu -i gpurun_out/r01_full_update_m128.ncu-rep --page source --csv > gpurun_out/r01_full_update_m128_source.csv 2>/dev/null
