VBK_LOOKAHEAD=0 ncu --set full --import-source on --clock-control none -k regex:'k_panel_diag' --launch-skip 3 -c 1 -f -o gpurun_out/r01_full_diag python profiles/fast_one.py dfl001 > /dev/null 2>&1
ncu -i gpurun_out/r01_full_diag.ncu-rep --page raw --csv > gpurun_out/r01_full_diag_raw.csv 2>/dev/null
ncu -i gpurun_out/r01_full_diag.ncu-rep --page source --csv > gpurun_out/r01_full_diag_source.csv 2>/dev/null
