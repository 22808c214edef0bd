( time timeout 1500 python profiles/fast_sweep.py ) > gpurun_out/r01_fast_sweep.jsonl 2> gpurun_out/r01_fast_sweep.err
tail -3 gpurun_out/r01_fast_sweep.err
tail -1 gpurun_out/r01_fast_sweep.jsonl | cut -c1-600
