set -x
timeout 900 python -m pytest tests/test_gpu.py -m gpu -x -q -k "batch or rowblock or blend or fast_mode_full" 2>&1 | tail -5
python bench.py --workload rowblock --steps 10 --warmup 3 > gpurun_out/s2_rowblock_n1.json 2> gpurun_out/s2_rowblock_n1.err; tail -3 gpurun_out/s2_rowblock_n1.err; cat gpurun_out/s2_rowblock_n1.json
python bench.py --workload batch --steps 2 --warmup 1 --batch-per-gpu 8 --streams 4 > gpurun_out/s2_batch_n1.json 2> gpurun_out/s2_batch_n1.err; tail -3 gpurun_out/s2_batch_n1.err; cat gpurun_out/s2_batch_n1.json
python bench.py --workload batch --steps 1 --warmup 1 --batch-per-gpu 8 --streams 1 > gpurun_out/s2_batch_n1_s1.json 2> gpurun_out/s2_batch_n1_s1.err; cat gpurun_out/s2_batch_n1_s1.json
python bench.py --workload batch --steps 1 --warmup 1 --batch-per-gpu 8 --streams 8 > gpurun_out/s2_batch_n1_s8.json 2> gpurun_out/s2_batch_n1_s8.err; cat gpurun_out/s2_batch_n1_s8.json
