TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
$TR bench.py --gpus 2 --workload rowblock --steps 10 --warmup 3 > gpurun_out/s6_rowblock_n2.json 2> gpurun_out/s6_rowblock_n2.err; tail -2 gpurun_out/s6_rowblock_n2.err; cat gpurun_out/s6_rowblock_n2.json
$TR bench.py --gpus 2 --workload batch --steps 1 --warmup 1 --batch-per-gpu 8 --streams 8 > gpurun_out/s6_batch_n2.json 2> gpurun_out/s6_batch_n2.err; tail -2 gpurun_out/s6_batch_n2.err; cat gpurun_out/s6_batch_n2.json
python bench.py --workload batch --steps 1 --warmup 1 --batch-per-gpu 8 --streams 8 > gpurun_out/s6_batch_n1.json 2> gpurun_out/s6_batch_n1.err; cat gpurun_out/s6_batch_n1.json
$TR bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/s6_default_n2.json 2> gpurun_out/s6_default_n2.err; tail -2 gpurun_out/s6_default_n2.err; cat gpurun_out/s6_default_n2.json | cut -c1-400
