#!/bin/bash
# Round-2 final measurements on one B200 (profiles/r02_summary.md quotes these files):
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash profiles/r02_final.sh'
set -u
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -q 2>&1 | tail -4 > $O/r02_pytest_gpu.log
python bench.py --impl reference --steps 3 --warmup 1 2>$O/r02_ref.err | tail -1 > $O/r02_bench_dfl001_reference.json
python bench.py 2>$O/r02_bench.err | tail -1 > $O/r02_bench_dfl001.json
python bench.py --workload pilot87 --no-solve-time 2>>$O/r02_bench.err | tail -1 > $O/r02_bench_pilot87.json
python bench.py --mode fast --no-solve-time 2>>$O/r02_bench.err | tail -1 > $O/r02_bench_dfl001_fast.json
python bench.py --workload mcf --mode fast --no-solve-time --no-strict --no-cpu-baseline 2>>$O/r02_bench.err | tail -1 > $O/r02_bench_mcf_32_25_fast.json
python profiles/strict_sweep.py > $O/r02_strict_sweep.jsonl 2>$O/r02_strict_sweep.err
tail -3 $O/r02_pytest_gpu.log; for f in $O/r02_bench_dfl001.json $O/r02_bench_dfl001_reference.json $O/r02_bench_pilot87.json $O/r02_bench_dfl001_fast.json $O/r02_bench_mcf_32_25_fast.json; do echo "$f: $(cut -c1-230 $f)"; done
wc -l $O/r02_strict_sweep.jsonl; tail -2 $O/r02_bench.err
