timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py > gpurun_out/s3_bench_dfl001.json 2> gpurun_out/s3_bench_dfl001.err; tail -2 gpurun_out/s3_bench_dfl001.err
python bench.py --workload pilot87 --no-strict > gpurun_out/s3_bench_pilot87.json 2> gpurun_out/s3_bench_pilot87.err
python - <<'PY'
import json
for n in ("dfl001","pilot87"):
    d=json.load(open(f"gpurun_out/s3_bench_{n}.json"))
    print(n, "value", round(d["value"],1), "ms/step", round(d["ms_per_step"],3), "factor ms", round(d["roofline"]["kernel_ms"],3), "fp64 frac", round(d["roofline"]["fp64"]["frac"],4), "e2e", round(d["e2e"]["value"],1), "cpu", d["cpu_baseline"]["ms_per_step"] if d["cpu_baseline"] else None, "parity", d["parity"])
PY
