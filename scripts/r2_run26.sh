cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests_s.txt 2>&1; tail -5 gpurun_out/r2_tests_s.txt
python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/r2_bench_k.json 2>gpurun_out/r2_bench_k.err; tail -3 gpurun_out/r2_bench_k.err
python bench.py --no-side-configs --no-e2e --no-cpu-baseline > gpurun_out/r2_bench_k_long.json 2>gpurun_out/r2_bench_k_long.err
python bench.py --no-side-configs --no-e2e --no-cpu-baseline --no-weight-sums > gpurun_out/r2_bench_k_long_nows.json 2>gpurun_out/r2_bench_k_long_nows.err
python - <<'PY'
import json
for f in ("r2_bench_k","r2_bench_k_long","r2_bench_k_long_nows"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["run"]["ms_per_step_one_in_flight"], d["stage_us"], d["roofline"]["frac"])
        if d.get("roofline_large"): print(json.dumps(d["roofline_large"])[:1500])
        if d.get("other_configs"): print(json.dumps(d["other_configs"])[:2500])
    except Exception as e: print(f, "ERR", e)
PY
