cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -k "assign or full_size or full_batch or graphed or whole_path or forward_train or label_assignment or pipeline" > gpurun_out/r2_tests_z.txt 2>&1; tail -3 gpurun_out/r2_tests_z.txt
timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python scripts/sanitize_driver.py 2>&1 | grep -E "RACECHECK SUMMARY|hazard|sanitize driver done" | sort | uniq -c | head
python bench.py --no-side-configs --no-e2e --no-cpu-baseline > gpurun_out/r2_bench_r.json 2>gpurun_out/r2_bench_r.err
python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/r2_bench_r20.json 2>gpurun_out/r2_bench_r.err
python - <<'PY'
import json
for f in ("r2_bench_r","r2_bench_r20"):
    d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    print(f, round(d["value"]), d["ms_per_step"], d["run"]["ms_per_step_one_in_flight"], {k:round(v,2) for k,v in d["stage_us"].items()})
    if d.get("roofline_large"): print("cfg5 assign us", d["roofline_large"]["assign"]["us"], "cfg3", d["other_configs"]["cfg3"]["assign"]["us"])
PY
python profiles/phase_timing.py 2>&1 | tail -12
