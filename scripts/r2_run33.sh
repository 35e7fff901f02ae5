cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests_y.txt 2>&1; tail -3 gpurun_out/r2_tests_y.txt
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_o.json 2>gpurun_out/r2_bench_o.err; tail -2 gpurun_out/r2_bench_o.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_o_ref.json 2>gpurun_out/r2_bench_o_ref.err; tail -2 gpurun_out/r2_bench_o_ref.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench_o.json").read().strip().splitlines()[-1])
print(round(d["value"]), d["ms_per_step"], d["e2e"], d["cpu_baseline"], d["roofline"], d["gpu_launches"], d["clocks"])
print({k:(v.get("loss_fwd_bwd"), v.get("assign",{}).get("us")) for k,v in [("cfg5",d["roofline_large"])]+list(d["other_configs"].items())})
r=json.loads(open("gpurun_out/r2_bench_o_ref.json").read().strip().splitlines()[-1]); print(r["value"], r["cpu_baseline"])
PY
