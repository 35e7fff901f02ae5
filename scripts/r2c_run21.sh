cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_tests_head.txt 2>&1; tail -2 gpurun_out/r2c_tests_head.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2c_head_n1.json 2> gpurun_out/r2c_head_n1.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2c_head_n1.json").read().strip().splitlines()[-1])
print(round(d["value"],1), d.get("ms_per_step"), d["e2e"]["value"], d["roofline"]["frac"], d["roofline_large"]["loss_fwd_bwd"]["frac"], d["run"]["ms_per_step_one_in_flight"], d["gpu_launches"], d["clocks"])
PY
