cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py -m gpu -x -q -k "loss or graphed or whole_path or full_size or full_batch_loss or autograd or reentrant or head or forward_train" > gpurun_out/r2c_tests18.txt 2>&1; tail -3 gpurun_out/r2c_tests18.txt
for w in cfg5 cfg2 cfg3; do
  LOSS_HINT=1 timeout 120 python scripts/loss_bench.py $w 200
  timeout 120 python scripts/loss_bench.py $w 200
done 2>&1 | grep "^{" > gpurun_out/r2c_loss_bench18.txt
python - <<'PY'
import json
for l in open("gpurun_out/r2c_loss_bench18.txt"):
    d=json.loads(l); print(d["workload"], d["env"], round(d["us"],2), round(d["frac_of_6553.9"],3))
PY
python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/r2c_bench18.json 2>gpurun_out/r2c_bench18.err; tail -2 gpurun_out/r2c_bench18.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2c_bench18.json").read().strip().splitlines()[-1])
print(round(d["value"]), d["ms_per_step"], d["run"]["ms_per_step_one_in_flight"], {k:round(v,2) for k,v in d["stage_us"].items()}, d["roofline"]["frac"], d["roofline_large"]["loss_fwd_bwd"]["frac"], d["other_configs"]["cfg3"]["loss_fwd_bwd"]["frac"])
PY
