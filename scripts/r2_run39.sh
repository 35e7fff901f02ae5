cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py -m gpu -x -q -k "loss or graphed or whole_path or full_size or full_batch_loss or autograd or reentrant" > gpurun_out/r2_tests_aa.txt 2>&1; tail -3 gpurun_out/r2_tests_aa.txt
for w in cfg5 cfg2 cfg3; do
  for ch in 1 0; do
  LOSS_HINT=1 RADET_LOSS_CHAIN=$ch timeout 120 python scripts/loss_bench.py $w 200
  RADET_LOSS_CHAIN=$ch timeout 120 python scripts/loss_bench.py $w 200
  done
done 2>&1 | grep "^{" > gpurun_out/r2_loss_bench21.txt
python - <<'PY'
import json
for l in open("gpurun_out/r2_loss_bench21.txt"):
    d=json.loads(l); print(d["workload"], d["env"], round(d["us"],2), round(d["frac_of_6553.9"],3))
PY
for ch in 1 0; do
RADET_LOSS_CHAIN=$ch python bench.py --no-side-configs --no-e2e --no-cpu-baseline > gpurun_out/r2_bench_s_$ch.json 2>gpurun_out/r2_bench_s_$ch.err; tail -2 gpurun_out/r2_bench_s_$ch.err
RADET_LOSS_CHAIN=$ch python bench.py --steps 20 --warmup 5 --no-side-configs --no-e2e --no-cpu-baseline > gpurun_out/r2_bench_s20_$ch.json 2>gpurun_out/r2_bench_s_$ch.err
done
python - <<'PY'
import json
for ch in (1,0):
  for f in ("r2_bench_s_","r2_bench_s20_"):
    d=json.loads(open(f"gpurun_out/{f}{ch}.json").read().strip().splitlines()[-1])
    print(ch, f, round(d["value"]), d["ms_per_step"], d["run"]["ms_per_step_one_in_flight"], {k:round(v,2) for k,v in d["stage_us"].items()})
PY
