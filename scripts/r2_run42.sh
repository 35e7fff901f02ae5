cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py -m gpu -x -q -k "loss or graphed or whole_path or full_size or full_batch_loss or autograd or reentrant" > gpurun_out/r2_tests_ab.txt 2>&1; tail -3 gpurun_out/r2_tests_ab.txt
for w in cfg5 cfg2 cfg3; do
  LOSS_HINT=1 timeout 120 python scripts/loss_bench.py $w 200
  timeout 120 python scripts/loss_bench.py $w 200
done 2>&1 | grep "^{" > gpurun_out/r2_loss_bench22.txt
python - <<'PY'
import json
for l in open("gpurun_out/r2_loss_bench22.txt"):
    d=json.loads(l); print(d["workload"], d["env"], round(d["us"],2), round(d["frac_of_6553.9"],3))
PY
LOSS_HINT=1 python scripts/dense_timeline.py cfg5 2>&1 | grep -v "Warning\|q = lambda\|_nanquantile\|per item\|setup->\|stage0->\|stage7->\| box planes n\| exit n\|slowest\|SM" | tee gpurun_out/r2_dense_timeline9.txt
python bench.py --no-side-configs --no-e2e --no-cpu-baseline > gpurun_out/r2_bench_t.json 2>gpurun_out/r2_bench_t.err; tail -2 gpurun_out/r2_bench_t.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench_t.json").read().strip().splitlines()[-1])
print(round(d["value"]), d["ms_per_step"], d["run"]["ms_per_step_one_in_flight"], {k:round(v,2) for k,v in d["stage_us"].items()}, d["roofline"]["frac"])
PY
