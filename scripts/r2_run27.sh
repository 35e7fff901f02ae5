cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py -m gpu -x -q -k "loss or graphed or whole_path or full_size or full_batch_loss" > gpurun_out/r2_tests_t.txt 2>&1; tail -3 gpurun_out/r2_tests_t.txt
for w in cfg5 cfg2 cfg3; do
  for ch in 0 2 3 5; do
  LOSS_HINT=1 RADET_DENSE_CHUNKS=$ch timeout 120 python scripts/loss_bench.py $w 100
  done
  RADET_DENSE_CHUNKS=0 timeout 120 python scripts/loss_bench.py $w 100
done 2>&1 | grep "^{" > gpurun_out/r2_loss_bench16.txt
python - <<'PY'
import json
for l in open("gpurun_out/r2_loss_bench16.txt"):
    d=json.loads(l); print(d["workload"], d["env"], round(d["us"],2), round(d["frac_of_6553.9"],3))
PY
