cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py -m gpu -x -q -k "loss or graphed or whole_path or full_size or full_batch_loss" > gpurun_out/r2_tests_x.txt 2>&1; tail -3 gpurun_out/r2_tests_x.txt
for w in cfg5 cfg2 cfg3; do
  LOSS_HINT=1 timeout 120 python scripts/loss_bench.py $w 200
  LOSS_HINT=1 RADET_DENSE_NOBALANCE=1 timeout 120 python scripts/loss_bench.py $w 200
done 2>&1 | grep "^{" > gpurun_out/r2_loss_bench20.txt
python - <<'PY'
import json
for l in open("gpurun_out/r2_loss_bench20.txt"):
    d=json.loads(l); print(d["workload"], d["env"], round(d["us"],2), round(d["frac_of_6553.9"],3))
PY
LOSS_HINT=1 python scripts/dense_timeline.py cfg5 2>&1 | grep -v "per item\|setup->\|stage0->\|stage7->\| box planes n\| exit n\|Warning\|q = lambda\|_nanquantile"
