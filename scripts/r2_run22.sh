cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py -m gpu -x -q -k "loss or graphed or whole_path or full_size or full_batch_loss" > gpurun_out/r2_tests_o.txt 2>&1; tail -4 gpurun_out/r2_tests_o.txt
for w in cfg5 cfg2 cfg3; do
  timeout 120 python scripts/loss_bench.py $w 100
  LOSS_HINT=1 timeout 120 python scripts/loss_bench.py $w 100
  RADET_LOSS_PDL=0 timeout 120 python scripts/loss_bench.py $w 100
done 2>&1 | grep "^{" > gpurun_out/r2_loss_bench11.txt
cat gpurun_out/r2_loss_bench11.txt
