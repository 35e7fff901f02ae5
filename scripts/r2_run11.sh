cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "loss or graphed or whole_path or full_size_cfg5" > gpurun_out/r2_tests_h.txt 2>&1; tail -5 gpurun_out/r2_tests_h.txt
for c in 1 296 600 1184; do
 for w in cfg2 cfg3 cfg5; do RADET_DENSE_CTAS=$c timeout 120 python scripts/loss_bench.py $w 100; done
 RADET_DENSE_CTAS=$c python bench.py --steps 20 --warmup 5 --no-side-configs --no-e2e --no-cpu-baseline > gpurun_out/r2_bench_c$c.json 2>gpurun_out/r2_bench_c$c.err
 RADET_DENSE_CTAS=$c python bench.py --no-side-configs --no-e2e --no-cpu-baseline > gpurun_out/r2_bench_c${c}_long.json 2>gpurun_out/r2_bench_c${c}_long.err
done 2>&1 | grep "^{" > gpurun_out/r2_loss_bench7.txt
