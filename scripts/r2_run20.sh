cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py -m gpu -x -q -k "loss or graphed or whole_path or full_size or full_batch_loss" > gpurun_out/r2_tests_n.txt 2>&1; tail -4 gpurun_out/r2_tests_n.txt
for w in cfg5 cfg2 cfg3; do timeout 120 python scripts/loss_bench.py $w 100; RADET_LOSS_IMPL=tma timeout 120 python scripts/loss_bench.py $w 100; done 2>&1 | grep "^{" > gpurun_out/r2_loss_bench10.txt
python bench.py --steps 20 --warmup 5 --no-side-configs --no-e2e --no-cpu-baseline > gpurun_out/r2_bench_j.json 2>gpurun_out/r2_bench_j.err
python bench.py --no-side-configs --no-e2e --no-cpu-baseline > gpurun_out/r2_bench_j_long.json 2>gpurun_out/r2_bench_j_long.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:loss_dense -s 3 -c 1 -o gpurun_out/r2_loss_cfg5_w python scripts/loss_bench.py cfg5 ncu > gpurun_out/r2_ncu_w.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:loss_dense -s 3 -c 1 -o gpurun_out/r2_loss_cfg2_w python scripts/loss_bench.py cfg2 ncu > gpurun_out/r2_ncu_w2.log 2>&1
