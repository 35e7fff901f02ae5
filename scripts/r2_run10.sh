cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 --no-side-configs --no-e2e --no-cpu-baseline > gpurun_out/r2_bench_fused.json 2> gpurun_out/r2_bench_fused.err
RADET_DENSE_IMPL=tma python bench.py --steps 20 --warmup 5 --no-side-configs --no-e2e --no-cpu-baseline > gpurun_out/r2_bench_tma.json 2> gpurun_out/r2_bench_tma.err
python bench.py --no-side-configs --no-e2e --no-cpu-baseline > gpurun_out/r2_bench_fused_long.json 2> gpurun_out/r2_bench_fused_long.err
RADET_DENSE_IMPL=tma python bench.py --no-side-configs --no-e2e --no-cpu-baseline > gpurun_out/r2_bench_tma_long.json 2> gpurun_out/r2_bench_tma_long.err
tail -3 gpurun_out/*.err
