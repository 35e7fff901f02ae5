cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests_l.txt 2>&1; tail -4 gpurun_out/r2_tests_l.txt
python bench.py --steps 20 --warmup 5 --no-side-configs --no-e2e --no-cpu-baseline > gpurun_out/r2_bench_f.json 2>gpurun_out/r2_bench_f.err
python bench.py --no-side-configs --no-e2e --no-cpu-baseline > gpurun_out/r2_bench_f_long.json 2>gpurun_out/r2_bench_f_long.err
RADET_NO_PDL=1 python bench.py --steps 20 --warmup 5 --no-side-configs --no-e2e --no-cpu-baseline > gpurun_out/r2_bench_g.json 2>gpurun_out/r2_bench_g.err
RADET_NO_PDL=1 python bench.py --no-side-configs --no-e2e --no-cpu-baseline > gpurun_out/r2_bench_g_long.json 2>gpurun_out/r2_bench_g_long.err
tail -2 gpurun_out/r2_bench_f.err
