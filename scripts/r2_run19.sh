cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests_m.txt 2>&1; tail -4 gpurun_out/r2_tests_m.txt
python bench.py --steps 20 --warmup 5 --no-side-configs --no-e2e --no-cpu-baseline > gpurun_out/r2_bench_h.json 2>gpurun_out/r2_bench_h.err
python bench.py --no-side-configs --no-e2e --no-cpu-baseline > gpurun_out/r2_bench_h_long.json 2>gpurun_out/r2_bench_h_long.err
RADET_SEPARATE_RANK=1 python bench.py --steps 20 --warmup 5 --no-side-configs --no-e2e --no-cpu-baseline > gpurun_out/r2_bench_i.json 2>gpurun_out/r2_bench_i.err
tail -2 gpurun_out/r2_bench_h.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_launches_b.csv python bench.py --profile --steps 12 --warmup 3 > /dev/null 2>&1
python profiles/launch_summary.py gpurun_out/r2_launches_b.csv
