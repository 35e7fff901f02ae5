cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py -m gpu -x -q -k "bboxes or nms or ops or whole_path or graphed or candidates" > gpurun_out/r2_tests_k.txt 2>&1; tail -4 gpurun_out/r2_tests_k.txt
python bench.py --steps 20 --warmup 5 --no-side-configs --no-e2e --no-cpu-baseline > gpurun_out/r2_bench_e.json 2>gpurun_out/r2_bench_e.err
python bench.py --no-side-configs --no-e2e --no-cpu-baseline > gpurun_out/r2_bench_e_long.json 2>gpurun_out/r2_bench_e_long.err
python profiles/phase_timing.py 2>&1 | head -12 > gpurun_out/r2_phase_timing3.txt; cat gpurun_out/r2_phase_timing3.txt
