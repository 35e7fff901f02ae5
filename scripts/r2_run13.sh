cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py -m gpu -x -q -k "assign or mt19937 or label_assignment or full_batch_assignment or whole_path or graphed" > gpurun_out/r2_tests_i.txt 2>&1; tail -15 gpurun_out/r2_tests_i.txt
python profiles/phase_timing.py 2>&1 | tail -12 > gpurun_out/r2_phase_timing2.txt; cat gpurun_out/r2_phase_timing2.txt
