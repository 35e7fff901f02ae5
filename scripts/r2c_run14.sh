cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "many_candidates or low_threshold" > gpurun_out/r2c_tests14.txt 2>&1; tail -15 gpurun_out/r2c_tests14.txt
