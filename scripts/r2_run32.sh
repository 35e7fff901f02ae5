cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tower.py -m gpu -x -q > gpurun_out/r2_tests_tower.txt 2>&1; tail -5 gpurun_out/r2_tests_tower.txt
python scripts/tower_bench.py 2>&1 | tail -4 | tee gpurun_out/r2_tower_bench.txt
