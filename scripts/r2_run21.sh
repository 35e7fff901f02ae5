cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python scripts/dense_timeline.py cfg5 > gpurun_out/r2_dense_timeline.txt 2>&1
python scripts/dense_timeline.py cfg2 >> gpurun_out/r2_dense_timeline.txt 2>&1
cat gpurun_out/r2_dense_timeline.txt
for w in cfg5 cfg2; do timeout 120 python scripts/loss_bench.py $w 100; done 2>&1 | grep "^{"
