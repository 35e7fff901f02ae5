cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python scripts/loss_timeline.py cfg2 > gpurun_out/r2_timeline3.txt 2>&1
python scripts/loss_timeline.py cfg5 >> gpurun_out/r2_timeline3.txt 2>&1
cat gpurun_out/r2_timeline3.txt
