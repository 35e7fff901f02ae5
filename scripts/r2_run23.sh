cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for w in cfg5; do
  python scripts/dense_timeline.py $w
  LOSS_HINT=1 python scripts/dense_timeline.py $w
  RADET_LOSS_PDL=0 python scripts/dense_timeline.py $w
done > gpurun_out/r2_dense_timeline3.txt 2>&1
cat gpurun_out/r2_dense_timeline3.txt
