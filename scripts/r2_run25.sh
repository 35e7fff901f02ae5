cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for w in cfg5 cfg2 cfg3; do
  timeout 120 python scripts/loss_bench.py $w 100
  LOSS_HINT=1 timeout 120 python scripts/loss_bench.py $w 100
done 2>&1 | grep "^{" > gpurun_out/r2_loss_bench15.txt
cat gpurun_out/r2_loss_bench15.txt
LOSS_HINT=1 python scripts/dense_timeline.py cfg5 > gpurun_out/r2_dense_timeline7.txt 2>&1
python scripts/dense_timeline.py cfg5 >> gpurun_out/r2_dense_timeline7.txt 2>&1
grep -v "per item\|setup->\|stage0->\|stage7->\| box planes n\| exit n\|Warning\|q = lambda\|_nanquantile" gpurun_out/r2_dense_timeline7.txt
