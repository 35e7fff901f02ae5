cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
LOSS_HINT=1 python scripts/dense_timeline.py cfg5 2>&1 | grep -v "Warning\|q = lambda\|_nanquantile\|per item\|setup->\|stage0->\|stage7->\| box planes n\| exit n" | tee gpurun_out/r2_dense_timeline8.txt
