cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "loss or graphed or whole_path or full_size_cfg5" > gpurun_out/r2_tests_g.txt 2>&1; tail -5 gpurun_out/r2_tests_g.txt
timeout 900 python -m pytest tests/test_gpu_full_size.py -m gpu -x -q > gpurun_out/r2_tests_full3.txt 2>&1; tail -15 gpurun_out/r2_tests_full3.txt
python scripts/loss_timeline.py cfg5 > gpurun_out/r2_timeline5.txt 2>&1
LOSS_HINT=1 python scripts/loss_timeline.py cfg5 >> gpurun_out/r2_timeline5.txt 2>&1
LOSS_HINT=1 python scripts/loss_timeline.py cfg2 >> gpurun_out/r2_timeline5.txt 2>&1
cat gpurun_out/r2_timeline5.txt
for w in cfg5 cfg2 cfg3; do
 timeout 120 python scripts/loss_bench.py $w 100
 LOSS_HINT=1 timeout 120 python scripts/loss_bench.py $w 100
 LOSS_HINT=1 RADET_FUSED_IPW=2 timeout 120 python scripts/loss_bench.py $w 100
 LOSS_HINT=1 RADET_FUSED_IPW=4.5 timeout 120 python scripts/loss_bench.py $w 100
done 2>&1 | grep "^{" > gpurun_out/r2_loss_bench6.txt
LOSS_HINT=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:loss_ -s 3 -c 1 -o gpurun_out/r2_loss_cfg5_v6 python scripts/loss_bench.py cfg5 ncu > gpurun_out/r2_ncu_v6.log 2>&1
