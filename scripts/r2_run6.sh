cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "loss or graphed or whole_path or full_size_cfg5" > gpurun_out/r2_tests_e.txt 2>&1; tail -5 gpurun_out/r2_tests_e.txt
python scripts/loss_timeline.py cfg2 > gpurun_out/r2_timeline2.txt 2>&1
python scripts/loss_timeline.py cfg5 >> gpurun_out/r2_timeline2.txt 2>&1
cat gpurun_out/r2_timeline2.txt
for w in cfg5 cfg2 cfg3; do
 for cg in 4 2; do RADET_FUSED_CG=$cg timeout 120 python scripts/loss_bench.py $w 100; done
 for t in 1.5 5; do RADET_FUSED_IPW=$t timeout 120 python scripts/loss_bench.py $w 100; done
done 2>&1 | grep "^{" > gpurun_out/r2_loss_bench4.txt
cat gpurun_out/r2_loss_bench4.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:loss_ -s 3 -c 1 -o gpurun_out/r2_loss_cfg5_v4 python scripts/loss_bench.py cfg5 ncu > gpurun_out/r2_ncu_v4.log 2>&1
