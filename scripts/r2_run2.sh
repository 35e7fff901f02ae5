set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "loss or graphed or whole_path or full_size_cfg5" > gpurun_out/r2_tests_b.txt 2>&1; tail -15 gpurun_out/r2_tests_b.txt
for w in cfg5 cfg2 cfg3; do
 for cg in 2 4; do RADET_FUSED_CG=$cg timeout 120 python scripts/loss_bench.py $w 100; done
 RADET_DENSE_IMPL=tma timeout 120 python scripts/loss_bench.py $w 100
done > gpurun_out/r2_loss_bench.txt 2>&1
cat gpurun_out/r2_loss_bench.txt
timeout 900 python -m pytest tests/test_gpu_full_size.py -m gpu -x -q > gpurun_out/r2_tests_full.txt 2>&1; tail -15 gpurun_out/r2_tests_full.txt
for cg in 2 4; do
RADET_FUSED_CG=$cg timeout 300 ncu --set full --clock-control none --import-source on -k regex:loss_ -s 3 -c 1 -o gpurun_out/r2_loss_cfg5_cg$cg python scripts/loss_bench.py cfg5 ncu > gpurun_out/r2_ncu_cg$cg.log 2>&1
done
RADET_FUSED_CG=2 timeout 300 ncu --set full --clock-control none --import-source on -k regex:loss_ -s 3 -c 1 -o gpurun_out/r2_loss_cfg2_cg2 python scripts/loss_bench.py cfg2 ncu > gpurun_out/r2_ncu_cfg2.log 2>&1
