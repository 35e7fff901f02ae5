set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "loss or graphed or whole_path or full_size_cfg5" > gpurun_out/r2_tests_d.txt 2>&1; tail -5 gpurun_out/r2_tests_d.txt
for w in cfg5 cfg2 cfg3; do
 for cg in 2 4; do RADET_FUSED_CG=$cg timeout 120 python scripts/loss_bench.py $w 100; done
 for t in 600 1200 2400; do RADET_FUSED_WARPS=$t timeout 120 python scripts/loss_bench.py $w 100; done
done > gpurun_out/r2_loss_bench3.txt 2>&1
grep "^{" gpurun_out/r2_loss_bench3.txt
RADET_FUSED_CG=2 timeout 300 ncu --set full --clock-control none --import-source on -k regex:loss_ -s 3 -c 1 -o gpurun_out/r2_loss_cfg5_v3 python scripts/loss_bench.py cfg5 ncu > gpurun_out/r2_ncu_v3.log 2>&1
RADET_FUSED_CG=2 timeout 300 ncu --set full --clock-control none --import-source on -k regex:loss_ -s 3 -c 1 -o gpurun_out/r2_loss_cfg2_v3 python scripts/loss_bench.py cfg2 ncu > gpurun_out/r2_ncu_cfg2_v3.log 2>&1
