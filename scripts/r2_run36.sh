cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 300 python scripts/sanitize_driver.py > gpurun_out/r2b_sanitize_plain.txt 2>&1; tail -2 gpurun_out/r2b_sanitize_plain.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2b_launches.csv python bench.py --profile --steps 12 --warmup 3 > gpurun_out/r2b_launches.log 2>&1
ncu --set full --clock-control none --import-source on --launch-skip 66 -c 11 -o gpurun_out/r2b_step python bench.py --profile --steps 12 --warmup 3 > gpurun_out/r2b_step.log 2>&1
LOSS_HINT=1 ncu --set full --clock-control none --import-source on -k regex:loss_ -s 4 -c 2 -o gpurun_out/r2b_loss_large python scripts/loss_bench.py cfg5 ncu > gpurun_out/r2b_loss_large.log 2>&1
ls -la gpurun_out/r2b_*
