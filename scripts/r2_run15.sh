cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --report-api-errors no --print-limit 20 python scripts/sanitize_driver.py > gpurun_out/r2_sanitizer_memcheck.log 2>&1
grep -E "ERROR SUMMARY|sanitize driver done" gpurun_out/r2_sanitizer_memcheck.log
timeout 1500 compute-sanitizer --tool initcheck --print-limit 20 python scripts/sanitize_driver.py > gpurun_out/r2_sanitizer_initcheck.log 2>&1
grep -E "ERROR SUMMARY|sanitize driver done" gpurun_out/r2_sanitizer_initcheck.log
for w in cfg5 cfg2; do
 timeout 120 python scripts/loss_bench.py $w 100
 RADET_LOSS_IMPL=hybrid timeout 120 python scripts/loss_bench.py $w 100
done 2>&1 | grep "^{" > gpurun_out/r2_loss_bench8.txt
for u in 12 16 20 24; do
 python bench.py --steps 20 --warmup 5 --inflight $u --no-side-configs --no-e2e --no-cpu-baseline > gpurun_out/r2_bench_U$u.json 2>gpurun_out/r2_bench_U$u.err
done
for u in 12 16; do
 python bench.py --inflight $u --no-side-configs --no-e2e --no-cpu-baseline > gpurun_out/r2_bench_U${u}_long.json 2>gpurun_out/r2_bench_U${u}_long.err
done
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests_all.txt 2>&1; tail -5 gpurun_out/r2_tests_all.txt
