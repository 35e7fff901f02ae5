cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 300 python scripts/sanitize_driver.py > gpurun_out/r2_sanitize_plain.txt 2>&1; tail -3 gpurun_out/r2_sanitize_plain.txt
for tool in memcheck racecheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_driver.py > gpurun_out/r2_sanitizer_$tool.log 2>&1
  echo "== $tool"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|sanitize driver done|smoke OK" gpurun_out/r2_sanitizer_$tool.log | sort | uniq -c | head -12
done
