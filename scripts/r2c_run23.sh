cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  extra=""; [ "$tool" = "memcheck" ] && extra="--report-api-errors no"
  timeout 900 compute-sanitizer --tool $tool $extra --print-limit 20 python scripts/sanitize_driver.py > gpurun_out/r2c_sanitizer_$tool.log 2>&1
  echo "== $tool"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|sanitize driver done|smoke OK|Traceback|Error" gpurun_out/r2c_sanitizer_$tool.log | sort | uniq -c | head -12
done
