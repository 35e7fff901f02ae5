cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_tests_final.txt 2>&1; tail -3 gpurun_out/r2c_tests_final.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c_smoke_final.txt 2>&1; tail -2 gpurun_out/r2c_smoke_final.txt
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2c_final_ref.json 2> gpurun_out/r2c_final_ref.err
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2c_final_n1.json 2> gpurun_out/r2c_final_n1.err
python bench.py > gpurun_out/r2c_final_n1_long.json 2> gpurun_out/r2c_final_n1_long.err
python profiles/phase_timing.py > gpurun_out/r2c_phase_timing.txt 2>&1
python - <<'PY'
import json
for f in ("r2c_final_ref","r2c_final_n1","r2c_final_n1_long"):
    d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    print(f, round(d["value"],1), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), (d.get("roofline") or {}).get("frac"), ((d.get("roofline_large") or {}).get("loss_fwd_bwd") or {}).get("frac"), (d.get("run") or {}).get("ms_per_step_one_in_flight"), d.get("stage_us"))
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2c_launches.csv python bench.py --profile --steps 12 --warmup 3 > gpurun_out/r2c_launches.log 2>&1
python profiles/launch_summary.py gpurun_out/r2c_launches.csv > gpurun_out/r2c_launch_summary.txt; cat gpurun_out/r2c_launch_summary.txt
ncu --set full --clock-control none --import-source on --launch-skip 66 -c 11 -o gpurun_out/r2c_step python bench.py --profile --steps 12 --warmup 3 > gpurun_out/r2c_ncu_step.log 2>&1
python profiles/ncu_table.py gpurun_out/r2c_step.ncu-rep > gpurun_out/r2c_ncu_step.txt 2>&1; cat gpurun_out/r2c_ncu_step.txt | cut -c1-200
for tool in memcheck racecheck synccheck; do
  extra=""; [ "$tool" = "memcheck" ] && extra="--report-api-errors no"
  timeout 900 compute-sanitizer --tool $tool $extra --print-limit 20 python scripts/sanitize_driver.py > gpurun_out/r2c_sanitizer_$tool.log 2>&1
  echo "== $tool"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|sanitize driver done|smoke OK" gpurun_out/r2c_sanitizer_$tool.log | sort | uniq -c | head -12
done
