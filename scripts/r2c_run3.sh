cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py -m gpu -x -q -k "assign or bboxes or detect or candidates or nms or whole_path or graphed or full_size" > gpurun_out/r2c_tests3.txt 2>&1; tail -3 gpurun_out/r2c_tests2.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2c_launches3.csv python profiles/phase_timing.py > gpurun_out/r2c_phase_timing3.txt 2>&1; tail -9 gpurun_out/r2c_phase_timing2.txt
python profiles/launch_summary.py gpurun_out/r2c_launches3.csv 2>&1 | tail -20
python bench.py --no-side-configs --no-e2e --no-cpu-baseline > gpurun_out/r2c_bench3.json 2>gpurun_out/r2c_bench3.err; tail -2 gpurun_out/r2c_bench3.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2c_bench3.json").read().strip().splitlines()[-1])
print(round(d["value"]), d["ms_per_step"], d["run"]["ms_per_step_one_in_flight"], {k:round(v,2) for k,v in d["stage_us"].items()}, d["roofline"]["frac"])
PY
