cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python profiles/phase_timing.py 2>&1 | grep -A12 "detect_bin per" 
echo "--- RADET_SELECT_HIST=1"
RADET_SELECT_HIST=1 python profiles/phase_timing.py 2>&1 | grep -A12 "detect_bin per" 
for m in 0 1; do
if [ $m = 1 ]; then export RADET_SELECT_HIST=1; fi
python bench.py --no-side-configs --no-e2e --no-cpu-baseline > gpurun_out/r2c_bench4_$m.json 2>gpurun_out/r2c_bench4_$m.err; tail -2 gpurun_out/r2c_bench4_$m.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2c_bench4_$m.json").read().strip().splitlines()[-1])
print(round(d["value"]), d["ms_per_step"], d["run"]["ms_per_step_one_in_flight"], {k:round(v,2) for k,v in d["stage_us"].items()}, d["roofline"]["frac"])
PY
done
