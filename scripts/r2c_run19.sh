cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py -m gpu -x -q -k "bboxes or detect or candidates or nms or whole_path or graphed or full_size" > gpurun_out/r2c_tests19.txt 2>&1; tail -2 gpurun_out/r2c_tests13.txt
python profiles/phase_timing.py 2>&1 | head -22
for i in 1 2; do
python bench.py --no-side-configs --no-e2e --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/r2c_bench19_$i.json 2>gpurun_out/r2c_bench19_$i.err; tail -2 gpurun_out/r2c_bench19_$i.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2c_bench19_$i.json").read().strip().splitlines()[-1])
print(round(d["value"]), d["ms_per_step"], d["run"]["ms_per_step_one_in_flight"], {k:round(v,2) for k,v in d["stage_us"].items()})
PY
done
python bench.py --no-side-configs --no-e2e --no-cpu-baseline > gpurun_out/r2c_bench19_long.json 2>gpurun_out/r2c_bench19_long.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2c_bench19_long.json").read().strip().splitlines()[-1])
print("long", round(d["value"]), d["ms_per_step"], d["run"]["ms_per_step_one_in_flight"])
PY
