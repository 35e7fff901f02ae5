cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py -m gpu -x -q -k "assign or whole_path or graphed or full_size or targets or mt19937 or uniform" > gpurun_out/r2c_tests22.txt 2>&1; tail -3 gpurun_out/r2c_tests22.txt
timeout 120 python profiles/phase_timing.py 2>&1 | tail -9
RADET_ASSIGN_DUO=0 timeout 120 python profiles/phase_timing.py 2>&1 | tail -3
timeout 300 python bench.py --no-side-configs --no-e2e --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/r2c_bench22.json 2>gpurun_out/r2c_bench22.err; tail -2 gpurun_out/r2c_bench22.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2c_bench22.json").read().strip().splitlines()[-1])
print(round(d["value"]), d["ms_per_step"], d["run"]["ms_per_step_one_in_flight"], {k:round(v,2) for k,v in d["stage_us"].items()})
PY
