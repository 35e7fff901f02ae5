cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for ch in 1 2 3 1 2; do
RADET_DENSE_CHUNKS=$ch python bench.py --no-side-configs --no-e2e --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/r2c_bench12_$ch.json 2>gpurun_out/r2c_bench12_$ch.err; tail -2 gpurun_out/r2c_bench12_$ch.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2c_bench12_$ch.json").read().strip().splitlines()[-1])
print("chunks=$ch", round(d["value"]), d["ms_per_step"], d["run"]["ms_per_step_one_in_flight"], {k:round(v,2) for k,v in d["stage_us"].items()}, round(d["roofline"]["frac"],4))
PY
done
for ch in 1 2; do
RADET_DENSE_CHUNKS=$ch python bench.py --no-side-configs --no-e2e --no-cpu-baseline > gpurun_out/r2c_bench12_long_$ch.json 2>gpurun_out/r2c_bench12_long_$ch.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2c_bench12_long_$ch.json").read().strip().splitlines()[-1])
print("long chunks=$ch", round(d["value"]), d["ms_per_step"], round(d["roofline"]["frac"],4))
PY
done
