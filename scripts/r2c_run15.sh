cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for u in 8 10 5 20 10 8; do
python bench.py --no-side-configs --no-e2e --no-cpu-baseline --steps 20 --warmup 5 --inflight $u > gpurun_out/r2c_bench15_$u.json 2>gpurun_out/r2c_bench15_$u.err; tail -2 gpurun_out/r2c_bench15_$u.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2c_bench15_$u.json").read().strip().splitlines()[-1])
print("inflight=$u", round(d["value"]), d["ms_per_step"], d["run"]["ms_per_step_one_in_flight"], [round(x,4) for x in d["run"]["block_ms"]])
PY
done
