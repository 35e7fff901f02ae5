cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
cp radet_b200/lib/libradet_b200.so /tmp/keep.so
for v in base head res64 res64_bin64 base head; do
cp radet_b200/lib/variants/$v.so radet_b200/lib/libradet_b200.so
python bench.py --no-side-configs --no-e2e --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/r2c_bench9_$v.json 2>gpurun_out/r2c_bench9_$v.err; tail -2 gpurun_out/r2c_bench9_$v.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2c_bench9_$v.json").read().strip().splitlines()[-1])
print("$v", round(d["value"]), d["ms_per_step"], d["run"]["ms_per_step_one_in_flight"], {k:round(v,2) for k,v in d["stage_us"].items()})
PY
done
cp /tmp/keep.so radet_b200/lib/libradet_b200.so
