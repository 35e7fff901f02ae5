cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
cp radet_b200/lib/libradet_b200.so /tmp/keep.so
for v in head mask256 head mask256; do
cp radet_b200/lib/variants/$v.so radet_b200/lib/libradet_b200.so
python bench.py --no-side-configs --no-e2e --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/r2c_bench25_$v.json 2>gpurun_out/r2c_bench25_$v.err; tail -2 gpurun_out/r2c_bench25_$v.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2c_bench25_$v.json").read().strip().splitlines()[-1])
print("$v", round(d["value"]), d["ms_per_step"], d["run"]["ms_per_step_one_in_flight"], {k:round(v,2) for k,v in d["stage_us"].items()})
PY
done
for v in head mask256; do
cp radet_b200/lib/variants/$v.so radet_b200/lib/libradet_b200.so
python bench.py --no-e2e --no-cpu-baseline > gpurun_out/r2c_bench25_long_$v.json 2>gpurun_out/r2c_bench25_long_$v.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2c_bench25_long_$v.json").read().strip().splitlines()[-1])
print("long $v", round(d["value"]), d["ms_per_step"], d["other_configs"]["cfg4"]["get_bboxes_vote"]["images_per_s"], d["roofline_large"]["get_bboxes_vote"]["us"])
PY
done
cp /tmp/keep.so radet_b200/lib/libradet_b200.so
