cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py -m gpu -x -q -k "bboxes or detect or candidates or nms or whole_path or graphed or full_size" > gpurun_out/r2c_tests10.txt 2>&1; tail -2 gpurun_out/r2c_tests10.txt
cp radet_b200/lib/libradet_b200.so /tmp/keep.so
for v in base head base head; do
cp radet_b200/lib/variants/$v.so radet_b200/lib/libradet_b200.so
python bench.py --no-side-configs --no-e2e --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/r2c_bench10_$v.json 2>gpurun_out/r2c_bench10_$v.err; tail -2 gpurun_out/r2c_bench10_$v.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2c_bench10_$v.json").read().strip().splitlines()[-1])
print("$v", round(d["value"]), d["ms_per_step"], d["run"]["ms_per_step_one_in_flight"], {k:round(v,2) for k,v in d["stage_us"].items()})
PY
done
cp /tmp/keep.so radet_b200/lib/libradet_b200.so
