cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py -m gpu -x -q -k "bboxes or detect or candidates or nms or whole_path or graphed or full_size or bbox2result or head" > gpurun_out/r2c_tests24.txt 2>&1; tail -3 gpurun_out/r2c_tests24.txt
for f in 1 0 1 0; do
RADET_RANK_FUSED=$f timeout 300 python bench.py --no-side-configs --no-e2e --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/r2c_bench24_$f.json 2>gpurun_out/r2c_bench24_$f.err; tail -2 gpurun_out/r2c_bench24_$f.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2c_bench24_$f.json").read().strip().splitlines()[-1])
print("fused=$f", round(d["value"]), d["ms_per_step"], d["run"]["ms_per_step_one_in_flight"], {k:round(v,2) for k,v in d["stage_us"].items()}, d["launches_per_step"])
PY
done
for f in 1 0; do
RADET_RANK_FUSED=$f timeout 300 python bench.py --no-e2e --no-cpu-baseline > gpurun_out/r2c_bench24_long_$f.json 2>gpurun_out/r2c_bench24_long_$f.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2c_bench24_long_$f.json").read().strip().splitlines()[-1])
print("long fused=$f", round(d["value"]), d["ms_per_step"], d["run"]["ms_per_step_one_in_flight"], d["other_configs"]["cfg4"]["get_bboxes_vote"]["images_per_s"], d["other_configs"]["cfg4"]["get_bboxes_nms"]["images_per_s"], d["roofline_large"]["get_bboxes_vote"]["us"])
PY
done
