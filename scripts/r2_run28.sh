cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for ch in 0 2 3 4; do
RADET_DENSE_CHUNKS=$ch python bench.py --no-side-configs --no-e2e --no-cpu-baseline > gpurun_out/r2_bench_l_ch$ch.json 2>gpurun_out/r2_bench_l.err
done
python - <<'PY'
import json
for ch in (0,2,3,4):
    d=json.loads(open(f"gpurun_out/r2_bench_l_ch{ch}.json").read().strip().splitlines()[-1])
    print(ch, round(d["value"]), d["ms_per_step"], d["run"]["ms_per_step_one_in_flight"], {k:round(v,2) for k,v in d["stage_us"].items()}, round(d["roofline"]["frac"],4))
PY
