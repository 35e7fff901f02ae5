cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for u in 8 12 16; do
python bench.py --inflight $u --no-side-configs --no-e2e --no-cpu-baseline > gpurun_out/r2_bench_q_$u.json 2>gpurun_out/r2_bench_q.err
python bench.py --inflight $u --steps 20 --warmup 5 --no-side-configs --no-e2e --no-cpu-baseline > gpurun_out/r2_bench_q20_$u.json 2>gpurun_out/r2_bench_q.err
done
python - <<'PY'
import json
for n in (8,12,16):
    d=json.loads(open(f"gpurun_out/r2_bench_q_{n}.json").read().strip().splitlines()[-1])
    e=json.loads(open(f"gpurun_out/r2_bench_q20_{n}.json").read().strip().splitlines()[-1])
    print(n, "long", round(d["value"]), d["ms_per_step"], "20-step", round(e["value"]), e["ms_per_step"])
PY
