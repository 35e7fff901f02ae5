cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for np_ in 3 4 6; do
RADET_E2E_INFLIGHT=$np_ python bench.py --steps 20 --warmup 5 --no-side-configs --no-cpu-baseline > gpurun_out/r2_bench_p_$np_.json 2>gpurun_out/r2_bench_p.err
done
python - <<'PY'
import json
for n in (3,4,6):
    d=json.loads(open(f"gpurun_out/r2_bench_p_{n}.json").read().strip().splitlines()[-1])
    e=d["e2e"]; print(n, round(e["value"]), e["ms_per_step"], e["h2d_GBps_per_gpu"], round(e["maps_resident"]["value"]), round(e["eager"]["value"]))
PY
