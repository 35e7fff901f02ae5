cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2b_final_ref.json 2> gpurun_out/r2b_final_ref.err
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2b_final_n1.json 2> gpurun_out/r2b_final_n1.err
python bench.py > gpurun_out/r2b_final_n1_long.json 2> gpurun_out/r2b_final_n1_long.err
python profiles/phase_timing.py > gpurun_out/r2b_phase_timing.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_tests_final.txt 2>&1; tail -3 gpurun_out/r2b_tests_final.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2b_smoke_final.txt 2>&1; cat gpurun_out/r2b_smoke_final.txt
python - <<'PY'
import json
for f in ("r2b_final_ref","r2b_final_n1","r2b_final_n1_long"):
    d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    print(f, round(d["value"],1), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), (d.get("roofline") or {}).get("frac"), ((d.get("roofline_large") or {}).get("loss_fwd_bwd") or {}).get("frac"))
PY
