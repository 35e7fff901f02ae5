cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
N=${NGPU:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2c_bench_n$N.json 2> gpurun_out/r2c_bench_n$N.err
tail -3 gpurun_out/r2c_bench_n$N.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2c_bench_n$N.json").read().strip().splitlines()[-1])
print(d["n_gpus"], round(d["value"]), d["ms_per_step"], d["e2e"]["value"], d["e2e"]["maps_resident"]["value"], d["run"]["host_enqueue_us_per_step_by_rank"], d.get("sync_num_pos"))
PY
if [ "$N" = "2" ]; then timeout 600 python -m pytest tests/test_gpu_nccl.py -m gpu -x -q 2>&1 | tail -2; fi
