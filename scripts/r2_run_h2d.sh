cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for n in 1 2 4 8; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2955$n scripts/h2d_ceiling.py 2>/dev/null | grep "^{" 
done > gpurun_out/r2_h2d_ceiling.txt
cat gpurun_out/r2_h2d_ceiling.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r2_bench_n4.json 2> gpurun_out/r2_bench_n4.err
