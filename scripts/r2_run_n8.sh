cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
N=${NGPU:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
tail -3 gpurun_out/r2_bench_n$N.err
nvidia-smi topo -m > gpurun_out/r2_topo$N.txt 2>&1; cat /sys/devices/system/node/online >> gpurun_out/r2_topo$N.txt; nproc >> gpurun_out/r2_topo$N.txt
