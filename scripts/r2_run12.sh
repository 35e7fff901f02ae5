cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
tail -5 gpurun_out/r2_bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --impl reference > gpurun_out/r2_bench_n2_ref.json 2> gpurun_out/r2_bench_n2_ref.err
nvidia-smi topo -m > gpurun_out/r2_topo2.txt 2>&1; cat /sys/devices/system/node/online >> gpurun_out/r2_topo2.txt
