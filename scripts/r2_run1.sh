set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( nproc; cat /sys/devices/system/node/online; ls /sys/devices/system/node/; nvidia-smi topo -m; python -c "import os;print(len(os.sched_getaffinity(0)))"; lscpu | head -30; free -g ) > gpurun_out/r2_topo.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.txt 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err
for u in 4 5 10 16; do
python bench.py --steps 20 --warmup 5 --inflight $u --no-e2e --no-cpu-baseline --no-side-configs > gpurun_out/r2_bench_u$u.json 2> gpurun_out/r2_bench_u$u.err
done
python bench.py --no-e2e --no-cpu-baseline --no-side-configs > gpurun_out/r2_bench_long.json 2> gpurun_out/r2_bench_long.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests_a.txt 2>&1
tail -3 gpurun_out/r2_tests_a.txt
