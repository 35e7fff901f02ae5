cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_final_ref.json 2> gpurun_out/r2_final_ref.err
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_final_n1.json 2> gpurun_out/r2_final_n1.err
python bench.py > gpurun_out/r2_final_n1_long.json 2> gpurun_out/r2_final_n1_long.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv python bench.py --profile --steps 12 --warmup 3 > gpurun_out/r2_launches.log 2>&1
ncu --set full --clock-control none --import-source on --launch-skip 66 -c 11 -o gpurun_out/r2_step python bench.py --profile --steps 12 --warmup 3 > gpurun_out/r2_step.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:loss_ -s 4 -c 2 -o gpurun_out/r2_loss_large python scripts/loss_bench.py cfg5 ncu > gpurun_out/r2_loss_large.log 2>&1
python profiles/phase_timing.py > gpurun_out/r2_phase_timing_final.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests_final.txt 2>&1; tail -3 gpurun_out/r2_tests_final.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke_final.txt 2>&1; cat gpurun_out/r2_smoke_final.txt
