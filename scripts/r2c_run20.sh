cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:detect_bin --launch-skip 2 -c 1 -f -o gpurun_out/r2c_bin python profiles/phase_timing.py > gpurun_out/r2c_bin_ncu.log 2>&1; tail -2 gpurun_out/r2c_bin_ncu.log
ncu --set full --clock-control none --import-source on -k regex:assign_resolve --launch-skip 2 -c 1 -f -o gpurun_out/r2c_resolve python profiles/phase_timing.py > gpurun_out/r2c_resolve_ncu.log 2>&1; tail -2 gpurun_out/r2c_resolve_ncu.log
ls -la gpurun_out/r2c_bin.ncu-rep gpurun_out/r2c_resolve.ncu-rep
