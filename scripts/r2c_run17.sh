cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:class_nms --launch-skip 2 -c 1 -f -o gpurun_out/r2c_nms python profiles/phase_timing.py > gpurun_out/r2c_nms_ncu.log 2>&1; tail -3 gpurun_out/r2c_nms_ncu.log
ls -la gpurun_out/r2c_nms.ncu-rep
