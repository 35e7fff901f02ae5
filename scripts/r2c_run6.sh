cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
export RADET_SELECT_HIST=1
python profiles/phase_timing.py 2>&1 | tail -9
