cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 60 build/casync_run 64 20 1 2>&1 | grep -E "batch|\.dw|total"
CASYNC_CHAIN=1 timeout 60 build/casync_run 64 20 0 2>&1 | grep -E "batch"
