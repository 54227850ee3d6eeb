cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for B in 48 64 96 128 256 1024; do
CASYNC_SPLIT=0 timeout 60 build/casync_run $B 20 0 2>&1 | tail -2 | head -1
timeout 60 build/casync_run $B 20 0 2>&1 | tail -2 | head -1
done
CASYNC_SPLIT=0 timeout 60 build/casync_run 64 4 0 2>&1 | tail -1
timeout 60 build/casync_run 64 4 0 2>&1 | tail -1
