cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for B in 1 2 4 8 16 32 64 256; do
CASYNC_GRAPH=0 timeout 60 build/casync_run $B 60 0 2>&1 | tail -2 | head -1
timeout 60 build/casync_run $B 60 0 2>&1 | tail -2 | head -1
done
