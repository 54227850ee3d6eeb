cd $GRAFT_REPO_ROOT
for B in 24 64 256; do
timeout 60 build/casync_run $B 20 0 2>&1 | tail -2
CASYNC_CHAIN=1 timeout 60 build/casync_run $B 20 0 2>&1 | tail -2
done
for B in 1 8; do
timeout 60 build/casync_run $B 50 0 2>&1 | tail -2 | head -1
CASYNC_CHAIN=1 timeout 60 build/casync_run $B 50 0 2>&1 | tail -2 | head -1
done
