cd $GRAFT_REPO_ROOT
for i in 4 20; do CASYNC_SPLIT=0 CASYNC_PHASE_DBG=$i timeout 60 build/casync_run 64 10 0 2>&1 | grep -A3 "phase dbg"; done
CASYNC_SPLIT=0 timeout 60 build/casync_run 64 20 1 2>&1 | grep -E "batch|down2.1|up2.0|total"
timeout 60 build/casync_run 64 30 0 2>&1 | tail -2 
