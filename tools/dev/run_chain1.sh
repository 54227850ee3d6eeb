cd $GRAFT_REPO_ROOT
for l in down4.1.pw2 fuse0.0.pw1 "attention_blocks.0|b1_w"; do
CASYNC_GRAPH=0 CASYNC_SPLIT=0 CASYNC_GEMM_DBG="$l" timeout 60 build/casync_run 64 10 0 2>&1 | grep -A2 "gemm dbg"
done
