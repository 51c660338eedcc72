cd $GRAFT_REPO_ROOT
tools/sweep_pool.sh "" "-DRTC_RESTORE_WORLD=0" 2>&1
echo "TLAS leaf 3:"
RTC_TLAS_LEAF=3 tools/sweep_pool.sh "-DRTC_RESTORE_WORLD=0" 2>&1
