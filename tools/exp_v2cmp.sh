cd /root/repo 2>/dev/null || cd $GRAFT_REPO_ROOT
JNE_KERNEL=v2 ./tools/exp_lib_compare.sh base shfl 2>&1 | grep variant | sed 's/variant/v2 variant/'
JNE_KERNEL=v1 ./tools/exp_lib_compare.sh base 2>&1 | grep variant | sed 's/variant/v1 variant/'
