cd $GRAFT_REPO_ROOT
( timeout 900 python -m pytest tests/test_gpu_sharded_merge.py tests/test_gpu_raymarch.py tests/test_gpu_sharded_raycast.py -m gpu -x -q ) > gpurun_out/r2_tests.log 2>&1
tail -30 gpurun_out/r2_tests.log
