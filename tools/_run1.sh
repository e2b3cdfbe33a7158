cd $GRAFT_REPO_ROOT
free -g | head -2; nproc
( time timeout 600 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q ) > gpurun_out/r2_tests.log 2>&1
tail -15 gpurun_out/r2_tests.log
