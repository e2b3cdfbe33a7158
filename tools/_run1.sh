cd $GRAFT_REPO_ROOT
( timeout 900 python -m pytest tests/test_gpu_export.py tests/test_gpu_driver.py -m gpu -x -q ) > gpurun_out/r2_tests.log 2>&1
tail -30 gpurun_out/r2_tests.log
