cd $GRAFT_REPO_ROOT
run() { echo "== $1 : $2 $3"; env $2 timeout 300 python tools/k1_ablate.py --steps 100 --ablate 0 $3 > gpurun_out/r2_ab_$1.log 2>&1; grep '^{"flags' gpurun_out/r2_ab_$1.log | cut -c1-300; }
run overlap A=1 ""
run overlap6 SFM_K1B_BLOCKS_PER_SM=6 ""
( timeout 600 python -m pytest tests/test_gpu_integrate.py tests/test_gpu_raymarch.py -m gpu -x -q ) > gpurun_out/r2_tests.log 2>&1
tail -3 gpurun_out/r2_tests.log
