cd $GRAFT_REPO_ROOT
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2_tests.log 2>&1
tail -4 gpurun_out/r2_tests.log
run() { echo "== $1 flags=$2 env=$3"; env $3 timeout 300 python tools/k1_ablate.py --steps 60 --ablate $4 --flags $2 $5 > gpurun_out/r2_ab_$1.log 2>&1; grep '^{"flags' gpurun_out/r2_ab_$1.log | cut -c1-250; }
run base 0 A=1 "0"
run slab40 0 A=1 "0" "--dims 1024 1024 1024 --slab 648 40"
run slab40z3 0 SFM_ZL_LOG2=3 "0" "--dims 1024 1024 1024 --slab 648 40"
run slab16 0 A=1 "0" "--dims 1024 1024 1024 --slab 656 16"
