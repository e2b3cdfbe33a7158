cd $GRAFT_REPO_ROOT
run() { echo "== $1 flags=$2 env=$3"; env $3 timeout 300 python tools/k1_ablate.py --steps 60 --ablate 0 --flags $2 > gpurun_out/r2_ab_$1.log 2>&1; grep '^{"flags' gpurun_out/r2_ab_$1.log | cut -c1-250; }
for v in f8 f8b3; do run $v 0 SFM_B200_LIB=$GRAFT_REPO_ROOT/build/libsfm_b200_$v.so; done
timeout 300 python tools/k1_ablate.py --steps 20 --ablate 0 16 2 > gpurun_out/r2_plain.log 2>&1 && grep '^{"flags' gpurun_out/r2_plain.log | cut -c1-330 && \
ncu --set full --clock-control none --import-source on -k regex:'integrate_kernel|classify_kernel' -s 20 -c 2 -o gpurun_out/r2_k1_split -f python tools/k1_ablate.py --steps 20 --ablate 0 > gpurun_out/r2_ncu.log 2>&1
tail -2 gpurun_out/r2_ncu.log
