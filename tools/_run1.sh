cd $GRAFT_REPO_ROOT
python bench.py --steps 200 --warmup 5 > gpurun_out/r1b_bench_n1.json 2> gpurun_out/r1b_bench_n1.err
tail -c 1500 gpurun_out/r1b_bench_n1.json; tail -3 gpurun_out/r1b_bench_n1.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r1b_bench_ref.json 2>> gpurun_out/r1b_bench_n1.err
cut -c1-600 gpurun_out/r1b_bench_ref.json
python bench.py --steps 8 --warmup 3 --no-cpu > gpurun_out/r1b_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1b_launches.csv python bench.py --steps 8 --warmup 3 --no-cpu > gpurun_out/r1b_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'integrate_kernel|classify_kernel' -s 12 -c 4 -o gpurun_out/r1b_prof_k1 -f python bench.py --steps 8 --warmup 3 --no-cpu > gpurun_out/r1b_ncu2.log 2>&1
tail -2 gpurun_out/r1b_ncu2.log
