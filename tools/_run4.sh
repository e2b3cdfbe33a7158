cd $GRAFT_REPO_ROOT
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 100 --warmup 5 --no-merge > gpurun_out/r2_bench_n4.json 2> gpurun_out/r2_bench_n4.err
tail -c 2500 gpurun_out/r2_bench_n4.json; tail -3 gpurun_out/r2_bench_n4.err
