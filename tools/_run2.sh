cd $GRAFT_REPO_ROOT
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/bench_fuse_sharded.py --gpus 2 --dims 256 256 256 --bins 32 --frames 12 --check > gpurun_out/r2_fuse_sharded_n2_check.json 2> gpurun_out/r2_fuse_n2.err
tail -2 gpurun_out/r2_fuse_sharded_n2_check.json; tail -5 gpurun_out/r2_fuse_n2.err | cut -c1-300
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 100 --warmup 5 --no-merge > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
python - <<'PY'
import json
for l in open('gpurun_out/r2_bench_n2.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['config']['z_slabs'], d['e2e']['ms_per_step']); print(d['per_rank'])
PY
