cd $GRAFT_REPO_ROOT
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 100 --warmup 5 --no-merge > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err
python - <<'PY'
import json
for l in open('gpurun_out/r2_bench_n8.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['config']['z_slabs'], d['e2e']['ms_per_step']); print(d['config']['slab_calibration_ms']); print(d['per_rank'])
PY
tail -3 gpurun_out/r2_bench_n8.err
