cd $GRAFT_REPO_ROOT
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2_tests.log 2>&1
tail -6 gpurun_out/r2_tests.log
python bench.py --steps 50 --warmup 5 --no-cpu > gpurun_out/r2_b.json 2>gpurun_out/r2_b.err
python - <<'PY'
import json
for l in open('gpurun_out/r2_b.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['fused_merge_path'])
PY
