#!/bin/bash
# compute-sanitizer on the 64^3 smoke (4 labelled frames + a ray-cast, __graft_entry__.smoke): memcheck and racecheck
# (shared-memory queues, the mbarrier-completed TMA copy, three streams with rotating contexts).  Run on a GPU box:
#   gpurun -- bash tools/sanitize_smoke.sh      -> gpurun_out/sanitizer_{memcheck,racecheck}.log
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 3 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_$tool.log 2>&1
  echo "== $tool rc=$?"; tail -4 gpurun_out/sanitizer_$tool.log
done
