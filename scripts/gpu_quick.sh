#!/bin/bash
# A short GPU session: the given pytest targets (one process each), logs under gpurun_out/.
#   bash scripts/gpu_quick.sh name1:"pytest args" name2:"pytest args" ...
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
rc=0
for spec in "$@"; do
  name=${spec%%:*}; args=${spec#*:}
  timeout 900 python -m pytest $args -q -m gpu -p no:cacheprovider --tb=short > gpurun_out/pytest_$name.log 2>&1
  r=$?
  echo "== $name: exit $r: $(tail -1 gpurun_out/pytest_$name.log)"
  if [ $r -ne 0 ]; then rc=1; grep -E "^(FAILED|ERROR|E  )" gpurun_out/pytest_$name.log | cut -c1-300 | head -12; fi
done
exit $rc
