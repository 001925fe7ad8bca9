#!/bin/bash
# One GPU-box session: build check, GPU parity tests (one process per group so a trapped kernel cannot
# poison the rest), smoke, a short bench.  Everything is logged under gpurun_out/.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
rc=0
run() {  # name, pytest args...
  local name=$1; shift
  timeout 600 python -m pytest "$@" -q -m gpu -p no:cacheprovider --tb=short > gpurun_out/pytest_$name.log 2>&1
  local r=$?
  echo "== $name: exit $r: $(tail -1 gpurun_out/pytest_$name.log)"
  if [ $r -ne 0 ]; then rc=1; grep -E "^(FAILED|ERROR|E  )" gpurun_out/pytest_$name.log | head -12; fi
}
run direct   tests/test_gpu_network.py -k "stem or depthwise or cpu_tensor"
run pw       tests/test_gpu_network.py -k "pointwise"
run head     tests/test_gpu_network.py -k "head"
run forward  tests/test_gpu_network.py -k "forward"
run detect   tests/test_gpu_detect.py
run nmslong  tests/test_gpu_nms_long.py
run match    tests/test_gpu_match_loss.py
run train    tests/test_gpu_train.py
run insitu   tests/test_gpu_train_insitu.py -s
run trainmod tests/test_gpu_train_modules.py
run metrics  tests/test_gpu_metrics.py
run gtbox    tests/test_gpu_gtbox.py
run nms120k  tests/test_gpu_zz_nms_oracle_120k.py
run nmsbig   tests/test_gpu_zz_nms_oracle_big.py
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; r=$?
echo "== smoke: exit $r: $(tail -1 gpurun_out/smoke.log)"; [ $r -ne 0 ] && rc=1
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; r=$?
echo "== bench: exit $r"; cat gpurun_out/bench.json; [ $r -ne 0 ] && { rc=1; tail -5 gpurun_out/bench.err; }
exit $rc
