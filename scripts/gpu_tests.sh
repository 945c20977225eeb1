#!/bin/bash
# Run every GPU test file in its own process (a CUDA fault in one must not poison the rest); logs -> gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
rc=0
for f in ${@:-tests/test_gpu_*.py}; do
  n=$(basename $f .py)
  timeout 600 python -m pytest $f -q -m gpu -rA --no-header -p no:cacheprovider > gpurun_out/$n.log 2>&1
  r=$?
  echo "== $n exit $r"; tail -n 25 gpurun_out/$n.log | cut -c1-220
  [ $r -ne 0 ] && rc=1
done
exit $rc
