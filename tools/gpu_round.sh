#!/bin/bash
# One GPU-box visit: the -m gpu suite (with durations), the default bench line, the reference arm.  Usage: tools/gpu_round.sh TAG [tests|bench|all]
TAG=${1:-x}; WHAT=${2:-all}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
nproc > gpurun_out/${TAG}_nproc.txt; cat /sys/fs/cgroup/cpu.max >> gpurun_out/${TAG}_nproc.txt 2>/dev/null
if [ "$WHAT" = all ] || [ "$WHAT" = tests ]; then
  timeout 2400 python -m pytest tests -m gpu -q --durations=25 -p no:cacheprovider > gpurun_out/${TAG}_tests.log 2>&1
  echo "tests exit $?" >> gpurun_out/${TAG}_tests.log
  tail -45 gpurun_out/${TAG}_tests.log
fi
if [ "$WHAT" = all ] || [ "$WHAT" = bench ]; then
  timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
  echo "bench exit $?"; tail -c 6000 gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
  timeout 600 python bench.py --impl reference > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err
  echo "ref exit $?"; cat gpurun_out/${TAG}_bench_ref.json
fi
