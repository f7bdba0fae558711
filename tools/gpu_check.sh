#!/bin/bash
# Runs every GPU parity test file in its own process (a trapped kernel kills only that file),
# then the micro-benchmarks.  Usage (on the GPU box): bash tools/gpu_check.sh [files...]
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
files="$@"
[ -z "$files" ] && files="tests/test_gpu_pf.py tests/test_gpu_crop.py tests/test_gpu_vit.py tests/test_gpu_tracker.py"
rc_all=0
for f in $files; do
  name=$(basename $f .py)
  timeout 600 python -m pytest $f -q -m gpu -x -s --tb=short > gpurun_out/$name.log 2>&1
  rc=$?
  echo "== $f rc=$rc"; tail -n 25 gpurun_out/$name.log
  [ $rc -ne 0 ] && rc_all=1
done
exit $rc_all
