#!/bin/bash
# full GPU parity suite + band / row A/B on one box
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gpu_tests.log 2>&1
echo "gpu tests rc=$?" >> gpurun_out/r02_gpu_tests.log
tail -4 gpurun_out/r02_gpu_tests.log
for p in 0 1; do
  echo "== auto profile $p: $(timeout 200 python tools/profile_recon.py 1024 3 $p 2>&1 | tail -1)"
  echo "== row  profile $p: $(HVQM4_ROW=1 timeout 200 python tools/profile_recon.py 1024 3 $p 2>&1 | tail -1)"
done 2>&1 | tee gpurun_out/r02_ab.txt
