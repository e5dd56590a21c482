#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for cap in 0 4096 12288 20480 30720; do
for p in 0 1; do
  echo "== band stage $cap profile $p: $(HVQM4_BAND=1 HVQM4_BAND_STAGE=$cap timeout 200 python tools/profile_recon.py 1024 3 $p 2>&1 | tail -1)"
done; done 2>&1 | tee gpurun_out/r02_stage_ab.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "band_kernel or stress" 2>&1 | tail -2
