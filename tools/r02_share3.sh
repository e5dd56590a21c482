#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "host cores: $(nproc); $(grep -m1 'model name' /proc/cpuinfo)"
for share in 0 128 160 192; do
  for rep in 1 2; do
    echo "== share $share: $(HVQM4_BATCH_TRACE=1 timeout 120 python tools/profile_e2e.py 1024 16 1 0 4 1 $share 2>&1 | grep -E 'fps|submitting|parser' | cut -c1-220 | tr '\n' ' ')"
  done
done
echo "== no read-back, share 0: $(timeout 120 python tools/profile_e2e.py 1024 16 0 0 4 1 0 2>&1 | grep -E 'fps' | cut -c1-120)"
echo "== no read-back, share 128: $(timeout 120 python tools/profile_e2e.py 1024 16 0 0 4 1 128 2>&1 | grep -E 'fps' | cut -c1-120)"
} 2>&1 | tee gpurun_out/r02_share3_ab.txt
