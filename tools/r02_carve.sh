#!/bin/bash
# how much the band kernels depend on the size of L1: the shared-memory carve-out pinned (percent of 228 KB)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
for c in -1 86 100; do
  echo "== tile  carveout $c: $(HVQM4_BAND_CARVEOUT=$c timeout 120 python tools/profile_recon.py 1024 3 0 2>&1 | tail -1)"
done
for c in -1 44 58 72 86 100; do
  echo "== plain carveout $c: $(HVQM4_BAND_TILE=0 HVQM4_BAND_CARVEOUT=$c timeout 120 python tools/profile_recon.py 1024 3 0 2>&1 | tail -1)"
done
} 2>&1 | tee gpurun_out/r02_carve_ab.txt
