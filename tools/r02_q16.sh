#!/bin/bash
# 16-bit queue entries over the nest table: tile band kernel at 75 KB per CTA (164 KB carve-out, 92 KB of L1)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
for rep in 1 2; do
echo "== tile   dense: $(timeout 120 python tools/profile_recon.py 1024 3 0 2>&1 | tail -1)"
echo "== plain  dense: $(HVQM4_BAND_TILE=0 timeout 120 python tools/profile_recon.py 1024 3 0 2>&1 | tail -1)"
done
echo "== tile   dense carve 58 (132 KB: one CTA per SM?): $(HVQM4_BAND_CARVEOUT=58 timeout 120 python tools/profile_recon.py 1024 3 0 2>&1 | tail -1)"
echo "== tile   dense carve 72: $(HVQM4_BAND_CARVEOUT=72 timeout 120 python tools/profile_recon.py 1024 3 0 2>&1 | tail -1)"
echo "== tile   dense carve 86: $(HVQM4_BAND_CARVEOUT=86 timeout 120 python tools/profile_recon.py 1024 3 0 2>&1 | tail -1)"
echo "== band (tile) realistic: $(HVQM4_BAND=1 timeout 120 python tools/profile_recon.py 1024 3 1 2>&1 | tail -1)"
echo "== auto realistic: $(timeout 120 python tools/profile_recon.py 1024 3 1 2>&1 | tail -1)"
for S in 16 64 256; do
echo "== tile S=$S dense: $(timeout 120 python tools/profile_recon.py $S 24 0 2>&1 | tail -1)"
done
echo "== parity: $(timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -1)"
} 2>&1 | tee gpurun_out/r02_q16_ab.txt
