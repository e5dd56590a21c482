#!/bin/bash
# end to end, GPU entropy stage with a share of the streams parsed by the host threads (HVQM4BatchSetHostShare)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "host cores: $(nproc)"
echo "== parity: $(timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k 'gpu_entropy or pipelined or staggered or registered or switch' 2>&1 | tail -1)"
for share in 0 64 96 128 160 192 256; do
  for rep in 1 2; do
    echo "== share $share dense:     $(timeout 120 python tools/profile_e2e.py 1024 16 1 0 4 1 $share 2>&1 | grep fps | tail -1 | cut -c1-100)"
  done
done
for share in 0 128; do
  echo "== share $share realistic: $(timeout 120 python tools/profile_e2e.py 1024 16 1 1 4 1 $share 2>&1 | grep fps | tail -1 | cut -c1-100)"
  echo "== share $share dense, tile band kernel behind the parser: $(HVQM4_BAND_BEHIND_PARSER=7 timeout 120 python tools/profile_e2e.py 1024 16 1 0 4 1 $share 2>&1 | grep fps | tail -1 | cut -c1-100)"
done
echo "== recon dense (tile default): $(timeout 120 python tools/profile_recon.py 1024 4 0 2>&1 | tail -1)"
echo "== recon dense (HVQM4_BAND_TILE=0): $(HVQM4_BAND_TILE=0 timeout 120 python tools/profile_recon.py 1024 4 0 2>&1 | tail -1)"
} 2>&1 | tee gpurun_out/r02_share_ab.txt
