#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "band_tile" > gpurun_out/r02_tile_tests.log 2>&1
echo "tile tests rc=$?" >> gpurun_out/r02_tile_tests.log
tail -4 gpurun_out/r02_tile_tests.log
for rep in 1 2; do
for p in 0 1; do
  echo "== band                 profile $p: $(HVQM4_BAND=1 timeout 200 python tools/profile_recon.py 1024 3 $p 2>&1 | tail -1)"
  echo "== band tile 8 rows     profile $p: $(HVQM4_BAND=1 HVQM4_BAND_TILE=1 timeout 200 python tools/profile_recon.py 1024 3 $p 2>&1 | tail -1)"
  echo "== band tile 4 rows     profile $p: $(HVQM4_BAND=1 HVQM4_BAND_TILE=1 HVQM4_BAND_ROWS=4 timeout 200 python tools/profile_recon.py 1024 3 $p 2>&1 | tail -1)"
  echo "== band 8-row CTA, 4-row bands  $p: $(HVQM4_BAND=1 HVQM4_BAND_ROWS=4 timeout 200 python tools/profile_recon.py 1024 3 $p 2>&1 | tail -1)"
done; done 2>&1 | tee gpurun_out/r02_tile_ab.txt
