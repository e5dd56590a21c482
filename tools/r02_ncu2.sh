#!/bin/bash
# ncu --set full of the default dense kernel after the shared-memory diet (tile variant, 16 warps, 75 KB + queues per CTA)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:recon_band_kernel -s 17 -c 2 -o gpurun_out/r02_band_tile16_dense -f \
    python tools/profile_recon.py 256 1 0 > gpurun_out/r02_band_tile16_dense.log 2>&1; echo "rc=$?"
ls -la gpurun_out/r02_band_tile16_dense.ncu-rep
