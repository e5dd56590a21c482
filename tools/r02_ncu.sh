#!/bin/bash
# round 2 evidence: ncu --set full of the four dense reconstruction schedules on one box (256 streams; the replay
# launches come after the 16 recording launches; two launches each = a P and a B picture step)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cap() { # name, kernel regex, env...
  name=$1; rx=$2; shift 2
  env "$@" timeout 600 ncu --set full --import-source on --clock-control none -k regex:$rx -s 17 -c 2 -o gpurun_out/$name -f \
    python tools/profile_recon.py 256 1 ${PROFILE:-0} > gpurun_out/$name.log 2>&1; echo "$name rc=$?"
}
cap r02_band_dense        recon_band_kernel  HVQM4_BAND=3
cap r02_band_tile_dense   recon_band_kernel  HVQM4_BAND=1 HVQM4_BAND_TILE=1
cap r02_row_dense         recon_row_kernel   HVQM4_ROW=1
cap r02_sweep_dense       recon_sweep_kernel HVQM4_SWEEP=1
PROFILE=1 cap r02_row_realistic   recon_row_kernel   HVQM4_ROW=1
ls -la gpurun_out/*.ncu-rep
