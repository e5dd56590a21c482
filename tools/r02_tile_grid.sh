#!/bin/bash
# band kernel, plain (auto CTA count) against the tile variant (band assembled in shared memory), over the grid size
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in "$@"; do
  cp variants/lib_$v.so hvqm4_b200/libhvqm4_b200.so
  for S in 16 32 64 128 256 512 1024; do
    R=$(( 4096 / S )); [ $R -gt 24 ] && R=24; [ $R -lt 3 ] && R=3
    for p in 0; do
      echo "== $v S=$S plain profile $p: $(HVQM4_BAND=1 timeout 200 python tools/profile_recon.py $S $R $p 2>&1 | tail -1)"
      echo "== $v S=$S tile  profile $p: $(HVQM4_BAND=1 HVQM4_BAND_TILE=1 timeout 200 python tools/profile_recon.py $S $R $p 2>&1 | tail -1)"
    done
  done
  echo "== $v parity: $(timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k 'band_tile or golden' 2>&1 | tail -1)"
done 2>&1 | tee gpurun_out/r02_tile_grid.txt
