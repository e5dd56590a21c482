#!/bin/bash
# tile band kernel (75 KB per CTA): warps per CTA
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in "$@"; do
  cp variants/lib_$v.so hvqm4_b200/libhvqm4_b200.so
  for rep in 1 2; do
    echo "== $v dense 1024: $(timeout 120 python tools/profile_recon.py 1024 3 0 2>&1 | tail -1)"
  done
  echo "== $v dense 128: $(timeout 120 python tools/profile_recon.py 128 24 0 2>&1 | tail -1)"
  echo "== $v parity: $(timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k 'golden or band' 2>&1 | tail -1)"
done 2>&1 | tee gpurun_out/r02_tile_warps2_ab.txt
