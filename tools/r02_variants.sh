#!/bin/bash
# A/B of prebuilt library variants (variants/lib_<name>.so) on one box: band kernel plain / tile, dense and realistic
cd "$(dirname "$0")/.."
cp hvqm4_b200/libhvqm4_b200.so /tmp/keep.so
for v in "$@"; do
  cp variants/lib_$v.so hvqm4_b200/libhvqm4_b200.so
  for p in 0 1; do
    echo "== $v plain profile $p: $(HVQM4_BAND=1 timeout 200 python tools/profile_recon.py 1024 3 $p 2>&1 | tail -1)"
    echo "== $v tile  profile $p: $(HVQM4_BAND=1 HVQM4_BAND_TILE=1 timeout 200 python tools/profile_recon.py 1024 3 $p 2>&1 | tail -1)"
  done
  echo "== $v parity: $(timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k 'golden_cfg5 or many_streams' 2>&1 | tail -1)"
done
cp /tmp/keep.so hvqm4_b200/libhvqm4_b200.so
