#!/bin/bash
# TMA small-box rate; band-kernel regression check (round-1 library vs head, auto mode without the sweep kernel)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 tools/ubench/tma_box > gpurun_out/r02_tma_box.txt 2>&1; echo "tma_box rc=$?"
cp hvqm4_b200/libhvqm4_b200.so /tmp/keep.so
for v in r1 head; do
  cp variants/lib_$v.so hvqm4_b200/libhvqm4_b200.so
  for p in 0 1; do
    echo "== $v profile $p: $(HVQM4_SWEEP=0 timeout 200 python tools/profile_recon.py 1024 3 $p 2>&1 | tail -1)"
  done
done > gpurun_out/r02_regress.txt 2>&1
cp /tmp/keep.so hvqm4_b200/libhvqm4_b200.so
cat gpurun_out/r02_regress.txt
tail -70 gpurun_out/r02_tma_box.txt
