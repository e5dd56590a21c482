#!/bin/bash
# round 2: row kernel parity + A/B against the band kernel (+ role trace with variants/lib_rowtrace.so if present)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "row" > gpurun_out/r02_row_tests.log 2>&1
echo "row tests rc=$?" >> gpurun_out/r02_row_tests.log
tail -5 gpurun_out/r02_row_tests.log
for p in 0 1; do
  echo "== band profile $p: $(HVQM4_SWEEP=0 timeout 200 python tools/profile_recon.py 1024 3 $p 2>&1 | tail -1)"
  echo "== row  profile $p: $(HVQM4_ROW=1 timeout 200 python tools/profile_recon.py 1024 3 $p 2>&1 | tail -1)"
done 2>&1 | tee gpurun_out/r02_row_ab.txt
if [ -f variants/lib_rowtrace.so ]; then
  cp hvqm4_b200/libhvqm4_b200.so /tmp/keep.so; cp variants/lib_rowtrace.so hvqm4_b200/libhvqm4_b200.so
  for p in 0 1; do timeout 200 python tools/profile_row_trace.py 1024 $p 2>&1 | tail -16; done | tee gpurun_out/r02_row_trace.txt
  cp /tmp/keep.so hvqm4_b200/libhvqm4_b200.so
fi
