#!/bin/bash
# one GPU end to end: host share x who fetches the bitstreams (host copy into the pinned arena | GPU gather from registered memory)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "host cores: $(nproc)"
for share in 0 96 128 160; do
  for rep in 1 2; do
    echo "== share $share host copy: $(HVQM4_BATCH_TRACE=1 timeout 120 python tools/profile_e2e.py 1024 16 1 0 4 1 $share 2>&1 | grep -E 'fps|submitting' | cut -c1-170 | tr '\n' ' ')"
    echo "== share $share gather:    $(HVQM4_BATCH_TRACE=1 E2E_REGISTER=1 timeout 120 python tools/profile_e2e.py 1024 16 1 0 4 1 $share 2>&1 | grep -E 'fps|submitting' | cut -c1-170 | tr '\n' ' ')"
  done
done
} 2>&1 | tee gpurun_out/r02_share_gather_ab.txt
