#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
nvidia-smi --query-gpu=name --format=csv,noheader; python -c "import torch; p=torch.cuda.get_device_properties(0); print('async engines', getattr(p,'async_engine_count', None))" 2>/dev/null
for split in 1 2 4; do
  for rep in 1 2; do
    echo "== d2h split $split, share 128: $(HVQM4_D2H_SPLIT=$split timeout 120 python tools/profile_e2e.py 1024 16 1 0 4 1 128 2>&1 | grep -E 'fps' | cut -c1-100)"
  done
done
echo "== parity split 2: $(HVQM4_D2H_SPLIT=2 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k 'pipelined or staggered' 2>&1 | tail -1)"
HVQM4_D2H_SPLIT=2 HVQM4_BATCH_TIMELINE=1 python tools/profile_e2e.py 1024 16 1 0 4 1 128 2>&1 | tail -14
} 2>&1 | tee gpurun_out/r02_d2h_split_ab.txt
