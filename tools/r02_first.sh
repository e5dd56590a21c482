#!/bin/bash
# round 2, first GPU look at the sweep kernel: parity tests that select it, then band vs sweep on the headline workload
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sweep" > gpurun_out/r02_sweep_tests.log 2>&1
echo "sweep tests rc=$?" >> gpurun_out/r02_sweep_tests.log
tail -5 gpurun_out/r02_sweep_tests.log
B="--no-e2e --no-cpu-baseline --no-realistic --steps 10 --warmup 3"
HVQM4_SWEEP=0 timeout 300 python bench.py $B > gpurun_out/r02_band.json 2> gpurun_out/r02_band.err; echo "band rc=$?"
HVQM4_SWEEP=1 timeout 300 python bench.py $B > gpurun_out/r02_sweep_h1.json 2> gpurun_out/r02_sweep_h1.err; echo "sweep h1 rc=$?"
HVQM4_SWEEP=1 HVQM4_SWEEP_H=2 timeout 300 python bench.py $B > gpurun_out/r02_sweep_h2.json 2> gpurun_out/r02_sweep_h2.err; echo "sweep h2 rc=$?"
HVQM4_SWEEP=0 timeout 300 python bench.py $B --profile 1 > gpurun_out/r02_band_real.json 2> gpurun_out/r02_band_real.err; echo "band real rc=$?"
HVQM4_SWEEP=1 timeout 300 python bench.py $B --profile 1 > gpurun_out/r02_sweep_real.json 2> gpurun_out/r02_sweep_real.err; echo "sweep real rc=$?"
for f in gpurun_out/r02_band.json gpurun_out/r02_sweep_h1.json gpurun_out/r02_sweep_h2.json gpurun_out/r02_band_real.json gpurun_out/r02_sweep_real.json; do
  python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], round(d["value"]), d["roofline"]["frac"], d["roofline"]["kernel"], d["config"]["launches_per_step"])
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
