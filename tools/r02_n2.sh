#!/bin/bash
# N-GPU end to end: host share of the entropy stage and who fetches the bitstreams, same box, with the submitting thread's trace
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
run() {
  tag=$1; shift
  HVQM4_BATCH_TRACE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 6 --warmup 3 \
    --no-realistic --no-cpu-baseline --no-e2e-host "$@" > gpurun_out/r02_n${N}_$tag.json 2> gpurun_out/r02_n${N}_$tag.err; echo "$tag rc=$?"
  grep "submitting thread" gpurun_out/r02_n${N}_$tag.err | tail -2
  python - "gpurun_out/r02_n${N}_$tag.json" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
e=d["e2e"]
print("  value", round(d["value"]), "e2e", round(e["value"]), "share", e.get("host_share_streams"), "ceiling", e.get("pcie_ceiling"), "frac", e.get("frac_of_ceiling"), "threads", d["details"]["host_threads_per_gpu"], "gather", d["details"].get("bitstreams_fetched_by_gpu"))
PY
}
{
nproc; free -g | head -2; nvidia-smi topo -m | head -12
run share0 --host-share 0
run auto
run share0_gather --host-share 0 --gather on
run auto_gather --gather on
} 2>&1 | tee gpurun_out/r02_n${N}_ab.txt
