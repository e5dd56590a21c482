#!/bin/bash
# end-of-round bench line at N GPUs of one box (default settings: what the driver runs)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 \
  > gpurun_out/r02_bench_n${N}_final.json 2> gpurun_out/r02_bench_n${N}_final.err; echo "rc=$?"
tail -3 gpurun_out/r02_bench_n${N}_final.err
python - "gpurun_out/r02_bench_n${N}_final.json" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
e=d["e2e"]
print("value", round(d["value"]), "strong", round(d.get("strong_scaling",{}).get("value",0)), "e2e", round(e["value"]), "ceiling", e.get("pcie_ceiling"), "frac", e.get("frac_of_ceiling"), "host", round(d["e2e_host_entropy"]["value"]), d["details"]["host_threads_per_gpu"], d["details"].get("bitstreams_fetched_by_gpu"))
PY
nproc
