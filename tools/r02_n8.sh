#!/bin/bash
# 8-GPU end-to-end: bitstreams copied by host threads vs fetched by the GPU, same box
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-8}
for g in off on; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 6 --warmup 3 \
    --gather $g --no-realistic > gpurun_out/r02_bench_n${N}_gather_$g.json 2> gpurun_out/r02_bench_n${N}_gather_$g.err; echo "gather $g rc=$?"
  python - "gpurun_out/r02_bench_n${N}_gather_$g.json" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
e=d["e2e"]
print("value", round(d["value"]), "strong", round(d.get("strong_scaling",{}).get("value",0)), "e2e", round(e["value"]), "ceiling", e.get("pcie_ceiling"), "frac", e.get("frac_of_ceiling"), "host", round(d["e2e_host_entropy"]["value"]), d["details"]["host_threads_per_gpu"])
PY
done
nproc; free -g | head -2
