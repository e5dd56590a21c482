"""Sums dram__bytes_read.sum + dram__bytes_write.sum and gpu__time_duration.sum per step from an
`ncu --csv --metrics ...` log of tools/profile_recon.py and prints / merges the result into
profiles/r02_traffic.json, stamped with the hash of the kernel sources it was captured from (bench.py only quotes a capture
of the sources it runs).
    python tools/ncu_traffic.py log.csv key launches_per_step first_launch n_steps"""
import csv
import json
import os
import sys

path, key = sys.argv[1], sys.argv[2]
per_step, first, n_steps = int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
rows = [r for r in csv.reader(open(path)) if len(r) > 10]
hdr = rows[0]
idi, mi, vi, ui, ki = hdr.index("ID"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("Kernel Name")
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "usecond": 1e-6, "nsecond": 1e-9, "msecond": 1e-3}
launch = {}
for r in rows[1:]:
    d = launch.setdefault(int(r[idi]), {"kernel": r[ki].split("(")[0][-28:]})
    d[r[mi]] = float(r[vi].replace(",", "")) * scale.get(r[ui], 1)
ids = sorted(launch)[first:first + per_step * n_steps]
tot_b = sum(launch[i].get("dram__bytes_read.sum", 0) + launch[i].get("dram__bytes_write.sum", 0) for i in ids)
tot_t = sum(launch[i].get("gpu__time_duration.sum", 0) for i in ids)
by_kernel = {}
for i in ids:
    k = launch[i]["kernel"]
    e = by_kernel.setdefault(k, [0, 0.0, 0.0])
    e[0] += 1
    e[1] += launch[i].get("gpu__time_duration.sum", 0)
    e[2] += launch[i].get("dram__bytes_read.sum", 0) + launch[i].get("dram__bytes_write.sum", 0)
print(f"{key}: {len(ids)} launches = {n_steps} steps; DRAM {tot_b / n_steps / 1e6:.1f} MB/step, {tot_t / n_steps * 1e6:.1f} us/step (ncu, cold cache, serialised)")
for k, e in by_kernel.items():
    print(f"   {k}: {e[0]} launches, {100 * e[1] / tot_t:.1f}% of time, {e[2] / e[0] / 1e6:.1f} MB DRAM per launch")
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root)
import bench  # noqa: E402  (kernel_sources_sha)
out = os.path.join(root, "profiles", "r02_traffic.json")
try:
    t = json.load(open(out))
    if t.get("kernel_sources_sha") != bench.kernel_sources_sha():
        raise ValueError("stale capture")
except (OSError, ValueError):
    t = {"how": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none on tools/profile_recon.py; "
                "bytes summed over the launches of one step (one picture per stream), averaged over the 16 steps of a GOP replay",
         "dram_bytes_per_step": {}, "ncu_us_per_step": {}}
t["dram_bytes_per_step"][key] = tot_b / n_steps
t["ncu_us_per_step"][key] = tot_t / n_steps * 1e6
t["kernel_sources_sha"] = bench.kernel_sources_sha()
json.dump(t, open(out, "w"), indent=1)
