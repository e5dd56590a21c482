"""Aggregates an `ncu --page source --csv --print-source cuda,sass` dump by CUDA source line.
    python tools/ncu_lines.py dump.csv [launch_index] [top_n]"""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
rows = list(csv.reader(open(path)))
# split into (file) sections; each begins with "File Path" row
sections, cur = [], None
for r in rows:
    if r and r[0] == "File Path":
        cur = {"file": r[1], "rows": [], "hdr": None}
        sections.append(cur)
    elif cur is not None:
        if r and r[0] == "Line No":
            cur["hdr"] = r
        elif cur["hdr"] and len(r) == len(cur["hdr"]):
            cur["rows"].append(r)
# sections repeat per launch; group by launch = count of first file occurrences
first = sections[0]["file"] if sections else None
launch, groups = -1, defaultdict(list)
for s in sections:
    if s["file"] == first:
        launch += 1
    groups[launch].append(s)


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return 0.0


agg = defaultdict(lambda: [0.0, 0.0, ""])
for s in groups[which]:
    h = s["hdr"]
    li, si = h.index("Line No"), h.index("Source")
    ie, sm = h.index("Instructions Executed"), h.index("# Samples")
    line, text = None, ""
    for r in s["rows"]:
        if r[li]:
            line, text = r[li], r[si]
        key = (s["file"].split("/")[-1], line)
        agg[key][0] += num(r[ie])
        agg[key][1] += num(r[sm])
        agg[key][2] = text
tot_i = sum(v[0] for v in agg.values()) or 1
tot_s = sum(v[1] for v in agg.values()) or 1
print(f"launch {which}: {tot_i:.0f} warp instructions, {tot_s:.0f} samples")
for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    print(f"{100 * v[0] / tot_i:6.2f}% inst {100 * v[1] / tot_s:6.2f}% samp  {f}:{l}  {v[2].strip()[:110]}")
