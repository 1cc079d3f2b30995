#!/usr/bin/env python
"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) next to the live per-kernel times of a bench
line.  usage: launch_summary.py launches.csv bench.json out.md "<command>" """
import csv
import json
import sys
from collections import OrderedDict

csv_path, bench_path, out, cmd = sys.argv[1:5]
rows = [r for r in csv.reader(open(csv_path)) if len(r) > 14 and r[0].isdigit()]
agg = OrderedDict()
for r in rows:
    name, ns = r[4], float(r[14].replace(",", ""))
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1; a[1] += ns / 1e6
tot = sum(a[1] for a in agg.values())
L = ["# ncu launch list", "", "`%s`" % cmd, "",
     "%d launches; times under ncu are serialised and cold-cache: compare shares, not absolutes.  Raw list: `%s`." % (
         len(rows), csv_path.split("/")[-1]), "",
     "| kernel | launches | total ms | avg ms | share |", "|---|---|---|---|---|"]
for k, (n, ms) in agg.items():
    L.append("| `%s` | %d | %.3f | %.3f | %.1f %% |" % (k, n, ms, ms / n, 100 * ms / tot))
libk = [k for k in agg if not k.startswith("b2c_")]
L += ["", "Kernels that are not ours on the list: %s." % (", ".join("`%s`" % k for k in libk) if libk else "none")]
line = json.loads(open(bench_path).read().strip().splitlines()[-1])
live = line["roofline"]["kernel_ms_per_step"]
per_step = line["roofline"].get("pipeline_launches_per_step", 1)
enc = {}
for k, ms in live.items():
    # the live profile folds the parse kernels into one key
    cands = [a for a in agg if a == k or (k == "b2c_lz_parse_kernel" and a.startswith("b2c_lz_parse"))]
    if cands:
        n, t = agg[cands[0]]
        enc[k] = (t / n * per_step, ms)
st, sl = sum(v[0] for v in enc.values()), sum(v[1] for v in enc.values())
L += ["", "Encode pipeline (%d launch(es) of each kernel per step): share of the step under ncu next to the share measured "
      "live by `bench.py` with CUDA events (`%s`, `roofline.kernel_ms_per_step`):" % (per_step, bench_path.split("/")[-1]), "",
      "| kernel | ncu ms/step | ncu share | live ms/step | live share |", "|---|---|---|---|---|"]
for k, (a, b) in enc.items():
    L.append("| `%s` | %.3f | %.1f %% | %.3f | %.1f %% |" % (k, a, 100 * a / st, b, 100 * b / sl))
open(out, "w").write("\n".join(L) + "\n")
print("wrote", out)
