"""Per-kernel summary of an ncu launch list (--metrics gpu__time_duration.sum --csv).
usage: python tools/launch_summary.py gpurun_out/r02_launches_a.csv [header line ...] > profiles/r02_launches_a_summary.txt"""
import collections
import csv
import re
import statistics
import sys

rows = []
with open(sys.argv[1]) as fh:
    lines = [l for l in fh if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    us = v / 1e3 if unit in ("ns", "nsecond") else v if unit in ("us", "usecond") else v * 1e3
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    name = re.sub(r"^void ", "", name).replace("wsegk::", "")
    rows.append((name, us))
by = collections.defaultdict(list)
for n, us in rows:
    by[n].append(us)
ours = {n: v for n, v in by.items() if not n.startswith(("at::", "void at::", "ncclDev", "vectorized", "elementwise", "nchw", "cunn", "upsample", "reduce_", "index", "triu", "Cat", "cat", "fill", "direct_copy"))}
tot = sum(sum(v) for v in ours.values())
for h in sys.argv[2:]:
    print(h)
print(f"({len(rows)} launches captured, {sum(len(v) for v in ours.values())} of them libwsegk kernels; cold-cache, serialised launches: shares, not absolute step times)\n")
for n, v in sorted(ours.items(), key=lambda kv: -sum(kv[1])):
    print(f"{n:46s} launches {len(v):4d}  median {statistics.median(v):8.1f} us  total {sum(v):9.1f} us  share {100 * sum(v) / tot:5.1f} %")
