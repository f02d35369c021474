"""Duration, DRAM bytes and achieved DRAM rate of every captured launch of an ncu --set full report, against the measured HBM
peak.  usage: python tools/hbm_summary.py rep.ncu-rep [header line ...]"""
import csv, io, json, os, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
peak = 6445.0
try:
    peak = float(json.load(open(os.path.join(root, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
def val(r, k):
    v = float(r[idx[k]].replace(",", ""))
    u = units[idx[k]]
    scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3, "us": 1.0, "ns": 1e-3}.get(u, 1.0)
    return v * scale
for h in sys.argv[2:]:
    print(h)
for r in rows[2:]:
    name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("wsegk::", "")
    us = val(r, "gpu__time_duration.sum")
    rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
    gbs = (rd + wr) / us / 1e3
    busy = r[idx["sm__throughput.avg.pct_of_peak_sustained_elapsed"]] if "sm__throughput.avg.pct_of_peak_sustained_elapsed" in idx else "?"
    print(f"{name:34s} {us:7.1f} us  dram read {rd / 1e6:7.2f} MB write {wr / 1e6:7.2f} MB  -> {gbs:6.0f} GB/s ({100 * gbs / peak:4.1f} % of {peak:.0f})  sm busy {busy} %")
