"""Per-kernel counts of the Blackwell-specific SASS mnemonics in the shipped library: tcgen05 MMA (UTC*MMA), TMEM loads /
stores (LDTM / STTM), tcgen05.commit (UTCBAR), bulk copies (UBLKCP) and tensor-map TMA (UTMALDG / UTMASTG).
usage: python tools/sass_summary.py [lib.so] > profiles/r02_sass_summary.txt   (cuobjdump only; runs without a GPU)"""
import collections
import os
import re
import subprocess
import sys

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(root, "wseg_b200", "csrc", "libwsegk.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip() or n
PAT = [("UTC*MMA", r"\bUTC[A-Z]*MMA"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"), ("UTCBAR", r"\bUTCBAR"),
       ("UBLKCP", r"\bUBLKCP"), ("UTMALDG", r"\bUTMALDG"), ("UTMASTG", r"\bUTMASTG"), ("SYNCS", r"\bSYNCS"),
       ("HMMA/IMMA", r"\b[HI]MMA\b"), ("instr", r"^\s+/\*[0-9a-f]{4,}\*/")]
counts, order, cur = collections.defaultdict(collections.Counter), [], None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        order.append(cur)
        continue
    if cur is None:
        continue
    for name, pat in PAT:
        if re.search(pat, line):
            counts[cur][name] += 1
print(f"# {os.path.relpath(lib, root)}: {len(order)} kernels, sm_100a SASS (cuobjdump -sass); counts of instructions per kernel")
print("# " + "  ".join(f"{n:>9s}" for n, _ in PAT) + "  kernel")
tot = collections.Counter()
for fn in sorted(order, key=lambda f: -counts[f]["UTC*MMA"] * 10**6 - counts[f]["UBLKCP"] * 10**3 - counts[f]["instr"]):
    c = counts[fn]
    tot.update(c)
    name = re.sub(r"\(.*", "", demangle(fn))
    print("  " + "  ".join(f"{c[n]:9d}" for n, _ in PAT) + "  " + name)
print("  " + "  ".join(f"{tot[n]:9d}" for n, _ in PAT) + "  TOTAL")
