"""Static SASS evidence of the in-tree library: per kernel, the counts of the Blackwell-specific instructions (tcgen05 MMA = UTC*MMA,
TMA = UTMALDG / UTMASTG / UBLKCP, tensor-memory loads / stores = LDTM / STTM, packed fp32 = FFMA2, cluster / DSMEM) and the total
instruction count.  usage: python tools/sass_static.py [path/to/liblpd_b200.so] > profiles/<tag>_sass_static.txt"""
import collections
import re
import subprocess
import sys
from pathlib import Path

so = sys.argv[1] if len(sys.argv) > 1 else str(Path(__file__).resolve().parent.parent / "lpd-net-pytorch_b200" / "liblpd_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
KEYS = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCMMA", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM", "UTCBAR", "UTCCP", "SYNCS", "FFMA2", "HFMA2", "HMMA", "FFMA", "UCGABAR", "LDS", "STS", "LDG", "STG", "RED", "ATOM"]
cur, per = None, collections.OrderedDict()
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur)
        per[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        per[cur]["_total"] += 1
        op = m.group(1)
        for k in KEYS:
            if op == k or (k in ("UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCMMA", "UTMALDG", "UTMASTG") and op.startswith(k)):
                per[cur][k] += 1
                break
tot = collections.Counter()
print(f"# {so}: {len(per)} kernels")
print(f"{'kernel':70s} {'instr':>7s}  Blackwell / memory opcodes")
for k, c in per.items():
    for kk, v in c.items():
        tot[kk] += v
    tags = " ".join(f"{kk}:{c[kk]}" for kk in KEYS if c[kk])
    print(f"{k[:70]:70s} {c['_total']:7d}  {tags}")
print("# totals: " + " ".join(f"{kk}:{tot[kk]}" for kk in KEYS if tot[kk]) + f" instructions:{tot['_total']}")
