"""Summarise the SASS source page of one kernel in an .ncu-rep: executed warp-instructions and stall samples per region.
usage: python tools/ncu_src_summary.py <report.ncu-rep> <kernel-regex> [top_n]"""
import csv, subprocess, sys, io, re, collections
rep, kre = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", f"regex:{kre}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hi = [i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r][0]
hdr = rows[hi]
si, ei, wi = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
out = []
for r in rows[hi + 1:]:
    try:
        out.append((r[0][-5:], r[si].strip(), int(r[ei]), int(r[wi])))
    except (ValueError, IndexError):
        pass
tot_i, tot_s = sum(o[2] for o in out), sum(o[3] for o in out)
print(f"total warp-inst {tot_i}  stall samples {tot_s}")
print("--- top stall lines")
for i in sorted(range(len(out)), key=lambda i: -out[i][3])[:topn]:
    o = out[i]
    print(f"{i:5d} {o[0]} {o[1][:70]:70s} exec {o[2]:10d} stall {o[3]:6d} ({100 * o[3] / tot_s:4.1f}%)")
