"""opcode histogram of an `ncu --page source --csv` export: python tools/sass_hist.py file.csv"""
import csv, collections, re, sys
rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
hdr = rows[h]; ia = hdr.index("Source"); ii = hdr.index("Instructions Executed"); isamp = hdr.index("# Samples")
ops = collections.Counter(); samp = collections.Counter(); tot = 0
for r in rows[h + 1:]:
    if len(r) <= ii or not r[ii].isdigit(): continue
    m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[ia].strip())
    op = m.group(2).split('.')[0] if m else r[ia][:10]
    n = int(r[ii]); ops[op] += n; samp[op] += int(r[isamp] or 0); tot += n
print("total warp instr", tot)
for op, n in ops.most_common(28): print(f"{op:12s} {n:12d} {100*n/tot:5.1f}%  samples {samp[op]}")
