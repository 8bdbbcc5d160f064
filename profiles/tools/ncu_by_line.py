"""Instructions executed per source line: joins `ncu --page source --csv` (per SASS instruction) with the line table of
`nvdisasm -c -g` of the same cubin.  usage: ncu_by_line.py <src.csv> <nvdisasm.txt> <mangled-substring> [top]"""
import csv
import re
import sys
from collections import defaultdict

src_csv, dis, fn = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 60
# nvdisasm: offsets -> (file, line)
line_of = {}
cur = None
infn = False
for ln in open(dis, errors="replace"):
    if ln.startswith(".text."):
        infn = fn in ln
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        inl = " (inlined)" if "inlined at" in ln else ""
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        line_of[int(m.group(1), 16)] = (cur, m.group(2).strip())
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ia, ie, isrc, ismp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
base = None
per_line = defaultdict(lambda: [0, 0, 0])
tot = 0
for r in rows[2:]:
    if len(r) <= ie:
        continue
    a = int(r[ia], 16)
    if base is None:
        base = a
    off = a - base
    n = int(r[ie] or 0)
    smp = int(r[ismp] or 0)
    loc = line_of.get(off, (None, ""))[0]
    per_line[loc][0] += n
    per_line[loc][1] += smp
    per_line[loc][2] += 1
    tot += n
print("total warp instructions", tot)
for loc, (n, smp, k) in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{str(loc):40s} {n:12d} {100.0 * n / tot:6.2f}%  samples {smp:7d}  sass {k}")
