"""Per-SASS-instruction execution counts with source lines: joins `ncu --page source --csv` with `nvdisasm -c -g`.
usage: ncu_stream.py <src.csv> <nvdisasm.txt> <function-substring> > stream.txt"""
import csv
import re
import sys

src_csv, dis, fn = sys.argv[1], sys.argv[2], sys.argv[3]
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
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        line_of[int(m.group(1), 16)] = (cur, m.group(2).strip())
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ia, ie, ismp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
base = None
for r in rows[2:]:
    if len(r) <= ie:
        continue
    a = int(r[ia], 16)
    if base is None:
        base = a
    off = a - base
    loc, ins = line_of.get(off, ((None, 0), ""))
    print(f"{off:06x} {int(r[ie] or 0):10d} {int(r[ismp] or 0):6d} {loc[0]}:{loc[1]} {ins}")
