"""Executed-instruction profile of a kernel from `ncu --page source --csv`: contiguous SASS regions with the same execution
count, with their share of all executed warp instructions.   python scripts/inst_regions.py <src.csv> [min share %]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
min_share = float(sys.argv[2]) if len(sys.argv) > 2 else 2.0
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]
ix = {k: i for i, k in enumerate(h)}
data = []
for r in rows[hi + 1:]:
    if len(r) < len(h) or r[0] in ("Address", "Kernel Name"):
        break
    data.append(r)
col = "Instructions Executed"
tot = sum(int(r[ix[col]] or 0) for r in data)
print(f"{tot} warp instructions executed, {len(data)} SASS instructions")
prev, start = None, 0
for n, r in enumerate(data + [None]):
    v = int(r[ix[col]] or 0) if r else None
    if v != prev:
        if prev is not None and prev * (n - start) > tot * min_share / 100:
            print(f"#{start:5d}-{n - 1:5d} {n - start:4d} instrs x {prev:9d} = {100 * prev * (n - start) / tot:5.1f} %  {data[start][ix['Source']].strip()[:70]}")
        prev, start = v, n
