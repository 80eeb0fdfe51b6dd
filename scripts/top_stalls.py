"""Top SASS instructions by warp-stall samples from `ncu --page source --csv` output, with the dominant stall reasons.
    python scripts/top_stalls.py gpurun_out/prof_<tag>_src_<kernel>.csv [N]"""
import csv
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
with open(path, newline="") as fh:
    rows = list(csv.reader(fh))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
for n, r in enumerate(rows[hdr_i + 1:]):
    if r and r[0] == "Kernel Name":          # next profiled launch: keep the first one only
        break
    if len(r) < len(hdr) or r[0] == "Address":
        continue
    samples = int(r[ix["# Samples"]] or 0)
    data.append((samples, n, r))
total = sum(d[0] for d in data)
print(f"total samples {total}, instructions {len(data)}")
agg = {h: sum(int(d[2][ix[h]] or 0) for d in data) for h in stall_cols}
print("stall totals:", ", ".join(f"{k[6:]}={v} ({100*v/max(total,1):.0f}%)" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
for samples, n, r in sorted(data, key=lambda d: -d[0])[:top]:
    st = sorted(((int(r[ix[h]] or 0), h[6:]) for h in stall_cols), reverse=True)[:3]
    print(f"{samples:7d} {100*samples/max(total,1):5.1f}%  #{n:5d} {r[ix['Source']].strip()[:70]:70s} " + " ".join(f"{h}={v}" for v, h in st if v))
