"""Phase timeline of one neighbour-aggregation launch of the C2 step: MESHTOK_GATHER_TRACE=k python scripts/gpu_gather_trace.py
k = layer (0..4).  Per CTA: start offset, prologue / CSR staging / wait for the x copies / row loop in cycles, total ns, SM."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from helpers import make_pair
import synth
from meshgpt_pytorch_b200._lib import lib

o, m = make_pair(use_residual_lfq=False)
v, f = synth.grid_batch("tri", 20, 20, list(range(64)))
v, f = v.cuda(), f.cuda()
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
for _ in range(5):
    flush.zero_()
    m.tokenize(v, f)
torch.cuda.synchronize()
buf = (C.c_longlong * 8192)()
n = lib.meshtok_debug_gather_trace(buf, 8192)
rows = [[buf[c * 8 + i] for i in range(8)] for c in range(1024) if buf[c * 8] > 0]
t0 = min(r[0] for r in rows)
t1 = max(r[6] for r in rows)
print(f"layer {os.environ.get('MESHTOK_GATHER_TRACE')}: {len(rows)} CTAs, kernel span {(t1 - t0) / 1e3:.2f} us (globaltimer)")
import statistics as st
def col(fn): return [fn(r) for r in rows]
for name, fn in (("start offset ns", lambda r: r[0] - t0), ("waits slices 1..", lambda r: r[2]), ("prologue + CSR", lambda r: r[3] - r[1]), ("wait copies", lambda r: r[4] - r[3]),
                 ("row loop", lambda r: r[5] - r[4]), ("total cycles", lambda r: r[5] - r[1]), ("total ns", lambda r: r[6] - r[0])):
    c = col(fn)
    print(f"  {name:16s} min {min(c):8d}  median {int(st.median(c)):8d}  max {max(c):8d}")
sms = {}
for r in rows:
    sms.setdefault(r[7], []).append((r[0] - t0, r[6] - t0))
print("  CTAs per SM:", sorted({len(x) for x in sms.values()}), " SMs used:", len(sms))
late = sorted(rows, key=lambda r: r[0])[-5:]
print("  last CTAs to start (ns after the first):", [(r[0] - t0, r[6] - r[0]) for r in late])
