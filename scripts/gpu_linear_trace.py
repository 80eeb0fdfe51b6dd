"""Timeline of one tcgen05 linear launch of the C2 step (cluster 0): MESHTOK_LINEAR_TRACE=k python scripts/gpu_linear_trace.py
k = ordinal of the linear launch within a step (0 = lin of layer 0, 1..4 = output linear + next lin, 5 = last output linear, 6 = pdc)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from helpers import make_pair
import synth  # tests/synth.py: pure-torch mesh generator (test + bench infrastructure, not product)
from meshgpt_pytorch_b200._lib import lib

o, m = make_pair(use_residual_lfq=False)
v, f = synth.grid_batch("tri", 20, 20, list(range(64)))
v, f = v.cuda(), f.cuda()
for _ in range(5):
    m.tokenize(v, f)
torch.cuda.synchronize()
buf = (C.c_longlong * 256)()
n = lib.meshtok_debug_linear_trace(buf, 256)
names = ["P issue", "P issued", "M wait acc", "M acc free", "M 1st stage", "E ready", "E acc full", "E stored"]
for rank in range(2):
    t0 = min(x for x in buf[rank * 128:(rank + 1) * 128] if x > 0)
    print(f"CTA {rank} (cycles since its first stamp / 1000)")
    for ev in range(8):
        row = [buf[(rank * 8 + ev) * 16 + u] for u in range(16)]
        print(f"  {names[ev]:12s} " + " ".join(f"{(x - t0) / 1000:7.1f}" if x > 0 else "      -" for x in row[:8]))
