"""Token agreement of the CUDA path with the CPU oracle: golden fixtures and a C2-shaped batch (default model, RVQ 16384x192)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import meshgpt_pytorch_b200 as M
import synth  # tests/synth.py: pure-torch mesh generator (test + bench infrastructure, not product)
from helpers import make_pair

def rate(a, b):
    m = b != -1
    return float((a[m] == b[m]).float().mean())

g = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
for name in ("lfq_tri", "rvq_tri", "lfq_quad_ragged"):
    fx = torch.load(os.path.join(g, f"pipeline_{name}.pt"))
    m = M.MeshAutoencoder(**fx["cfg"]); m.load_reference_state_dict(fx["state_dict"]); m = m.cuda()
    c = m.tokenize(fx["vertices"].cuda(), fx["faces"].cuda()).cpu()
    print(f"fixture {name}: token agreement {rate(c, fx['codes']):.5f}")
for kind, cfg, nxny, n in (("tri", dict(use_residual_lfq=False), (20, 20), 8), ("quad", dict(quads=True), (20, 40), 8), ("tri", dict(), (20, 20), 8)):
    o, m = make_pair(**cfg)
    v, f = synth.grid_batch(kind, *nxny, list(range(n)))
    want = o.tokenize(v, f)
    got = m.tokenize(v.cuda(), f.cuda()).cpu()
    both = (got == want).all(-1)
    print(f"{kind} {cfg}: {n} x 800 faces, token agreement {rate(got, want):.5f}, tokens with both levels equal {float(both.float().mean()):.5f}")
