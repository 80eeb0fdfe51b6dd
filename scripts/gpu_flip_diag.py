"""Which derived-feature bin differs between the CUDA path and the oracle on the C4 test batch, and how close to a bin edge the
feature is: python scripts/gpu_flip_diag.py   (prints one line per face whose project_in row differs by more than 1e-4)."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from helpers import make_pair
import synth
from oracle import meshtok_oracle as O

o, m = make_pair(use_residual_lfq=False)
g = torch.Generator().manual_seed(1234)
nf = torch.randint(100, 801, (64,), generator=g).tolist()
v, f = synth.ragged_batch("tri", 20, 20, nf, range(0, 64))
codes, st = o.forward_codes(vertices=v, faces=f, return_stages=True)
m.keep_taps = True
got = m.tokenize(v.cuda(), f.cuda())
t = m.taps()
x0 = t["x0"].cpu()
diff = (x0 - st["project_in"]).abs().max(-1).values
bad = (diff > 1e-4).nonzero()[:, 0].tolist()
print(f"{len(bad)} of {x0.shape[0]} face rows differ by more than 1e-4 (max {float(diff.max()):.3e})")
fm = (f != -1).all(-1)
rows = fm.nonzero()                                   # (n, 2): mesh, face of every packed row
lo, hi = o.coor_range
b_idx = torch.arange(f.shape[0])[:, None, None]
fc = v[b_idx, f.masked_fill(~fm[:, :, None], 0).long()]
feats = O.derived_face_features(fc)
def frac_pos(x, lo_, hi_, n):
    u = (x - lo_) / (hi_ - lo_) * n - 0.5             # distance of u from the nearest rounding boundary (k + 0.5)
    return (u - torch.floor(u) - 0.5).abs()
for r in bad[:20]:
    b, fi = rows[r].tolist()
    ang, area, nrm = feats["angles"][b, fi], feats["area"][b, fi], feats["normals"][b, fi]
    print(f"row {r} (mesh {b}, face {fi}): |dx| {float(diff[r]):.3e}")
    print("   angles ", [f"{float(a):.7f}" for a in ang], " dist to bin edge (in bins):", [f"{float(d):.2e}" for d in frac_pos(ang, 0.0, math.pi, o.num_discrete_angle)])
    print("   area   ", f"{float(area):.7e}", " dist:", f"{float(frac_pos(area, 0.0, (hi - lo) ** 2, o.num_discrete_area)):.2e}")
    print("   normals", [f"{float(a):.7f}" for a in nrm], " dist:", [f"{float(d):.2e}" for d in frac_pos(nrm, lo, hi, o.num_discrete_normals)])
    print("   coords ", fc[b, fi].flatten().tolist())
