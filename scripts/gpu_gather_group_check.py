"""Tokens of a 160-mesh ragged batch on the full-size RVQ model: written to a file (first run), compared with the file (second run).
Used by scripts/gpu_gather_group_ab.sh to check MESHTOK_GATHER_GROUP=1 (read once per process) against the default aggregation kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from helpers import make_pair
import synth  # tests/synth.py: pure-torch mesh generator (test + bench infrastructure, not product)

o, m = make_pair(use_residual_lfq=False)
g = torch.Generator().manual_seed(7)
counts = torch.randint(100, 801, (160,), generator=g).tolist()
v, f = synth.ragged_batch("tri", 20, 20, counts, range(160))
m.keep_taps = True
codes = m.tokenize(v.cuda(), f.cuda()).cpu()
taps = {k: t.cpu() for k, t in m.taps().items() if k.startswith("layer")}
path = sys.argv[1]
if os.path.exists(path):
    want = torch.load(path)
    same = torch.equal(codes, want["codes"])
    layers = all(torch.equal(taps[k], want["taps"][k]) for k in taps)
    print(f"MESHTOK_GATHER_GROUP={os.environ.get('MESHTOK_GATHER_GROUP', '0')}: codes identical {same}, {len(taps)} layer outputs bit-identical {layers}")
    sys.exit(0 if same and layers else 1)
torch.save(dict(codes=codes, taps=taps), path)
print("reference tokens written")
