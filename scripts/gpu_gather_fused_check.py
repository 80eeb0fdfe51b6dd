"""GPU: the aggregation fused into the SAGE output linears (MESHTOK_GATHER_FUSED=1, default) against the separate aggregation launches
(=0) on the full-size RVQ model: every layer output and every code must be bit-identical (same neighbour order, same rounding, same
MMA order).  The switch is read once per process, hence subprocesses."""
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCRIPT = r"""
import sys, torch
sys.path.insert(0, {root!r}); sys.path.insert(0, {tests!r})
from helpers import make_pair
import synth
o, m = make_pair(use_residual_lfq=False)
g = torch.Generator().manual_seed(5)
counts = torch.randint(100, 801, (48,), generator=g).tolist() + [800] * 16 + [1, 2, 3]
v, f = synth.ragged_batch("tri", 20, 20, counts, list(range(len(counts))))
m.keep_taps = True
out = m._run(vertices=v.cuda(), faces=f.cuda(), check_status=True)
taps = m.taps()
res = dict(codes=out["codes"].cpu())
for k, t in taps.items():
    if k.startswith("layer") or k in ("x0", "averaged"):
        res[k] = t.cpu().clone()
torch.save(res, sys.argv[1])
"""


def run(flag, path):
    env = dict(os.environ, MESHTOK_GATHER_FUSED=flag)
    subprocess.run([sys.executable, "-c", SCRIPT.format(root=ROOT, tests=os.path.join(ROOT, "tests")), path], check=True, env=env)


if __name__ == "__main__":
    import torch
    with tempfile.TemporaryDirectory() as d:
        run("0", os.path.join(d, "a.pt"))
        run("1", os.path.join(d, "b.pt"))
        a, b = torch.load(os.path.join(d, "a.pt")), torch.load(os.path.join(d, "b.pt"))
        bad = 0
        for k in a:
            same = torch.equal(a[k], b[k])
            diff = (a[k].double() - b[k].double()).abs().max().item() if a[k].numel() else 0.0
            print(f"{k:16s} shape {tuple(a[k].shape)} identical {same} max|diff| {diff:.3e}")
            bad += not same
        print("FUSED GATHER", "OK: bit-identical" if bad == 0 else f"MISMATCH in {bad} tensors")
        sys.exit(1 if bad else 0)
