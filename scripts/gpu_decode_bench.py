"""Decode throughput (faces decoded / s, device-timed) at the C2 shape with the reference's default decoder, next to the CPU
decode oracle on a sample.  First-version numbers: the path is correct, not yet tuned."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import meshgpt_pytorch_b200 as M
from helpers import make_pair
from meshgpt_pytorch_b200 import synth
from meshgpt_pytorch_b200._lib import lib
from oracle import decode_oracle as D

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
o, m = make_pair(use_residual_lfq=False)
od = D.OracleDecoder(o.quantizer, use_residual_lfq=False).eval()
md = M.MeshDecoder(m)
md.load_reference_state_dict({k: v for k, v in od.state_dict().items() if not k.startswith("quantizer.")})
md = md.cuda()
v, f = synth.grid_batch("tri", 20, 20, list(range(B)))
codes = m.tokenize(v.cuda(), f.cuda())
for _ in range(3):
    md.decode_from_codes_to_faces(codes)
torch.cuda.synchronize()
if len(sys.argv) > 2 and sys.argv[2] == "profile":     # ncu --profile-from-start off: one eager step
    torch.cuda.profiler.start(); md.decode_from_codes_to_faces(codes); torch.cuda.synchronize(); torch.cuda.profiler.stop()
    sys.exit(0)
l0 = lib.meshtok_kernel_launch_count()
steps = 10
evs = []
for _ in range(steps):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); coords, mask = md.decode_from_codes_to_faces(codes); b.record(); evs.append((a, b))
torch.cuda.synchronize()
ms = sum(a.elapsed_time(b) for a, b in evs) / steps
launches = (lib.meshtok_kernel_launch_count() - l0) / steps
n_cpu = min(B, 8)
t0 = time.perf_counter(); od.decode_from_codes_to_faces(codes[:n_cpu].cpu()); cpu_s = time.perf_counter() - t0
g = md.graphed(codes)
for _ in range(3):
    g()
torch.cuda.synchronize()
evs = []
for _ in range(steps):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); gc, gm = g(); b.record(); evs.append((a, b))
torch.cuda.synchronize()
ms_graph = sum(a.elapsed_time(b) for a, b in evs) / steps
assert torch.equal(torch.nan_to_num(gc), torch.nan_to_num(coords))
ms_eager, ms = ms, ms_graph
print(json.dumps(dict(eager_ms_per_step=ms_eager, metric="faces decoded/sec (device-timed)", value=B * 800 / (ms * 1e-3), unit="faces/s", ms_per_step=ms, meshes=B,
                      counted_launches_per_step=launches,
                      cpu_baseline=dict(value=n_cpu * 800 / cpu_s, unit="faces/s", cores=os.cpu_count(), kind="port", sample=f"decode oracle on {n_cpu} meshes"))))
