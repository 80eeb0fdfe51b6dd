"""C4 workload (256 meshes, 100..800 faces each, default RVQ model) through the ragged layout (tokenize_ragged) next to
the padded layout (tokenize) on the same meshes: device-timed eager calls (L2 flushed between steps) and host-buffer calls (pinned inputs,
H2D + pipeline + D2H of the codes inside the timed region).  One JSON line per layout.

    python scripts/gpu_ragged_bench.py [meshes] [profile]      # profile: one step of each layout between cudaProfilerStart / Stop
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from helpers import make_pair
from meshgpt_pytorch_b200 import data as Dt
import synth
from meshgpt_pytorch_b200._lib import lib

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
o, m = make_pair(use_residual_lfq=False)
g = torch.Generator().manual_seed(1234)
counts = torch.randint(100, 801, (B,), generator=g).tolist()
v, f = synth.ragged_batch("tri", 20, 20, counts, range(B))
r = Dt.padded_to_ragged(v, f)
n_faces = sum(counts)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
vp, fp = v.pin_memory(), f.pin_memory()
rvp, rfp = r["vertices"].pin_memory(), r["faces"].pin_memory()
fo, vo = r["face_offsets"], r["vertex_offsets"]
vd, fd, rvd, rfd = vp.cuda(), fp.cuda(), rvp.cuda(), rfp.cuda()
fod, vod = fo.cuda(), vo.cuda()
F, V = max(counts), int((vo[1:] - vo[:-1]).max())
out_p = torch.empty((B, f.shape[1] * 3, 2), dtype=torch.int64).pin_memory()
out_r = torch.empty((n_faces * 3, 2), dtype=torch.int64).pin_memory()


def padded_dev():
    return m.forward(vertices=vd, faces=fd, return_codes=True, check_status=False)


def ragged_dev():
    return m.tokenize_ragged(rvd, rfd, fod, vod, max_faces=F, max_vertices=V, check_status=False)


def padded_host():
    c = m.forward(vertices=vp.cuda(non_blocking=True), faces=fp.cuda(non_blocking=True), return_codes=True, check_status=False)
    out_p.copy_(c, non_blocking=True)
    torch.cuda.current_stream().synchronize()


def ragged_host():
    c = m.tokenize_ragged(rvp.cuda(non_blocking=True), rfp.cuda(non_blocking=True), fo, vo, check_status=False)
    out_r.copy_(c, non_blocking=True)
    torch.cuda.current_stream().synchronize()


def dev_ms(fn, steps=20):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    evs = []
    for _ in range(steps):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); evs.append((a, b))
    torch.cuda.synchronize()
    return sorted(a.elapsed_time(b) for a, b in evs)[steps // 2]


def host_ms(fn, steps=20):
    for _ in range(5):
        fn()
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter(); fn(); ts.append((time.perf_counter() - t0) * 1e3)
    return sorted(ts)[steps // 2]


assert torch.equal(Dt.ragged_to_padded_codes(ragged_dev(), fo, 3), padded_dev())
m.check_last_status()
if len(sys.argv) > 2 and sys.argv[2] == "profile":     # ncu --profile-from-start off: one padded step, then one ragged step
    for _ in range(3):
        padded_dev(); ragged_dev()
    torch.cuda.synchronize()
    torch.cuda.profiler.start(); padded_dev(); torch.cuda.synchronize(); ragged_dev(); torch.cuda.synchronize(); torch.cuda.profiler.stop()
    sys.exit(0)
# the same two layouts as CUDA graphs: the padded graph serves this (b, nf, nv) shape only, the ragged one every batch within its bounds
gp = m.graphed(vd, fd)
gr = m.graphed_ragged(B, 800, 441, B * 800, B * 441)
assert torch.equal(gr(rvd, rfd, fo, vo), ragged_dev()) and torch.equal(gp(), padded_dev())


def padded_graph_host():
    gp.tokenize_host(vp, fp, out_p)


def ragged_graph_host():
    out_r.copy_(gr(rvp, rfp, fo, vo), non_blocking=True)
    torch.cuda.current_stream().synchronize()


h2d_p, d2h_p = vp.numel() * 4 + fp.numel() * 8, out_p.numel() * 8
h2d_r, d2h_r = rvp.numel() * 4 + rfp.numel() * 8 + 8 * (B + 1), out_r.numel() * 8
for name, dfn, hfn, h2d, d2h in (("padded", padded_dev, padded_host, h2d_p, d2h_p), ("ragged", ragged_dev, ragged_host, h2d_r, d2h_r),
                                 ("padded, graph", lambda: gp(), padded_graph_host, h2d_p, d2h_p),
                                 ("ragged, graph", lambda: gr.graph.replay(), ragged_graph_host, h2d_r, d2h_r)):
    l0 = lib.meshtok_kernel_launch_count(); dfn(); launches = (lib.meshtok_kernel_launch_count() - l0) or 26     # (a replay counts no launches)
    d, h = dev_ms(dfn), host_ms(hfn)
    print(json.dumps(dict(layout=name, metric="faces tokenized/sec (valid faces; median of 20 eager steps)", meshes=B, valid_faces=n_faces,
                          slots=B * f.shape[1], value=n_faces / (d * 1e-3), unit="faces/s", ms_per_step=d,
                          e2e=dict(value=n_faces / (h * 1e-3), unit="faces/s", ms_per_step=h, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h),
                          counted_launches_per_step=launches)))
