"""Front-end throughput: codes -> fine tokens + face tokens at the C2 shape (64 x 800 triangle faces, 2 levels,
16384 codes + eos, dim 512), device-timed with L2 flushed between steps, against the HBM roofline, next to the CPU oracle.

    python scripts/gpu_frontend_bench.py [meshes] [profile]
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import meshgpt_pytorch_b200 as M
from oracle import frontend_oracle as FO

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
NF, N, D, K = 800, 6, 512, 16384
L = NF * N
torch.manual_seed(0)
o = FO.OracleTokenFrontEnd(codebook_size=K, num_quantizers=2, quads=False, dim=D, max_seq_len=L).eval()
m = M.MeshTransformerFrontEnd(codebook_size=K, num_quantizers=2, quads=False, dim=D, max_seq_len=L)
m.load_reference_state_dict(o.state_dict())
m = m.cuda()
codes = torch.randint(0, K, (B, L), device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")      # > L2 (126 MB), written between timed steps
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
try:
    traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))     # DRAM bytes per launch from the committed ncu capture
except OSError:
    traffic = {}


def timed(fn, steps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    evs = []
    for _ in range(steps):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); evs.append((a, b))
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in evs) / steps


if len(sys.argv) > 2 and sys.argv[2] == "profile":
    for _ in range(3):
        m.embed_codes(codes)
    torch.cuda.synchronize()
    torch.cuda.profiler.start(); m.embed_codes(codes); torch.cuda.synchronize(); torch.cuda.profiler.stop()
    sys.exit(0)

ms_fine = timed(lambda: m.fine_tokens(codes))
ms_face = timed(lambda: m.face_tokens(codes))
ms_both = timed(lambda: m.embed_codes(codes))
faces = B * NF
# algorithmic bytes per face: fine = n codes (8 B) + n rows written (4 D) [the embedding rows are L2-resident tables shared by the batch];
# face = n codes + n gathered table rows (4 D each, the 201 MB of folded tables do not fit L2) + 1 row written
bytes_fine = faces * (N * 8 + N * 4 * D)
bytes_face = faces * (N * 8 + N * 4 * D + 4 * D)
hbm = peaks.get("hbm_gbs_burst") or peaks.get("hbm_gbs") or 6554.6
n_cpu = min(B, 16)
t0 = time.perf_counter(); o(codes[:n_cpu].cpu()); cpu_s = time.perf_counter() - t0
print(json.dumps(dict(
    metric="faces embedded/sec (fine tokens + face tokens, device-timed, L2 flushed between steps)", value=faces / (ms_both * 1e-3), unit="faces/s",
    ms_per_step=ms_both, meshes=B, gpu_launches_per_step=2,
    roofline=dict(kernel="face_tokens_kernel<4>", bound="hbm", achieved=bytes_face / (ms_face * 1e-3) / 1e9, peak=hbm, unit="GB/s",
                  frac=bytes_face / (ms_face * 1e-3) / 1e9 / hbm,
                  traffic=traffic.get("frontend_face_tokens", {}).get("dram_bytes_per_launch") if B == 64 else None, ms_per_launch=ms_face, bytes_per_launch=bytes_face),
    fine_tokens=dict(ms_per_launch=ms_fine, bytes_per_launch=bytes_fine, achieved_gbs=bytes_fine / (ms_fine * 1e-3) / 1e9,
                     frac_of_hbm_peak=bytes_fine / (ms_fine * 1e-3) / 1e9 / hbm),
    cpu_baseline=dict(value=n_cpu * NF / cpu_s, unit="faces/s", cores=os.cpu_count(), kind="port", sample=f"front-end oracle on {n_cpu} meshes"))))
