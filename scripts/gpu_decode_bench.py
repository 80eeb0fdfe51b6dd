"""Decode throughput (faces decoded / s, device-timed) at the C2 shape with the reference's default decoder, next to the CPU
decode oracle on a sample.  Prints one JSON line for the fused schedule (graph replay = the number to quote, eager beside it) and
one for the stage-by-stage schedule (MESHTOK_DEC_FUSED=0).

    python scripts/gpu_decode_bench.py [meshes] [profile | fusedonly]      # profile: one eager step between cudaProfilerStart / Stop
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import meshgpt_pytorch_b200 as M
from helpers import make_pair
import synth  # tests/synth.py: pure-torch mesh generator (test + bench infrastructure, not product)
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
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")      # > L2 (126 MB), written between timed steps


def timed(fn, steps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    evs = []
    for _ in range(steps):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); out = fn(); b.record(); evs.append((a, b))
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in evs) / steps, out


if len(sys.argv) > 2 and sys.argv[2] == "profile":     # ncu --profile-from-start off: one eager step
    for _ in range(3):
        md.decode_from_codes_to_faces(codes)
    torch.cuda.synchronize()
    torch.cuda.profiler.start(); md.decode_from_codes_to_faces(codes); torch.cuda.synchronize(); torch.cuda.profiler.stop()
    sys.exit(0)

n_cpu = min(B, 8)
t0 = time.perf_counter(); od.decode_from_codes_to_faces(codes[:n_cpu].cpu()); cpu_s = time.perf_counter() - t0
cpu = dict(value=n_cpu * 800 / cpu_s, unit="faces/s", cores=os.cpu_count(), kind="port", sample=f"decode oracle on {n_cpu} meshes")
for schedule in (("fused",) if "fusedonly" in sys.argv else ("fused", "staged")):
    os.environ["MESHTOK_DEC_FUSED"] = "1" if schedule == "fused" else "0"
    md.decode_from_codes_to_faces(codes)          # first call: builds the weight images (one-time launches, not part of a step)
    l0 = lib.meshtok_kernel_launch_count()
    md.decode_from_codes_to_faces(codes)
    launches = lib.meshtok_kernel_launch_count() - l0
    ms_eager, (coords, mask) = timed(lambda: md.decode_from_codes_to_faces(codes))
    g = md.graphed(codes)
    ms_graph, (gc, gm) = timed(lambda: g())
    assert torch.equal(torch.nan_to_num(gc), torch.nan_to_num(coords))
    # roofline of the dominant kernel class, the tcgen05 convolutions (tensor bound): issued bf16 flops = 3 MMAs per split-bf16 product x
    # 2 R N K, over the event-timed duration of every convolution launch of one eager step (L2 flushed before the step)
    from meshgpt_pytorch_b200 import decoder as dec_mod
    flush.fill_(1)
    dec_mod.conv_trace = []
    md.decode_from_codes_to_faces(codes)
    torch.cuda.synchronize()
    trace, dec_mod.conv_trace = dec_mod.conv_trace, None
    conv_ms = sum(a.elapsed_time(b) for *_, a, b in trace)
    conv_flops = sum(6.0 * r * n * k for r, n, k, *_ in trace)
    peak = peaks["bf16_tflops_sustained"]
    roofline = None if not trace else dict(
        kernel=f"linear_tc_pair(_tail)_kernel, conv mode: {len(trace)} launches per step", bound="tensor", achieved=conv_flops / (conv_ms * 1e-3) / 1e12,
        peak=peak, unit="TFLOP/s", frac=conv_flops / (conv_ms * 1e-3) / 1e12 / peak, traffic=None, peak_source="measured (bf16_tflops_sustained)",
        flops_per_step=conv_flops, conv_ms_per_step=conv_ms, conv_share_of_eager_step=conv_ms / ms_eager)
    print(json.dumps(dict(schedule=schedule, roofline=roofline, metric="faces decoded/sec (device-timed, graph replay, L2 flushed between steps)",
                          value=B * 800 / (ms_graph * 1e-3), unit="faces/s", ms_per_step=ms_graph, eager_ms_per_step=ms_eager, meshes=B,
                          counted_launches_per_step=launches, cpu_baseline=cpu)))
