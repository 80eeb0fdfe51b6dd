#!/usr/bin/env python
"""bench.py - faces tokenized / second on B200 (BASELINE.json metric), one JSON line on rank 0.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|c1|c4] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step = one tokenize() of one batch of synthetic meshes per rank (workload c2 = BASELINE configs[1]: 64 meshes x
800 triangle faces, ResidualVQ 16384 x 192, 2 quantizers, random-init weights).  Meshes shard across ranks with no
data-path collective (weak scaling: every rank tokenizes its own 64 meshes per step); the only exchange is the NCCL
all-reduce of the codebook-usage histogram, inside the timed region.

  value       device-timed (CUDA events around each step, inputs resident in HBM, L2 flushed between steps,
              max over ranks) whole-job faces/s.
  e2e         same metric through the host-buffer call MeshAutoencoder.tokenize_host: pinned host inputs ->
              H2D -> pipeline -> D2H of the int64 codes, all inside the timed region.
  roofline    the dominant kernel (tcgen05 RVQ search): 2*R*K*D flops per launch / its CUDA-event duration,
              against MEASURED_PEAKS.json bf16_tflops_sustained (fp16 and bf16 share the kind::f16 tensor pipe).
  cpu_baseline the oracle (CPU restatement of the reference path) on a bounded sample, all host threads.
--impl reference times that oracle alone (the reference package itself cannot be imported: 13 of its
dependencies are not installed and there is no network).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "faces tokenized/sec (device-timed, max over ranks)"
UNIT = "faces/s"

WORKLOADS = {
    "c1": dict(kind="tri", nx=0, ny=0, batch=2, cfg=dict(num_discrete_coors=128), desc="README: batch 2 x 64 random tri faces, 121 vertices, default (LFQ)"),
    "c2": dict(kind="tri", nx=20, ny=20, batch=64, cfg=dict(use_residual_lfq=False), desc="batch 64 x 800 tri faces (20x20 grid, 441 vertices), ResidualVQ 16384x192 x2"),
    "c3": dict(kind="quad", nx=20, ny=40, batch=64, cfg=dict(quads=True), desc="batch 64 x 800 quad faces (20x40 grid, 861 vertices), residual LFQ"),
    "c4": dict(kind="tri", nx=20, ny=20, batch=256, cfg=dict(use_residual_lfq=False), desc="batch 256, faces U[100,800] padded with -1, ResidualVQ"),
    # BASELINE configs[4] (100k meshes sharded over the GPUs) is run as a stream of launches of this size per rank:
    # activations (14 GB) >> L2, the regime the HBM-roofline figures of the gather/scatter/feature kernels are quoted in.
    "c5": dict(kind="tri", nx=20, ny=20, batch=1024, cfg=dict(use_residual_lfq=False), desc="batch 1024 x 800 tri faces per launch (the 100k-mesh sharded run's launch size), ResidualVQ 16384x192 x2"),
}


def make_inputs(name, first_seed, batch):
    from meshgpt_pytorch_b200 import synth
    w = WORKLOADS[name]
    if name == "c1":
        return synth.readme_batch(first_seed, batch=batch)
    if name == "c4":
        g = torch.Generator().manual_seed(1234 + first_seed)
        nf = torch.randint(100, 801, (batch,), generator=g).tolist()
        return synth.ragged_batch(w["kind"], w["nx"], w["ny"], nf, range(first_seed, first_seed + batch))
    return synth.grid_batch(w["kind"], w["nx"], w["ny"], range(first_seed, first_seed + batch))


def valid_faces(faces):
    return int((faces != -1).all(-1).sum())


def build_models(name, want_cuda):
    """Oracle (CPU) with torch.manual_seed(42) init + explicit codebook; CUDA model loaded from its state dict."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers
    cfg = WORKLOADS[name]["cfg"]
    oracle = helpers.make_oracle(seed=42, **cfg)
    model = None
    if want_cuda:
        import meshgpt_pytorch_b200 as M
        model = M.MeshAutoencoder(**cfg)
        model.load_reference_state_dict(oracle.state_dict())
        model = model.cuda().eval()
    return oracle, model


class ClockSampler:
    """SM clock and throttle reasons of one GPU while the bench runs: NVML polled every ~2 ms from a thread (nvidia-smi -lms
    100 as the fallback), so that even a 20 ms timed region gets several samples.  window_begin() / window_end() bracket the
    device-timed loop; the reported median is over the samples inside the window (all samples if none fell inside)."""
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, index):
        self.rows, self.proc, self.nvml, self.stop_flag = [], None, None, False
        self.t0 = self.t1 = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            try:   # the CUDA device may not be NVML device `index` (CUDA_VISIBLE_DEVICES): go by UUID
                uuid = "GPU-" + str(torch.cuda.get_device_properties(index).uuid)
                self.handle = pynvml.nvmlDeviceGetHandleByUUID(uuid)
            except Exception:
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _poll(self):
        n = self.nvml
        while not self.stop_flag:
            try:
                sm = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                try:
                    mask = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                except Exception:
                    mask = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                self.rows.append((time.perf_counter(), sm, self.smax, {name for name, bit in self.REASONS if mask & bit}))
            except Exception:
                pass
            time.sleep(0.002)

    def _pump(self):
        names = [n for n, _ in self.REASONS]
        for line in self.proc.stdout:
            r = [c.strip() for c in line.split(",")]
            try:
                self.rows.append((time.perf_counter(), float(r[0]), float(r[1]),
                                  {n for n, v in zip(names, r[3:7]) if v.lower().startswith("active")}))
            except (ValueError, IndexError):
                continue

    def window_begin(self):
        self.t0 = time.perf_counter()

    def window_end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        if self.nvml is None and self.proc is None:
            return None
        self.stop_flag = True
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        else:
            self.thread.join(timeout=1)
        rows = list(self.rows)
        if not rows:
            return None
        inside = [r for r in rows if self.t0 is not None and self.t1 is not None and self.t0 <= r[0] <= self.t1]
        use = inside or rows
        sm = sorted(r[1] for r in use)
        reasons = set()
        for r in use:
            reasons |= r[3]
        return dict(sm_mhz=sm[len(sm) // 2], sm_min_mhz=sm[0], sm_max_mhz=max(r[2] for r in use), reasons=sorted(reasons),
                    samples=len(use), samples_total=len(rows), source="nvml" if self.nvml is not None else "nvidia-smi",
                    window="device-timed loop" if inside else "whole measurement")


def section_work(model, n_faces, n_vrows, n_edges, batch, F, V):
    """Algorithmic work of each pipeline section per step: bytes each tensor is read/written once
    for the memory-bound kernels, flops for the tensor-core kernels."""
    nvf, D, Q, K = model.num_vertices_per_face, model.dim_codebook, model.num_quantizers, model.codebook_size
    dims_in = [D] + list(model.encoder_dims[:-1])
    dims_out = list(model.encoder_dims)
    N, R, E = n_faces, n_vrows, n_edges
    faces_b = batch * F * nvf * 8
    work = {
        "topology": ("hbm", faces_b + 4 * (E + N + N * nvf + R + 2 * batch * F + batch * V)),
        "face_embed": ("hbm", faces_b + batch * V * 12 + N * D * 4),
        "sage_gather_mean": ("hbm", sum(2 * N * d * 4 for d in dims_in) + len(dims_in) * 4 * (E + N)),
        "sage_row_norm": ("hbm", 3 * N * dims_out[-1] * 4),
        # read y once, write the vertex means once, the vertex incidence CSR; for RVQ models the kernel also emits the level-0
        # residual copy, row scales and the fp16 row image (2 D + 32 B per row) the search kernel reads
        "scatter_mean": ("hbm", N * nvf * D * 4 + R * D * 4 + 4 * (N * nvf + R)
                         + (0 if getattr(model, "use_residual_lfq", True) else R * (D * 4 + 4 + 2 * D + 32))),
        "gather_codes": ("hbm", faces_b + R * Q * 4 + batch * F * nvf * Q * 8),
        "lfq": ("hbm", R * D * 4 + R * Q * 4),
        # fp32-equivalent flops; the split-bf16 scheme issues 3 bf16 MMAs per product
        "sage_linear": ("tensor", sum(2.0 * N * (di * di + 2 * di * do) for di, do in zip(dims_in, dims_out))),
        "pdc_linear": ("tensor", 2.0 * N * dims_out[-1] * nvf * D),
        "rvq_search_tc": ("tensor", 2.0 * R * K * D * Q),
    }
    return work


def cpu_time_oracle(oracle, name, n_meshes, repeats, warmup):
    torch.set_num_threads(os.cpu_count() or 1)
    v, f = make_inputs(name, 0, n_meshes)
    times = []
    for i in range(warmup + repeats):
        t0 = time.perf_counter()
        oracle.tokenize(v, f)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return valid_faces(f), times


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    oracle, _ = build_models(args.workload, want_cuda=False)
    n = min(args.cpu_meshes, WORKLOADS[args.workload].get("batch", args.cpu_meshes))
    faces, times = cpu_time_oracle(oracle, args.workload, n, args.steps, args.warmup)
    ms = 1e3 * sum(times) / len(times)
    val = faces / (ms / 1e3)
    line = dict(impl="reference", metric=METRIC, value=val, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=ms, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
                config=dict(workload=args.workload + ": " + WORKLOADS[args.workload]["desc"],
                            note="CPU oracle (restated reference path); the reference package is not importable here"),
                cpu_baseline=dict(value=val, unit=UNIT, cores=os.cpu_count(), kind="port",
                                  sample=f"{n} of the workload's meshes per step ({faces} faces), mean of {args.steps} steps"),
                e2e=dict(value=val, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-meshes", type=int, default=64, help="meshes in the CPU baseline sample (capped at the batch size)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--pipeline-depth", type=int, default=2, help="batches in flight in the host-to-host pipeline (e2e)")
    ap.add_argument("--no-hist", action="store_true", help="experiment: skip the codebook-usage histogram")
    ap.add_argument("--hist-every-step", action="store_true", help="all-reduce the histogram after every step instead of once after the last")
    ap.add_argument("--precomputed-edges", action="store_true",
                    help="pass face_edges (derived once, outside the timed region) instead of deriving them inside tokenize (BASELINE config 4)")
    ap.add_argument("--no-graph", action="store_true", help="launch the pipeline kernel by kernel instead of replaying its CUDA graph")
    ap.add_argument("--ncu-window", type=int, default=0,
                    help="profiling aid: after the warm-up run this many device steps between cudaProfilerStart/Stop and exit "
                         "(use with ncu --profile-from-start off); prints no bench line")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        return run_reference_arm(args, rank, world)

    import torch.distributed as dist
    assert torch.cuda.is_available(), "bench.py --impl b200 needs a CUDA device (there is no CPU path)"
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from meshgpt_pytorch_b200 import _lib
    lib = _lib.lib
    w = WORKLOADS[args.workload]
    oracle, model = build_models(args.workload, want_cuda=True)
    batch = w["batch"]
    v_cpu, f_cpu = make_inputs(args.workload, rank * batch, batch)
    n_faces_rank = valid_faces(f_cpu)
    v_dev, f_dev = v_cpu.cuda(), f_cpu.cuda()
    e_dev = None
    if args.precomputed_edges:
        import meshgpt_pytorch_b200 as M
        e_dev = M.derive_face_edges_from_faces(f_dev)            # (b, emax, 2) int64, -1 padded: what a dataset cache holds
    v_pin, f_pin = v_cpu.pin_memory(), f_cpu.pin_memory()
    Q, K = model.num_quantizers, model.codebook_size
    hist = torch.zeros(Q, K, dtype=torch.int64, device="cuda")
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")      # > 126 MB L2
    nsec = lib.meshtok_profile_sections()
    sec_names = [lib.meshtok_profile_section_name(i).decode() for i in range(nsec)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    step_hist = torch.zeros_like(hist)

    # The codebook-usage histogram is a statistic of the JOB: every rank accumulates its own on the device (the kernels add
    # into step_hist) and the ranks exchange it ONCE, after the last step, inside the timed region ("reduced once
    # per run (or per batch)").  --hist-every-step all-reduces after every step instead (a cross-rank sync per step).
    def reduce_hist():
        if world > 1 and not args.no_hist:
            dist.all_reduce(step_hist, op=dist.ReduceOp.SUM, async_op=True).wait()

    def eager_step():
        model.forward(vertices=v_dev, faces=f_dev, face_edges=e_dev, return_codes=True, histogram=None if args.no_hist else step_hist, check_status=False)
        if args.hist_every_step:
            reduce_hist()
            hist.add_(step_hist)
            step_hist.zero_()

    # the whole pipeline of a batch as ONE CUDA graph (the C ABI never allocates or synchronises)
    graphed = None
    launches_per_graph = 0
    if not args.no_graph:
        eager_step()                                                            # first call: also builds the weight images
        l0 = lib.meshtok_kernel_launch_count()
        eager_step()                                                            # the same call the graph captures
        launches_per_graph = lib.meshtok_kernel_launch_count() - l0
        graphed = model.graphed(v_dev, f_dev, histogram=None if args.no_hist else step_hist, face_edges=e_dev)

    def device_step():
        if graphed is None:
            return eager_step()
        graphed()
        if args.hist_every_step:
            reduce_hist()
            hist.add_(step_hist)
            step_hist.zero_()

    codes_pin = torch.empty((batch, f_cpu.shape[1] * f_cpu.shape[2], Q), dtype=torch.int64, pin_memory=True)

    def host_step():
        if e_dev is not None:
            return      # the host-buffer entries take vertices and faces only
        if graphed is None:
            model.tokenize_host(v_pin, f_pin, pinned_out=codes_pin)
        else:
            graphed.tokenize_host(v_pin, f_pin, pinned_out=codes_pin)

    def timed(step_fn, steps, profile=False, finalize=None):
        evs = []
        if profile:
            lib.meshtok_profile_enable(1)
        barrier()
        wall0 = time.perf_counter()
        for _ in range(steps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            step_fn()
            b.record()
            evs.append((a, b))
        if finalize is not None:    # the job's one exchange: timed, and charged to the steps
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            finalize()
            b.record()
            evs.append((a, b))
        barrier()
        wall = time.perf_counter() - wall0
        sec_ms = (C.c_float * nsec)()
        sec_n = (C.c_int32 * nsec)()
        if profile:
            lib.meshtok_profile_enable(0)
            _lib.check(lib.meshtok_profile_collect(sec_ms, sec_n, nsec))
        ms = sum(a.elapsed_time(b) for a, b in evs) / steps
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), wall, list(sec_ms), list(sec_n)

    # warm-up (also sizes the workspace and the edge buffer)
    for _ in range(args.warmup):
        device_step()
    model.check_last_status()
    for _ in range(max(1, args.warmup // 2)):
        host_step()
    torch.cuda.synchronize()

    if args.ncu_window:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        for _ in range(args.ncu_window):
            flush.zero_()
            eager_step()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        print(json.dumps(dict(ncu_window_steps=args.ncu_window, workload=args.workload)), flush=True)
        return

    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = lib.meshtok_kernel_launch_count()
    if sampler:
        sampler.window_begin()
    step_hist.zero_()
    ms_dev, wall_dev, _, _ = timed(device_step, args.steps, finalize=None if args.hist_every_step else reduce_hist)
    if sampler:
        sampler.window_end()
    launches = lib.meshtok_kernel_launch_count() - launches0
    if not args.no_hist:   # the job's histogram: every valid token of every step of every rank, once per quantizer level
        got_tokens = int((hist if args.hist_every_step else step_hist).sum().item())
        want_tokens = args.steps * world * n_faces_rank * model.num_vertices_per_face * Q
        assert got_tokens == want_tokens, f"histogram holds {got_tokens} tokens, expected {want_tokens}"
    if graphed is not None:
        launches = launches_per_graph * args.steps          # kernels inside the replayed graphs
    # second pass, kernel by kernel, with cudaEvent pairs around every pipeline section (kept out of the headline timing)
    prof_steps = min(args.steps, 10)
    for _ in range(2):
        eager_step()
    _, _, sec_ms, sec_n = timed(eager_step, prof_steps, profile=True)
    model.check_last_status()
    ms_e2e_serial, _, _, _ = timed(host_step, args.steps)
    ms_e2e = ms_e2e_serial
    if graphed is not None and e_dev is None:
        # the public host-to-host API for a stream of batches: HostPipeline double-buffers, so the H2D copy of step i+1 and
        # the D2H copy of step i-1 overlap the kernels of step i; every step's copies are inside the timed region
        pipe = model.host_pipeline(v_dev, f_dev, depth=args.pipeline_depth)
        for _ in range(2 * args.pipeline_depth):
            pipe.submit(v_pin, f_pin)
        pipe.drain()
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for slot in pipe.slots:
            slot["stream"].wait_event(ev0)
        for _ in range(args.steps):
            pipe.submit(v_pin, f_pin)
        last = pipe.drain()
        assert torch.equal(last[-1][1], codes_pin), "pipelined and single-call host paths disagree"
        for slot in pipe.slots:
            torch.cuda.current_stream().wait_stream(slot["stream"])
        ev1.record()
        ev1.synchronize()
        t = torch.tensor([ev0.elapsed_time(ev1) / args.steps], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t.item())
    clocks = sampler.stop() if sampler else None

    total_faces = torch.tensor([n_faces_rank], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(total_faces)
    total_faces = int(total_faces.item())

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        t = model.taps()
        R = t["n_vrows"]
        D = model.dim_codebook
        sec = {n: dict(ms_per_step=sec_ms[i] / prof_steps, launches_per_step=sec_n[i] / prof_steps) for i, n in enumerate(sec_names) if sec_n[i]}
        hbm_peak = peaks.get("hbm_gbs") or 6650.0
        tc_peak = peaks.get("bf16_tflops_sustained") or 1400.0
        work = section_work(model, t["n_faces"], R, t["n_edges"], batch, f_cpu.shape[1], v_cpu.shape[1])
        for n, d in sec.items():
            if n not in work or d["ms_per_step"] <= 0:
                continue
            bound, amount = work[n]
            if bound == "hbm":
                gbs = amount / (d["ms_per_step"] * 1e-3) / 1e9
                d.update(bound="hbm", algorithmic_bytes=int(amount), achieved_gbs=gbs, frac_of_hbm_peak=gbs / hbm_peak)
            else:
                tf = amount / (d["ms_per_step"] * 1e-3) / 1e12
                mult = 1.0 if n == "rvq_search_tc" else 3.0
                d.update(bound="tensor", flops=amount, achieved_tflops=tf, tensor_pipe_tflops=tf * mult, frac_of_bf16_peak=tf * mult / tc_peak)
        roofline = None
        if not model.use_residual_lfq and "rvq_search_tc" in sec:
            per_launch_ms = sec_ms[sec_names.index("rvq_search_tc")] / max(1, sec_n[sec_names.index("rvq_search_tc")])
            flops = 2.0 * R * K * D
            achieved = flops / (per_launch_ms * 1e-3) / 1e12
            peak = peaks.get("bf16_tflops_sustained") or 1400.0
            # DRAM bytes per launch of this kernel from the committed ncu --set full capture (profiles/traffic.json), or null
            traffic, kernel_name = None, "rvq_tc2_search_kernel (CTA pair) / rvq_tc_search_kernel"
            try:
                tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
                traffic = tj["rvq_search"]["dram_bytes_per_launch"]
                kernel_name = tj["rvq_search"].get("kernel", kernel_name)
            except Exception:
                pass
            roofline = dict(kernel=kernel_name, bound="tensor", achieved=achieved, peak=peak, unit="TFLOP/s",
                            frac=achieved / peak, traffic=traffic, peak_source="measured" if peaks else "fallback",
                            flops_per_launch=flops, ms_per_launch=per_launch_ms)
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            n_cpu = min(args.cpu_meshes, batch)
            faces_s, times = cpu_time_oracle(oracle, args.workload, n_cpu, repeats=5, warmup=1)
            cpu = dict(value=faces_s / min(times), unit=UNIT, cores=os.cpu_count(), kind="port",
                       sample=f"oracle on {n_cpu} of the workload's {batch} meshes ({faces_s} faces), best of 5 after 1 warm-up, "
                              f"{sum(times):.1f} s of CPU work")
        h2d = v_pin.numel() * v_pin.element_size() + f_pin.numel() * f_pin.element_size()
        d2h = codes_pin.numel() * codes_pin.element_size()
        line = dict(metric=METRIC, value=total_faces / (ms_dev * 1e-3), unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=ms_dev, higher_is_better=True, scaling="weak", vs_baseline=None,
                    dtype="f32 activations (split-bf16 tcgen05 linears), f16 tcgen05 RVQ screening + f32 re-score",
                    data="synthetic",
                    config=dict(workload=args.workload + ": " + w["desc"], meshes_per_rank=batch, faces_per_step=total_faces,
                                quantizer_rows_per_rank=R, l2="flushed between steps (512 MiB memset outside the timed events)",
                                sharding="meshes by rank, weights replicated; (Q,K) int64 histogram accumulated per rank on the device, " + ("NCCL all-reduce after every step" if args.hist_every_step else "ONE NCCL all-reduce after the last step, inside the timed region"),
                                launch="one CUDA graph replay per step" if graphed is not None else "kernel by kernel",
                                face_edges="precomputed (user-supplied edge lists -> CSR by target on the device)" if e_dev is not None else "derived inside tokenize",
                                wall_ms_per_step_incl_flush=1e3 * wall_dev / args.steps,
                                rvq_rows_rechecked_by_exact_scan=t.get("rvq_exact_rechecks")),
                    e2e=None if e_dev is not None else dict(value=total_faces / (ms_e2e * 1e-3), unit=UNIT, h2d_bytes_per_step=h2d * world, d2h_bytes_per_step=d2h * world,
                             ms_per_step=ms_e2e, api=f"HostPipeline.submit ({args.pipeline_depth} batches in flight, graph replays)" if graphed is not None else "tokenize_host",
                             single_call_ms=ms_e2e_serial, single_call_value=total_faces / (ms_e2e_serial * 1e-3)),
                    gpu_launches=int(launches), gpu_launches_per_step=launches / args.steps, sections=sec, roofline=roofline, cpu_baseline=cpu, clocks=clocks)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
