#!/usr/bin/env python
"""bench.py - faces tokenized / second on B200 (BASELINE.json metric), one JSON line on rank 0.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|c1|c4|c5] [--impl b200|reference] [--workloads c1,c3,c4,c4pe,c5|none]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Headline (same workload at every N, so that the driver's scaling efficiency compares like with like): a step = one tokenize() of
one batch of synthetic meshes per rank, workload c2 = BASELINE configs[1]: 64 meshes x 800 triangle faces, ResidualVQ 16384 x 192,
2 quantizers, random-init weights; every step takes a DIFFERENT resident batch (4 rotating batches per rank, seeds = global mesh
index).  Meshes shard across ranks with no data-path collective (weak scaling); the only exchange is the NCCL all-reduce of the
codebook-usage histogram, inside the timed region.

  value        device-timed (CUDA events around each step, inputs resident in HBM, L2 flushed between steps, max over ranks)
               whole-job faces/s.
  e2e          the same metric through the host-buffer API (HostPipeline = libmeshtok's native meshtok_pipeline_*): pinned host
               inputs -> H2D -> pipeline -> D2H of the int64 codes and the status word, every step inside the timed region.
  roofline     the dominant kernel (tcgen05 RVQ search): 2*R*K*D flops per launch / its CUDA-event duration, against
               MEASURED_PEAKS.json bf16_tflops (the BURST figure: the kernel runs for ~130 us inside a sub-second window; the
               sustained figure is reported next to it); fp16 and bf16 share the kind::f16 tensor pipe.
  cpu_baseline the oracle (CPU restatement of the reference path) on a bounded sample, all host threads.
  workloads    the other BASELINE configs, a few steps each, outside the headline timing: c1 (README call), c3 (quads + LFQ),
               c4 (256 ragged meshes; derived and precomputed edges), and c5 = configs[4], the sharded 100k-mesh run's regime:
               1 024 meshes per launch per rank generated ON THE DEVICE by global mesh index (weak, >= 1 s timed, histogram
               all-reduced per batch on a side stream), a fixed 8 192-mesh collection (strong scaling) and a ragged 4 096-mesh
               collection cut by balanced_contiguous_shards (Sigma-faces balance).
--impl reference times the reference's CPU implementation of the path: the real package if baseline/_ref holds an importable
copy (probed at run time and reported), else the oracle (13 of the reference's dependencies are not installed).
The reference arm never imports the product package (no libmeshtok.so in its process).
"""
import argparse
import ctypes as C
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

METRIC = "faces tokenized/sec (device-timed, max over ranks)"
UNIT = "faces/s"
ROTATE = 4           # distinct resident batches per rank the steps cycle through

WORKLOADS = {
    "c1": dict(kind="tri", nx=0, ny=0, batch=2, cfg=dict(num_discrete_coors=128), desc="README: batch 2 x 64 random tri faces, 121 vertices, default (LFQ)"),
    "c2": dict(kind="tri", nx=20, ny=20, batch=64, cfg=dict(use_residual_lfq=False), desc="batch 64 x 800 tri faces (20x20 grid, 441 vertices), ResidualVQ 16384x192 x2"),
    "c3": dict(kind="quad", nx=20, ny=40, batch=64, cfg=dict(quads=True), desc="batch 64 x 800 quad faces (20x40 grid, 861 vertices), residual LFQ"),
    "c4": dict(kind="tri", nx=20, ny=20, batch=256, cfg=dict(use_residual_lfq=False), desc="batch 256, faces U[100,800] padded with -1, ResidualVQ"),
    # BASELINE configs[4] (100k meshes sharded over the GPUs) is run as a stream of launches of this size per rank:
    # activations (14 GB) >> L2, the regime the HBM-roofline figures of the gather/scatter/feature kernels are quoted in.
    "c5": dict(kind="tri", nx=20, ny=20, batch=1024, cfg=dict(use_residual_lfq=False), desc="batch 1024 x 800 tri faces per launch (the 100k-mesh sharded run's launch size), ResidualVQ 16384x192 x2"),
}
L2_NOTE = "GPU arm: flushed between steps (512 MiB memset outside the timed events); CPU arm: not applicable"


def make_inputs(name, first_seed, batch):
    import synth
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


def workload_config(name, world, precomputed_edges=False):
    """The workload-defining part of the JSON line - the SAME dict in the b200 and the reference arm."""
    w = WORKLOADS[name]
    _, f = make_inputs(name, 0, w["batch"])
    return dict(workload=name + ": " + w["desc"], meshes_per_rank=w["batch"], faces_per_step=valid_faces(f) * world,
                face_edges="precomputed (user-supplied edge lists)" if precomputed_edges else "derived inside tokenize",
                inputs=f"{ROTATE} distinct batches per rank in rotation, seeds = global mesh index", l2=L2_NOTE)


def build_oracle(name):
    import helpers
    return helpers.make_oracle(seed=42, **WORKLOADS[name]["cfg"])


def build_model(name, oracle):
    import meshgpt_pytorch_b200 as M
    model = M.MeshAutoencoder(**WORKLOADS[name]["cfg"])
    model.load_reference_state_dict(oracle.state_dict())
    return model.cuda().eval()


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


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


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
        # per level: candidate entries (6 x 32 B), the residual row read and written, the winning codebook row, the fp16 row image
        "rvq_rescore": ("hbm", Q * R * (192 + 3 * D * 4 + 2 * D)),
        # fp32-equivalent flops; the split-bf16 scheme issues 3 bf16 MMAs per product
        "sage_linear": ("tensor", sum(2.0 * N * (di * di + 2 * di * do) for di, do in zip(dims_in, dims_out))),
        "pdc_linear": ("tensor", 2.0 * N * dims_out[-1] * nvf * D),
        "rvq_search_tc": ("tensor", 2.0 * R * K * D * Q),
    }
    return work


def annotate_sections(model, sec_ms, sec_n, sec_names, steps, taps, batch, F, V, peaks):
    sec = {n: dict(ms_per_step=sec_ms[i] / steps, launches_per_step=sec_n[i] / steps) for i, n in enumerate(sec_names) if sec_n[i]}
    hbm_peak = peaks.get("hbm_gbs") or 6650.0
    tc_burst = peaks.get("bf16_tflops") or 1600.0
    tc_sust = peaks.get("bf16_tflops_sustained") or 1400.0
    work = section_work(model, taps["n_faces"], taps["n_vrows"], taps["n_edges"], batch, F, V)
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
            d.update(bound="tensor", flops=amount, achieved_tflops=tf, tensor_pipe_tflops=tf * mult, frac_of_bf16_peak=tf * mult / tc_burst,
                     frac_of_bf16_sustained=tf * mult / tc_sust)
    return sec


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


def probe_reference_package():
    """is there an importable copy of the reference under baseline/_ref on THIS machine?  Probed in a
    subprocess (a failing import must not poison this one)."""
    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(ref_dir) or not os.listdir(ref_dir):
        return dict(found=False, importable=False, why="baseline/_ref is absent or empty (pip install of the reference fails offline: setup_requires pytest-runner + 13 missing dependencies)")
    code = f"import sys; sys.path.insert(0, {ref_dir!r}); import meshgpt_pytorch; print(meshgpt_pytorch.__file__)"
    try:
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    except Exception as e:   # noqa: BLE001
        return dict(found=True, importable=False, why=f"probe failed: {e}")
    if r.returncode == 0:
        return dict(found=True, importable=True, why=r.stdout.strip())
    return dict(found=True, importable=False, why=(r.stderr.strip().splitlines() or ["import failed"])[-1][:200])


def time_reference_package(name, oracle, n_meshes, repeats, warmup):
    """The UNMODIFIED reference package from baseline/_ref, random-init weights of the oracle (same state-dict keys), CPU."""
    sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
    from meshgpt_pytorch import MeshAutoencoder   # noqa: E402
    torch.set_num_threads(os.cpu_count() or 1)
    ref = MeshAutoencoder(**WORKLOADS[name]["cfg"]).eval()
    missing = ref.load_state_dict(oracle.state_dict(), strict=False)
    v, f = make_inputs(name, 0, n_meshes)
    times = []
    for i in range(warmup + repeats):
        t0 = time.perf_counter()
        ref.tokenize(vertices=v, faces=f)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return valid_faces(f), times, f"reference package, {len(missing.missing_keys)} keys not in the oracle's state dict (decoder half)"


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    oracle = build_oracle(args.workload)
    n = min(args.cpu_meshes, WORKLOADS[args.workload]["batch"])
    probe = probe_reference_package()
    kind, how = "port", "CPU oracle (restated reference path); the reference package is not importable here"
    faces = times = None
    if probe["importable"]:
        try:
            faces, times, how = time_reference_package(args.workload, oracle, n, args.steps, args.warmup)
            kind = "reference"
        except Exception as e:   # noqa: BLE001
            probe["why"] = f"import worked, tokenize failed: {e}"[:200]
    if times is None:
        faces, times = cpu_time_oracle(oracle, args.workload, n, args.steps, args.warmup)
    ms = 1e3 * sum(times) / len(times)
    val = faces / (ms / 1e3)
    line = dict(impl="reference", metric=METRIC, value=val, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=ms, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
                config=workload_config(args.workload, world, args.precomputed_edges),
                cpu_baseline=dict(value=val, unit=UNIT, cores=os.cpu_count(), kind=kind,
                                  sample=f"{n} of the workload's meshes per step ({faces} faces), mean of {args.steps} steps; {how}"),
                e2e=dict(value=val, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0,
                reference_package=probe)
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------------------
# the b200 arm
# ---------------------------------------------------------------------------------------------------------------------------


class Ctx:
    """Everything the measurements share: the library handle, the L2 flush buffer, the process group."""

    def __init__(self, args, rank, world, local):
        import torch.distributed as dist
        from meshgpt_pytorch_b200 import _lib
        self.args, self.rank, self.world, self.local, self.dist, self._lib = args, rank, world, local, dist, _lib
        self.lib = _lib.lib
        self.flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")      # > 126 MB L2
        self.nsec = self.lib.meshtok_profile_sections()
        self.sec_names = [self.lib.meshtok_profile_section_name(i).decode() for i in range(self.nsec)]
        self.peaks = load_peaks()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, x):
        t = torch.tensor([x], dtype=torch.int64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t)
        return int(t.item())

    def timed(self, step_fn, steps, profile=False, finalize=None, pre=None):
        """K steps, each bracketed by CUDA events on the current stream; the L2 flush and `pre(i)` (making step i's inputs resident)
        happen before the start event.  Returns (mean ms per step, max over ranks), wall seconds, section ms, section launches."""
        evs = []
        if profile:
            self.lib.meshtok_profile_enable(1)
        self.barrier()
        wall0 = time.perf_counter()
        for i in range(steps):
            if pre is not None:
                pre(i)
            self.flush.zero_()
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
        self.barrier()
        wall = time.perf_counter() - wall0
        sec_ms = (C.c_float * self.nsec)()
        sec_n = (C.c_int32 * self.nsec)()
        if profile:
            self.lib.meshtok_profile_enable(0)
            self._lib.check(self.lib.meshtok_profile_collect(sec_ms, sec_n, self.nsec))
        ms = sum(a.elapsed_time(b) for a, b in evs) / steps
        return self.max_over_ranks(ms), wall, list(sec_ms), list(sec_n)


def measure_workload(ctx, name, oracle, model, steps, warmup, precomputed_edges=False, want_e2e=True, want_cpu=True, cpu_meshes=16,
                     hist_every_step=False, no_hist=False, no_graph=False, pipeline_depth=2, headline=False, sampler=None):
    """One BASELINE workload on this rank's GPU: device-timed graph replays over ROTATE distinct resident batches, the section
    pass, the host-to-host pipeline and (rank 0, N = 1) the CPU baseline.  Returns the dict that goes into the JSON line."""
    import meshgpt_pytorch_b200 as M
    args, rank, world, lib = ctx.args, ctx.rank, ctx.world, ctx.lib
    w = WORKLOADS[name]
    batch = w["batch"]
    rot = ROTATE if name != "c1" else 1          # (c1 is the README call on its one pair of tensors)
    host = [make_inputs(name, (k * world + rank) * batch, batch) for k in range(rot)]
    F = max(f.shape[1] for _, f in host)
    V = max(v.shape[1] for v, _ in host)

    def pad_to(v, f):       # ragged workloads: every rotating batch in ONE padded shape (one graph)
        if v.shape[1] == V and f.shape[1] == F:
            return v, f
        v2 = torch.full((v.shape[0], V, 3), -1.0)
        v2[:, : v.shape[1]] = v
        f2 = torch.full((f.shape[0], F, f.shape[2]), -1, dtype=f.dtype)
        f2[:, : f.shape[1]] = f
        return v2, f2
    host = [pad_to(v, f) for v, f in host]
    faces_per_batch = [valid_faces(f) for _, f in host]
    dev = [(v.cuda(), f.cuda()) for v, f in host]
    edges = None
    if precomputed_edges:
        el = [M.derive_face_edges_from_faces(f) for _, f in dev]          # what a dataset cache holds: (b, emax, 2) int64, -1 padded
        E = max(e.shape[1] for e in el)
        edges = [torch.nn.functional.pad(e, (0, 0, 0, E - e.shape[1]), value=-1) for e in el]
    Q, K = model.num_quantizers, model.codebook_size
    hist = torch.zeros(Q, K, dtype=torch.int64, device="cuda")
    step_hist = torch.zeros_like(hist)
    use_hist = not no_hist

    def reduce_hist():
        if world > 1 and use_hist:
            ctx.dist.all_reduce(step_hist, op=ctx.dist.ReduceOp.SUM, async_op=True).wait()

    cur = [0]

    def eager_step():
        v, f = dev[cur[0] % rot]
        e = edges[cur[0] % rot] if edges is not None else None
        model.forward(vertices=v, faces=f, face_edges=e, return_codes=True, histogram=step_hist if use_hist else None, check_status=False)
        if hist_every_step:
            reduce_hist()
            hist.add_(step_hist)
            step_hist.zero_()

    graphed, launches_per_graph = None, 0
    if not no_graph:
        eager_step()                                                            # first call: also builds the weight images
        l0 = lib.meshtok_kernel_launch_count()
        eager_step()                                                            # the same call the graph captures
        launches_per_graph = lib.meshtok_kernel_launch_count() - l0
        graphed = model.graphed(dev[0][0], dev[0][1], histogram=step_hist if use_hist else None, face_edges=None if edges is None else edges[0])

    def make_resident(i):        # outside the timed events: step i's batch -> the graph's static inputs (device-to-device)
        cur[0] = i
        if graphed is not None:
            v, f = dev[i % rot]
            graphed.vertices.copy_(v)
            graphed.faces.copy_(f)
            if edges is not None:
                graphed.face_edges.copy_(edges[i % rot])

    def device_step():
        if graphed is None:
            return eager_step()
        graphed.graph.replay()
        if hist_every_step:
            reduce_hist()
            hist.add_(step_hist)
            step_hist.zero_()

    for i in range(warmup):
        make_resident(i)
        device_step()
    model.check_last_status()
    torch.cuda.synchronize()
    if headline and args.ncu_window:
        torch.cuda.profiler.start()
        for i in range(args.ncu_window):
            cur[0] = i
            ctx.flush.zero_()
            eager_step()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return dict(ncu_window_steps=args.ncu_window, workload=name)

    launches0 = lib.meshtok_kernel_launch_count()
    if sampler:
        sampler.window_begin()
    step_hist.zero_()
    hist.zero_()
    ms_dev, wall_dev, _, _ = ctx.timed(device_step, steps, finalize=None if hist_every_step else reduce_hist, pre=make_resident)
    if sampler:
        sampler.window_end()
    launches = lib.meshtok_kernel_launch_count() - launches0
    faces_timed = sum(faces_per_batch[i % rot] for i in range(steps))
    if use_hist:   # the job's histogram: every valid token of every step of every rank, once per quantizer level
        got_tokens = int((hist if hist_every_step else step_hist).sum().item())
        want_tokens = ctx.sum_over_ranks(faces_timed) * model.num_vertices_per_face * Q if (world > 1) else faces_timed * model.num_vertices_per_face * Q
        assert got_tokens == want_tokens, f"{name}: histogram holds {got_tokens} tokens, expected {want_tokens}"
    if graphed is not None:
        launches = launches_per_graph * steps          # kernels inside the replayed graphs
    total_faces_step = ctx.sum_over_ranks(faces_timed) / steps

    # second pass, kernel by kernel, with cudaEvent pairs around every pipeline section (kept out of the headline timing)
    prof_steps = min(steps, 10)
    for i in range(2):
        cur[0] = i
        eager_step()
    _, _, sec_ms, sec_n = ctx.timed(eager_step, prof_steps, profile=True, pre=lambda i: cur.__setitem__(0, i))
    model.check_last_status()
    taps = model.taps()
    sections = annotate_sections(model, sec_ms, sec_n, ctx.sec_names, prof_steps, taps, batch, F, V, ctx.peaks)

    out = dict(value=total_faces_step / (ms_dev * 1e-3), unit=UNIT, ms_per_step=ms_dev, meshes_per_rank=batch, faces_per_step=total_faces_step,
               quantizer_rows_per_rank=taps["n_vrows"], gpu_launches_per_step=launches / steps, sections=sections,
               rvq_rows_rechecked_by_exact_scan=taps.get("rvq_exact_rechecks"),
               face_edges="precomputed (user-supplied edge lists -> CSR by target on the device)" if edges is not None else "derived inside tokenize",
               launch="one CUDA graph replay per step" if graphed is not None else "kernel by kernel",
               wall_ms_per_step_incl_flush=1e3 * wall_dev / steps)
    out["_launches"] = int(launches)
    out["_taps"] = dict(R=taps["n_vrows"])

    # host to host: pinned inputs -> H2D -> pipeline -> D2H codes + status, through libmeshtok's native batch pipeline
    if want_e2e and edges is None:
        pins = [(v.pin_memory(), f.pin_memory()) for v, f in host]
        pipe = model.host_pipeline(dev[0][0], dev[0][1], depth=pipeline_depth)
        for i in range(2 * pipeline_depth):
            pipe.submit(*pins[i % rot])
        pipe.drain()
        ctx.barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        streams = pipe.streams()
        ev0.record()
        for s in streams:
            s.wait_event(ev0)
        t0 = time.perf_counter()
        n_done = 0
        for i in range(steps):
            n_done += pipe.submit(*pins[i % rot]) is not None
        last = pipe.drain()
        n_done += len(last)
        wall_e2e = time.perf_counter() - t0
        for s in streams:
            torch.cuda.current_stream().wait_stream(s)
        ev1.record()
        ev1.synchronize()
        assert n_done == steps
        want_last = model.tokenize(*dev[(steps - 1) % rot]).cpu()
        assert torch.equal(last[-1][1], want_last), f"{name}: pipelined host path and the device path disagree"
        ms_e2e = ctx.max_over_ranks(ev0.elapsed_time(ev1) / steps)
        ms_wall = ctx.max_over_ranks(1e3 * wall_e2e / steps)
        h2d = sum(t.numel() * t.element_size() for t in pins[0])
        d2h = pipe.outs[0].numel() * pipe.outs[0].element_size() + 16
        out["e2e"] = dict(value=total_faces_step / (ms_e2e * 1e-3), unit=UNIT, h2d_bytes_per_step=h2d * world, d2h_bytes_per_step=d2h * world,
                          ms_per_step=ms_e2e, host_wall_ms_per_step=ms_wall,
                          api=f"HostPipeline.submit = meshtok_pipeline_submit (native; {pipeline_depth} slots in flight, one graph launch per batch, status word in pinned memory)")
        pipe.close()
    if want_cpu and rank == 0 and world == 1:
        n_cpu = min(cpu_meshes, batch)
        faces_s, times = cpu_time_oracle(oracle, name, n_cpu, repeats=3, warmup=1)
        out["cpu_baseline"] = dict(value=faces_s / min(times), unit=UNIT, cores=os.cpu_count(), kind="port",
                                   sample=f"oracle on {n_cpu} of the workload's {batch} meshes ({faces_s} faces), best of 3 after 1 warm-up, {sum(times):.1f} s of CPU work")
    del graphed
    return out


def measure_c5(ctx, oracle, model, meshes_per_launch, min_seconds):
    """BASELINE configs[4]: the sharded 100k-mesh run, in its three aspects (every mesh generated ON THE DEVICE from its global index)."""
    import synth
    from meshgpt_pytorch_b200 import sharding
    rank, world, lib, dist = ctx.rank, ctx.world, ctx.lib, ctx.dist
    Bm = meshes_per_launch
    Q, K, nvf = model.num_quantizers, model.codebook_size, model.num_vertices_per_face
    out = {}
    # ---- weak scaling: every rank streams launches of Bm meshes; 4 rotating resident batches; histogram all-reduced PER BATCH on a side stream
    res = [synth.grid_batch_device("tri", 20, 20, (k * world + rank) * Bm, Bm, "cuda") for k in range(ROTATE)]
    step_hist = torch.zeros(Q, K, dtype=torch.int64, device="cuda")
    snap = [torch.zeros_like(step_hist) for _ in range(2)]
    total_hist = torch.zeros_like(step_hist)
    g = model.graphed(res[0][0], res[0][1], histogram=step_hist)
    side = torch.cuda.Stream()
    main = torch.cuda.current_stream()
    snap_free = [torch.cuda.Event() for _ in range(2)]
    for e in snap_free:
        e.record()

    def step(i):
        v, f = res[i % ROTATE]
        g.vertices.copy_(v)          # device-to-device: the next batch of the stream becomes the graph's input
        g.faces.copy_(f)
        g.graph.replay()
        k = i & 1
        main.wait_event(snap_free[k])
        snap[k].copy_(step_hist)
        step_hist.zero_()
        side.wait_stream(main)
        with torch.cuda.stream(side):                     # overlaps the next batch's kernels
            if world > 1:
                dist.all_reduce(snap[k], op=dist.ReduceOp.SUM)
            total_hist.add_(snap[k])
            snap_free[k].record()
    for i in range(3):
        step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    step(0)
    e1.record()
    e1.synchronize()
    steps = max(8, int(math.ceil(min_seconds * 1e3 / max(e0.elapsed_time(e1), 1e-3))))
    if world > 1:
        t = torch.tensor([steps], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        steps = int(t.item())
    total_hist.zero_()
    ctx.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(steps):
        step(i)
    main.wait_stream(side)
    b.record()
    ctx.barrier()
    ms = ctx.max_over_ranks(a.elapsed_time(b))
    faces = world * steps * Bm * 800
    assert int(total_hist.sum().item()) == faces * nvf * Q, "c5 weak: histogram token count"
    g.check()
    # kernel by kernel on one launch (sections + the HBM fractions this regime is quoted for)
    lib.meshtok_profile_enable(1)
    for i in range(2):
        ctx.flush.zero_()
        model.forward(vertices=res[i][0], faces=res[i][1], return_codes=True, check_status=False)
    torch.cuda.synchronize()
    sec_ms = (C.c_float * ctx.nsec)()
    sec_n = (C.c_int32 * ctx.nsec)()
    lib.meshtok_profile_enable(0)
    ctx._lib.check(lib.meshtok_profile_collect(sec_ms, sec_n, ctx.nsec))
    taps = model.taps()
    out["weak"] = dict(value=faces / (ms * 1e-3), unit=UNIT, ms_per_launch=ms / steps, launches_timed=steps, seconds_timed=ms * 1e-3,
                       meshes_per_launch_per_rank=Bm, faces_per_launch_per_rank=Bm * 800, scaling="weak",
                       inputs=f"{ROTATE} rotating device-generated batches per rank (seeds = global mesh index), copied device-to-device into the graph's inputs inside the timed region",
                       histogram="all-reduced (NCCL) after EVERY batch on a side stream, overlapped with the next batch" if world > 1 else "accumulated per batch on a side stream (1 rank: no collective)",
                       sections=annotate_sections(model, list(sec_ms), list(sec_n), ctx.sec_names, 2, taps, Bm, 800, 441, ctx.peaks))
    del g
    # ---- host to host at this launch size
    pins = [(v.cpu().pin_memory(), f.cpu().pin_memory()) for v, f in res[:2]]
    pipe = model.host_pipeline(res[0][0], res[0][1], depth=2)
    for i in range(3):
        pipe.submit(*pins[i % 2])
    pipe.drain()
    ctx.barrier()
    n_e2e = max(4, steps // 8)
    streams = pipe.streams()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for s in streams:
        s.wait_event(ev0)
    for i in range(n_e2e):
        pipe.submit(*pins[i % 2])
    pipe.drain()
    for s in streams:
        main.wait_stream(s)
    ev1.record()
    ev1.synchronize()
    ms_e2e = ctx.max_over_ranks(ev0.elapsed_time(ev1) / n_e2e)
    out["weak"]["e2e"] = dict(value=world * Bm * 800 / (ms_e2e * 1e-3), unit=UNIT, ms_per_launch=ms_e2e,
                              h2d_bytes_per_step=world * sum(t.numel() * t.element_size() for t in pins[0]),
                              d2h_bytes_per_step=world * (pipe.outs[0].numel() * 8 + 16), api="meshtok_pipeline_submit, 2 slots")
    pipe.close()
    del pipe, pins
    # ---- strong scaling: a FIXED collection of 8 x Bm meshes (global indices 0 ..), contiguous shards, launches of <= Bm meshes
    M_total = 8 * Bm
    lo, hi = sharding.shard_range(M_total, rank, world)
    shard = [synth.grid_batch_device("tri", 20, 20, s, min(Bm, hi - s), "cuda") for s in range(lo, hi, Bm)]
    graphs = {}
    for v, f in shard:
        if v.shape[0] not in graphs:
            graphs[v.shape[0]] = model.graphed(v, f, histogram=step_hist)
    step_hist.zero_()
    for v, f in shard[:1]:
        graphs[v.shape[0]](v, f)
    step_hist.zero_()
    ctx.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for v, f in shard:
        graphs[v.shape[0]](v, f)
    if world > 1:
        dist.all_reduce(step_hist, op=dist.ReduceOp.SUM)
    b.record()
    ctx.barrier()
    ms = ctx.max_over_ranks(a.elapsed_time(b))
    assert int(step_hist.sum().item()) == M_total * 800 * nvf * Q, "c5 strong: histogram token count"
    out["strong"] = dict(value=M_total * 800 / (ms * 1e-3), unit=UNIT, ms_total=ms, collection_meshes=M_total, meshes_per_rank=hi - lo,
                         launches_per_rank=len(shard), scaling="strong", histogram="ONE all-reduce after the rank's last launch, inside the timed region")
    del graphs, shard
    # ---- ragged collection (C4-like, faces U[100, 800] by hash of the mesh index) cut by Sigma faces, ragged layout, one graph for all launches
    M_r = 4 * Bm
    idx = torch.arange(M_r, dtype=torch.int64)
    nf_all = (100 + (synth._hash01(idx, 4) * 701).long().clamp(max=700)).tolist()
    cuts = sharding.balanced_contiguous_shards(nf_all, world)
    lo, hi = cuts[rank]
    faces_rank = sum(nf_all[lo:hi])
    faces_by_rank = [sum(nf_all[a_:b_]) for a_, b_ in cuts]
    launches = []
    from meshgpt_pytorch_b200 import data as Dt
    for s in range(lo, hi, Bm):
        n = min(Bm, hi - s)
        v, f, nf = synth.ragged_batch_device("tri", 20, 20, s, n, "cuda")
        r = Dt.padded_to_ragged(v, f)
        launches.append((r["vertices"], r["faces"], r["face_offsets"].cuda(), r["vertex_offsets"].cuda()))
    step_hist.zero_()
    if launches:
        gr = model.graphed_ragged(Bm, 800, 441, max(x[1].shape[0] for x in launches), max(x[0].shape[0] for x in launches), histogram=step_hist)
        gr(*launches[0])
    step_hist.zero_()
    ctx.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for x in launches:
        gr(*x)
    if world > 1:
        dist.all_reduce(step_hist, op=dist.ReduceOp.SUM)
    b.record()
    ctx.barrier()
    ms = ctx.max_over_ranks(a.elapsed_time(b))
    total_faces = sum(nf_all)
    assert int(step_hist.sum().item()) == total_faces * nvf * Q, "c5 ragged: histogram token count"
    out["ragged_sharded"] = dict(value=total_faces / (ms * 1e-3), unit=UNIT, ms_total=ms, collection_meshes=M_r, valid_faces=total_faces,
                                 faces_by_rank=faces_by_rank, imbalance_max_over_mean=max(faces_by_rank) / (total_faces / world),
                                 sharding="balanced_contiguous_shards (Sigma faces), ragged layout (CSR of meshes), one CUDA graph serves every launch",
                                 scaling="strong")
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS), help="the headline workload (c2 = BASELINE configs[1])")
    ap.add_argument("--workloads", default="c1,c3,c4,c4pe,c5", help="extra BASELINE workloads reported in the line's `workloads` object (none: skip)")
    ap.add_argument("--c5-meshes", type=int, default=1024, help="meshes per launch per rank of the c5 measurements")
    ap.add_argument("--c5-seconds", type=float, default=1.0, help="least timed seconds of the c5 weak-scaling stream")
    ap.add_argument("--cpu-meshes", type=int, default=64, help="meshes in the CPU baseline sample (capped at the batch size)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--pipeline-depth", type=int, default=2, help="batches in flight in the host-to-host pipeline (e2e)")
    ap.add_argument("--no-hist", action="store_true", help="experiment: skip the codebook-usage histogram")
    ap.add_argument("--hist-every-step", action="store_true", help="all-reduce the histogram after every step instead of once after the last")
    ap.add_argument("--precomputed-edges", action="store_true",
                    help="pass face_edges (derived once, outside the timed region) instead of deriving them inside tokenize (BASELINE config 4)")
    ap.add_argument("--no-graph", action="store_true", help="launch the pipeline kernel by kernel instead of replaying its CUDA graph")
    ap.add_argument("--no-e2e", action="store_true", help="experiments: skip the host-to-host measurement (the line then has no e2e: not a bench value)")
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
    ctx = Ctx(args, rank, world, local)
    peaks = ctx.peaks

    oracle = build_oracle(args.workload)
    model = build_model(args.workload, oracle)
    sampler = ClockSampler(local) if rank == 0 and not args.ncu_window else None
    head = measure_workload(ctx, args.workload, oracle, model, args.steps, args.warmup, precomputed_edges=args.precomputed_edges,
                            want_e2e=not args.no_e2e, want_cpu=not args.no_cpu_baseline, cpu_meshes=args.cpu_meshes, hist_every_step=args.hist_every_step,
                            no_hist=args.no_hist, no_graph=args.no_graph, pipeline_depth=args.pipeline_depth, headline=True, sampler=sampler)
    if args.ncu_window:
        if rank == 0:
            print(json.dumps(head), flush=True)
        return
    clocks = sampler.stop() if sampler else None

    extra = {}
    wanted = [] if args.workloads in ("none", "") else [x.strip() for x in args.workloads.split(",") if x.strip()]
    sub_steps = max(5, min(args.steps, 10))
    for wname in wanted:
        base = "c4" if wname == "c4pe" else wname
        if base not in WORKLOADS or (base == args.workload and wname != "c4pe"):
            continue
        try:
            same_model = WORKLOADS[base]["cfg"] == WORKLOADS[args.workload]["cfg"]
            o2 = oracle if same_model else build_oracle(base)
            m2 = model if same_model else build_model(base, o2)
            if base == "c5":
                extra["c5"] = dict(workload="c5: " + WORKLOADS["c5"]["desc"], **measure_c5(ctx, o2, m2, args.c5_meshes, args.c5_seconds))
            else:
                r = measure_workload(ctx, base, o2, m2, sub_steps, 3, precomputed_edges=(wname == "c4pe"), want_e2e=(wname != "c4pe"),
                                     want_cpu=not args.no_cpu_baseline and wname != "c4pe", cpu_meshes=16)
                r.pop("_launches", None), r.pop("_taps", None)
                r["workload"] = base + ": " + WORKLOADS[base]["desc"]
                extra[wname] = r
            if not same_model:
                del m2
            torch.cuda.empty_cache()
        except Exception as e:   # noqa: BLE001  (an extra workload must never cost the headline line)
            extra[wname] = dict(error=f"{type(e).__name__}: {e}"[:300])

    if rank == 0:
        R = head["_taps"]["R"]
        D, K = model.dim_codebook, model.codebook_size
        sec = head["sections"]
        roofline = None
        if not model.use_residual_lfq and "rvq_search_tc" in sec:
            s = sec["rvq_search_tc"]
            per_launch_ms = s["ms_per_step"] / max(1.0, s["launches_per_step"])
            flops = 2.0 * R * K * D
            achieved = flops / (per_launch_ms * 1e-3) / 1e12
            peak = peaks.get("bf16_tflops") or 1600.0
            sust = peaks.get("bf16_tflops_sustained")
            # DRAM bytes per launch of this kernel from the committed ncu --set full capture (profiles/traffic.json), or null
            traffic, kernel_name = None, "rvq_tc2_search_kernel (CTA pair) / rvq_tc_search_kernel"
            try:
                tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
                traffic = tj["rvq_search"]["dram_bytes_per_launch"]
                kernel_name = tj["rvq_search"].get("kernel", kernel_name)
            except Exception:
                pass
            roofline = dict(kernel=kernel_name, bound="tensor", achieved=achieved, peak=peak, unit="TFLOP/s",
                            frac=achieved / peak, traffic=traffic,
                            peak_source=("MEASURED_PEAKS.json bf16_tflops (burst: the kernel is timed alone, ~130 us per launch)" if peaks else "fallback 1600 (no MEASURED_PEAKS.json)"),
                            frac_of_sustained_peak=(achieved / sust) if sust else None, sustained_peak=sust,
                            flops_per_launch=flops, ms_per_launch=per_launch_ms)
        cfg = workload_config(args.workload, world, args.precomputed_edges)
        line = dict(metric=METRIC, value=head["value"], unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=head["ms_per_step"], higher_is_better=True, scaling="weak", vs_baseline=None,
                    dtype="f32 activations (split-bf16 tcgen05 linears), f16 tcgen05 RVQ screening + f32 re-score",
                    data="synthetic", config=cfg,
                    details=dict(quantizer_rows_per_rank=R,
                                 sharding="meshes by rank, weights replicated; (Q,K) int64 histogram accumulated per rank on the device, " + ("NCCL all-reduce after every step" if args.hist_every_step else "ONE NCCL all-reduce after the last step, inside the timed region"),
                                 launch=head["launch"], wall_ms_per_step_incl_flush=head["wall_ms_per_step_incl_flush"],
                                 rvq_rows_rechecked_by_exact_scan=head["rvq_rows_rechecked_by_exact_scan"]),
                    e2e=head.get("e2e"), gpu_launches=head["_launches"], gpu_launches_per_step=head["gpu_launches_per_step"], sections=sec,
                    roofline=roofline, cpu_baseline=head.get("cpu_baseline"), clocks=clocks, workloads=extra,
                    reference_package=probe_reference_package())
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
