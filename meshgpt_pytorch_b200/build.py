"""Builds libmeshtok.so (hand-written sm_100a CUDA kernels + the C ABI) in-tree with nvcc.

    python meshgpt_pytorch_b200/build.py            # incremental (run as a script: importing the package needs the .so)
    python meshgpt_pytorch_b200/build.py --force

nvcc cross-compiles without a GPU; the resulting .so is git-ignored but travels to the GPU box with
the repo snapshot.  No torch C++ extension is involved: the library is loaded with ctypes.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ_DIR = os.path.join(HERE, "lib", "obj")
LIB_PATH = os.path.join(HERE, "lib", "libmeshtok.so")

ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
                "--extended-lambda", "-Xptxas", "-v"]
# files whose float chains must round like the reference's separate fp32 ops (no FMA contraction)
NO_FMA = {"features.cu"}


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest(paths) -> str:
    h = hashlib.sha256()
    for p in sorted(paths):
        with open(p, "rb") as fh:
            h.update(p.encode() + b"\0" + fh.read())
    return h.hexdigest()


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(HERE, "..", "include", "meshtok.h"))
    return hs


def build_library(force: bool = False, verbose: bool = False, experiments: bool = False, variant: str = "", defines=()) -> str:
    """experiments: a SECOND library, lib/libmeshtok_exp.so, compiled with -DMESHTOK_EXPERIMENTS - the timing experiments that skip work
    (and so produce wrong results) exist only there; load it with MESHTOK_LIBRARY=.../libmeshtok_exp.so.  The product build has none."""
    global OBJ_DIR, LIB_PATH, COMMON_FLAGS
    if experiments:
        OBJ_DIR = os.path.join(HERE, "lib", "obj_exp")
        LIB_PATH = os.path.join(HERE, "lib", "libmeshtok_exp.so")
        COMMON_FLAGS = COMMON_FLAGS + ["-DMESHTOK_EXPERIMENTS"]
    if variant:   # A/B builds: lib/libmeshtok_<variant>.so with extra -D flags (python build.py --variant early -DMESHTOK_PDL_EARLY=0)
        OBJ_DIR = os.path.join(HERE, "lib", "obj_" + variant)
        LIB_PATH = os.path.join(HERE, "lib", f"libmeshtok_{variant}.so")
        COMMON_FLAGS = COMMON_FLAGS + list(defines)
    os.makedirs(OBJ_DIR, exist_ok=True)
    nvcc = _nvcc()
    hdr_digest = _digest(_headers())
    jobs = []
    for src in sources():
        path = os.path.join(CSRC, src)
        obj = os.path.join(OBJ_DIR, src[:-3] + ".o")
        stamp = obj + ".sha"
        want = _digest([path]) + hdr_digest + " ".join(ARCH_FLAGS + COMMON_FLAGS)
        have = open(stamp).read() if os.path.isfile(stamp) and os.path.isfile(obj) else ""
        if force or want != have:
            flags = list(COMMON_FLAGS) + (["-fmad=false"] if src in NO_FMA else [])
            jobs.append((src, [nvcc, *ARCH_FLAGS, *flags, "-c", path, "-o", obj], stamp, want))

    def run(job):
        src, cmd, stamp, want = job
        r = subprocess.run(cmd, capture_output=True, text=True)
        return src, r, stamp, want

    relink = force or not os.path.isfile(LIB_PATH) or bool(jobs)
    log_lines = []
    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as pool:
        for src, r, stamp, want in pool.map(run, jobs):
            log_lines.append(f"==== {src}\n{r.stderr}")
            if r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
                raise RuntimeError(f"nvcc failed on {src}")
            with open(stamp, "w") as fh:
                fh.write(want)
    if log_lines:
        with open(os.path.join(OBJ_DIR, "ptxas.log"), "w") as fh:
            fh.write("\n".join(log_lines))
        if verbose:
            sys.stderr.write("\n".join(log_lines))
    if relink:
        objs = [os.path.join(OBJ_DIR, s[:-3] + ".o") for s in sources()]
        cmd = [nvcc, *ARCH_FLAGS, "-shared", "-o", LIB_PATH, *objs, "-lcudart_static", "-lpthread", "-ldl", "-lrt"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return LIB_PATH


if __name__ == "__main__":
    _variant = sys.argv[sys.argv.index("--variant") + 1] if "--variant" in sys.argv else ""
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv, experiments="--experiments" in sys.argv, variant=_variant,
                        defines=[a for a in sys.argv if a.startswith("-D")]))
