"""Writes profiles/r02_sass_blackwell.md: per-kernel counts of the SASS mnemonics that identify the tcgen05 / TMEM / bulk-copy / TMA paths
in the built library (cuobjdump -sass; runs in the build container, no GPU needed).   python scripts/sass_evidence.py"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "meshgpt_pytorch_b200", "lib", "libmeshtok.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
pat = re.compile(r"\b(UTCHMMA(?:\.2CTA)?|UTCBAR(?:\.2CTA)?(?:\.MULTICAST)?|LDTM(?:\.x\d+)?|UBLKCP(?:\.S\.G)?|UTMALDG(?:\.\dD)?|SYNCS\.ARRIVE\.TRANS64|UCGABAR_\w+|LDGSTS|HMMA)\b")


def short(n):
    """the kernel's own name out of the mangled symbol: a length prefix followed by exactly that many characters ending in 'kernel'"""
    for m in re.finditer(r"(\d{1,2})(?=[A-Za-z_])", n):
        cand = n[m.end(): m.end() + int(m.group(1))]
        if cand.endswith("kernel"):
            return cand
    return n[:60]


cur, per, examples = None, collections.OrderedDict(), {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = short(m.group(1))
        continue
    if cur is None:
        continue
    for mm in pat.finditer(line):
        k = mm.group(1)
        per.setdefault(cur, collections.Counter())[k] += 1
        examples.setdefault(k, re.sub(r"\s+/\*.*$", "", line.strip())[:120])
tot = collections.Counter()
for c in per.values():
    tot.update(c)
keys = sorted(tot)
L = ["# SASS evidence (round 2): Blackwell instructions in libmeshtok.so", "",
     "`python scripts/sass_evidence.py` = `cuobjdump -sass meshgpt_pytorch_b200/lib/libmeshtok.so` (sm_100a only), run in the build container at the",
     "commit this file belongs to: counts of the mnemonics that identify the tcgen05 / TMEM / bulk-copy / TMA paths, per kernel (all",
     "instantiations of a kernel summed).", "",
     "| mnemonic | what it is |", "|---|---|",
     "| `UTCHMMA`, `UTCHMMA.2CTA` | `tcgen05.mma` (single CTA / `cta_group::2` CTA pair): 5th-generation tensor-core MMA, accumulators in TMEM |",
     "| `LDTM.x16` / `.x32` | `tcgen05.ld`: accumulator read-back from TMEM in the epilogues |",
     "| `UTCBAR(.2CTA)(.MULTICAST)` | `tcgen05.commit` -> mbarrier arrive (multicast to both CTAs of a pair) |",
     "| `UBLKCP.S.G` | `cp.async.bulk.shared.global`: 1-D bulk copies of pre-tiled operand images, one elected thread per stage |",
     "| `UTMALDG.2D` | `cp.async.bulk.tensor.2d`: tensor-map TMA loads (the x slices of the neighbour-aggregation kernel, round 2) |",
     "| `SYNCS.ARRIVE.TRANS64` | `mbarrier.arrive.expect_tx` |",
     "| `UCGABAR_ARV` / `_WAIT` | `barrier.cluster` of the CTA-pair kernels |",
     "| `LDGSTS` | `cp.async` 16-byte copies (round-1 staging paths, kept as A/B partners) |",
     "| `HMMA` | legacy `mma.sync`: must be absent |", "",
     "| kernel | " + " | ".join(f"`{k}`" for k in keys) + " |", "|---|" + "---:|" * len(keys)]
for fn, c in per.items():
    if any(k.startswith(("UTC", "LDTM", "UBLK", "UTMA")) for k in c):
        L.append(f"| `{fn}` | " + " | ".join(str(c.get(k, "")) for k in keys) + " |")
L.append("| **whole library** | " + " | ".join(str(tot.get(k, 0)) for k in keys) + " |")
L += ["", "One instruction of each kind as it appears in the SASS:", "", "```"] + [examples[k] for k in keys] + ["```", "",
      f"`HMMA` occurrences: {tot.get('HMMA', 0)} - no kernel falls back to `mma.sync`."]
path = os.path.join(ROOT, "profiles", "r02_sass_blackwell.md")
open(path, "w").write("\n".join(L) + "\n")
print(path)
