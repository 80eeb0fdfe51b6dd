"""Prints a compact view of a bench.py JSON line (stdin or file)."""
import json
import signal
import sys
signal.signal(signal.SIGPIPE, signal.SIG_DFL)

line = [l for l in (open(sys.argv[1]) if len(sys.argv) > 1 else sys.stdin) if l.startswith("{")][-1]
d = json.loads(line)


def sections(sec, indent="  "):
    for k, v in (sec or {}).items():
        extra = ""
        if v.get("bound") == "hbm":
            extra = f"  {v['achieved_gbs']:8.0f} GB/s = {v['frac_of_hbm_peak']:.2f} of HBM peak"
        elif v.get("bound") == "tensor":
            extra = f"  {v['tensor_pipe_tflops']:8.0f} TF/s on the tensor pipe = {v['frac_of_bf16_peak']:.2f} of burst bf16 peak"
        print(f"{indent}{k:18s} {v['ms_per_step']*1e3:8.1f} us  x{v['launches_per_step']:.0f}{extra}")


e = d.get("e2e") or {}
print(f"value {d['value']/1e6:.2f} M faces/s  ms/step {d['ms_per_step']:.3f}  e2e {(e.get('value') or float('nan'))/1e6:.2f} M ({e.get('ms_per_step')} ms dev, {e.get('host_wall_ms_per_step')} ms wall)  launches/step {d.get('gpu_launches_per_step')}  n_gpus {d['n_gpus']}")
sections(d.get("sections"))
r = d.get("roofline")
if r:
    print(f"  roofline {r['kernel']}: {r['achieved']:.0f} {r['unit']} = {r['frac']:.3f} of {r['peak']} burst ({r.get('frac_of_sustained_peak')} of sustained) ({r['ms_per_launch']*1e3:.1f} us/launch)")
print("  rechecked rows:", (d.get("details") or {}).get("rvq_rows_rechecked_by_exact_scan"), " cpu:", (d.get("cpu_baseline") or {}).get("value"), " clocks:", d.get("clocks"))
for name, w in (d.get("workloads") or {}).items():
    if "error" in w:
        print(f"  [{name}] ERROR {w['error']}")
        continue
    if name == "c5":
        for part in ("weak", "strong", "ragged_sharded"):
            p = w.get(part) or {}
            print(f"  [c5 {part}] {p.get('value', float('nan'))/1e6:.2f} M faces/s  " + "  ".join(f"{k}={p[k]}" for k in ("ms_per_launch", "launches_timed", "seconds_timed", "ms_total", "imbalance_max_over_mean") if k in p)
                  + (f"  e2e {p['e2e']['value']/1e6:.2f} M" if "e2e" in p else ""))
            if part == "weak":
                sections(p.get("sections"), "      ")
        continue
    ee = w.get("e2e") or {}
    print(f"  [{name}] {w['value']/1e6:.2f} M faces/s  {w['ms_per_step']:.3f} ms/step  e2e {(ee.get('value') or float('nan'))/1e6:.2f} M  cpu {(w.get('cpu_baseline') or {}).get('value')}")
    if "-v" in sys.argv:
        sections(w.get("sections"), "      ")
