"""Prints a compact view of a bench.py JSON line (stdin or file)."""
import signal
import json
import sys
signal.signal(signal.SIGPIPE, signal.SIG_DFL)

line = [l for l in (open(sys.argv[1]) if len(sys.argv) > 1 else sys.stdin) if l.startswith("{")][-1]
d = json.loads(line)
print(f"value {d['value']/1e6:.2f} M faces/s  ms/step {d['ms_per_step']:.3f}  e2e {((d.get('e2e') or {}).get('value') or float('nan'))/1e6:.2f} M  launches {d.get('gpu_launches')}  n_gpus {d['n_gpus']}")
for k, v in (d.get("sections") or {}).items():
    extra = ""
    if v.get("bound") == "hbm":
        extra = f"  {v['achieved_gbs']:8.0f} GB/s = {v['frac_of_hbm_peak']:.2f} of HBM peak"
    elif v.get("bound") == "tensor":
        extra = f"  {v['tensor_pipe_tflops']:8.0f} TF/s on the tensor pipe = {v['frac_of_bf16_peak']:.2f} of bf16 peak"
    print(f"  {k:18s} {v['ms_per_step']*1e3:8.1f} us  x{v['launches_per_step']:.0f}{extra}")
r = d.get("roofline")
if r:
    print(f"  roofline {r['kernel']}: {r['achieved']:.0f} {r['unit']} = {r['frac']:.3f} of {r['peak']} ({r['ms_per_launch']*1e3:.1f} us/launch)")
print("  rechecked rows:", d["config"].get("rvq_rows_rechecked_by_exact_scan"), " cpu:", (d.get("cpu_baseline") or {}).get("value"), " clocks:", d.get("clocks"))
