"""Compact view of bench.py JSON lines: step / e2e ms and per-kernel ms (python tools/bench_summary.py file.json ...)."""
import json
import sys

for path in sys.argv[1:]:
    try:
        line = [l for l in open(path).read().splitlines() if l.startswith("{")][-1]
        d = json.loads(line)
    except Exception as ex:
        print(f"{path}: unreadable ({ex})")
        continue
    ks = " ".join(f"{k.replace('k_', '')}={v['ms_per_step']:.2f}" for k, v in d.get("kernels", {}).items() if v["ms_per_step"] >= 0.05)
    e2e = d.get("e2e") or {}
    print(f"{path}: value {d['value']:.1f} Gb/s step {d['ms_per_step']:.2f} ms | e2e {e2e.get('value', 0):.1f} Gb/s {e2e.get('ms_per_step', 0):.2f} ms | launches {d.get('gpu_launches')} | {ks}")
    if d.get("phases"):
        print("    phases", {k: round(v, 2) for k, v in d["phases"]["ms"].items()})
    if d.get("file_level"):
        fl = d["file_level"]
        print("    file level: ref %.2fs  gpu 3 calls %.2fs %s  fused %.2fs  parity %s" % (fl["reference_tools_s"], fl["gpu_three_calls_s"],
              [round(x, 2) for x in fl["gpu_three_calls_s_each"]], fl["gpu_fused_run_profile_s"], fl["parity_vs_reference_tools"]))
    if d.get("cpu_baseline"):
        cb = d["cpu_baseline"]
        print("    cpu baseline: %.4f Gb/s (%s cores), net of fixed %.4f, fixed %.1fs" % (cb["value"], cb["cores"], cb.get("value_net_of_fixed_cost", 0), cb.get("fixed_seconds", 0)))
