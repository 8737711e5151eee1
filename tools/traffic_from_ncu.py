"""ncu launch list (--metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv) of ONE profile step
-> profiles/traffic.json: DRAM bytes, launches and serialised duration per kernel of this library.

    python tools/traffic_from_ncu.py gpurun_out/traffic_step.csv cfg2_1M_5kb_ont_k4 1000000 profiles/traffic.json
"""
import collections
import csv
import json
import re
import sys


def main(src, workload, reads, dst):
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    hdr = rows[0]
    ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    idi = hdr.index("ID")
    acc = collections.OrderedDict()
    seen = collections.defaultdict(set)
    for r in rows[1:]:
        m = re.search(r"(k2?_[a-z0-9_]+)", r[ki])
        if not m:
            continue
        name = m.group(1)
        a = acc.setdefault(name, {"dram_bytes_read": 0.0, "dram_bytes_write": 0.0, "gpu_time_ms": 0.0, "launches": 0})
        val = float(r[vi].replace(",", ""))
        if r[mi] == "dram__bytes_read.sum":
            a["dram_bytes_read"] += val
        elif r[mi] == "dram__bytes_write.sum":
            a["dram_bytes_write"] += val
        elif r[mi] == "gpu__time_duration.sum":
            a["gpu_time_ms"] += val / 1e6
        seen[name].add(r[idi])
    for name, a in acc.items():
        a["launches"] = len(seen[name])
        a["dram_bytes_per_step"] = a["dram_bytes_read"] + a["dram_bytes_write"]
    out = {"workload": workload, "reads": int(reads),
           "source": f"ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none over one whole step "
                     f"(tools/prof_step.py, launch list kept as profiles/{src.split('/')[-1]}); durations are cold-cache and serialised",
           "kernels": acc}
    with open(dst, "w") as f:
        json.dump(out, f, indent=1)
    tot = sum(a["dram_bytes_per_step"] for a in acc.values())
    print(f"{len(acc)} kernels, {tot / 1e9:.1f} GB of DRAM traffic per step")
    for name, a in acc.items():
        print(f"  {name:18s} x{a['launches']:<3d} {a['gpu_time_ms']:8.3f} ms  R {a['dram_bytes_read'] / 1e9:7.2f} GB  W {a['dram_bytes_write'] / 1e9:7.2f} GB")


if __name__ == "__main__":
    main(*sys.argv[1:5])
