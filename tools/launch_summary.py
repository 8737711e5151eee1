"""Sum an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (profiles/*_launches_*_summary.csv)."""
import csv, sys, collections
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if not l.startswith("=="))]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[1:]:
    if len(r) <= vi: continue
    v = float(r[vi].replace(",", ""))
    v = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui].replace("usecond", "us").replace("nsecond", "ns").replace("msecond", "ms"), 1e-6) * v
    name = r[ki].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
    tot[name] += v; cnt[name] += 1
allms = sum(tot.values())
print("kernel,launches,total_ms,share_of_all_launch_time")
for k, v in tot.most_common():
    print(f'"{k}",{cnt[k]},{v:.3f},{v / allms:.4f}')
