"""Experiment: L2-resident partitioned count+search vs the direct kernels (ms, Gbases/s)."""
import json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lrbinner_b200.profile import PartitionWorkspace, dev_count, dev_search, dev_mirror, dev_table15_partitioned
from lrbinner_b200.synth import CONFIGS, SynthSpec

dev = torch.device("cuda:0")
torch.cuda.init(); torch.zeros(1, device=dev)
cfg_name = sys.argv[2] if len(sys.argv) > 2 else "cfg2_1M_5kb_ont_k4"
cfg = CONFIGS[cfg_name]
n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else cfg["n_reads"]
spec = SynthSpec(n_reads, lengths=cfg["lengths"], errors=cfg["errors"], seed=cfg["seed"])
dr, layout = spec.device_reads(dev)
L, n = spec.total_bases, spec.n_reads
z = lambda *s: torch.zeros(s, dtype=torch.int32, device=dev)
table, hist, sums = z(2 ** 30), z(n, 10), z(n)
ws = PartitionWorkspace(dr)
def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
def direct():
    table.zero_(); hist.zero_(); sums.zero_()
    dev_count(dr, table); dev_mirror(table); dev_search(dr, table, 32, 10, hist, sums)
ms = timeit(direct)
ref = (hist.clone(), sums.clone(), table.clone())
print(json.dumps({"exp": "direct", "cfg": cfg_name, "reads": n, "Mbases": L / 1e6, "ms": ms, "Gbases_s": L / ms / 1e6}), flush=True)
for shift in (26, 25, 24):
    def fused():
        table.zero_(); hist.zero_(); sums.zero_()
        dev_table15_partitioned(dr, ws, table, True, 32, 10, hist, sums, log2_bucket_keys=shift)
        dev_mirror(table)
    ms = timeit(fused)
    ok = bool(torch.equal(hist, ref[0]) and torch.equal(sums, ref[1]) and torch.equal(table, ref[2]))
    def count_only():
        table.zero_()
        dev_table15_partitioned(dr, ws, table, True, log2_bucket_keys=shift)
    ms_c = timeit(count_only)
    def search_only():
        hist.zero_(); sums.zero_()
        dev_table15_partitioned(dr, ws, table, False, 32, 10, hist, sums, log2_bucket_keys=shift)
    ms_s = timeit(search_only)
    print(json.dumps({"exp": "partitioned", "log2_bucket_keys": shift, "buckets": 2 ** (30 - shift), "fused_ms": ms, "fused_Gbases_s": L / ms / 1e6,
                      "count_only_ms": ms_c, "search_only_ms": ms_s, "equal_to_direct": ok}), flush=True)
