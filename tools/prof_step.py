"""One device-resident profile step on a reduced read set, for ncu captures (never a bench value)."""
import argparse, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lrbinner_b200.profile import COMP_WIDTH, PartitionWorkspace, dev_composition, dev_mirror
from lrbinner_b200.synth import CONFIGS, SynthSpec

ap = argparse.ArgumentParser()
ap.add_argument("--reads", type=int, default=200000)
ap.add_argument("--config", default="cfg2_1M_5kb_ont_k4")
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--bucket-log2", type=int, default=24)
a = ap.parse_args()
cfg = CONFIGS[a.config]
dev = torch.device("cuda", 0)
spec = SynthSpec(a.reads, lengths=cfg["lengths"], errors=cfg["errors"], seed=cfg["seed"])
dr, _ = spec.device_reads(dev)
k, P = cfg["k"], COMP_WIDTH[cfg["k"]]
n = spec.n_reads
table = torch.zeros(2 ** 30, dtype=torch.int32, device=dev)
comp = torch.zeros((n, P), dtype=torch.int32, device=dev)
hist = torch.zeros((n, 10), dtype=torch.int32, device=dev)
sums = torch.zeros(n, dtype=torch.int32, device=dev)
ws = PartitionWorkspace(dr)
for _ in range(a.steps):
    comp.zero_(); hist.zero_(); sums.zero_()          # the table is WRITTEN by the count (apply mode bit 3): no memset
    dev_composition(dr, k, comp)
    ws.build(True, log2_bucket_keys=a.bucket_log2)
    ws.apply(table, count=True, overwrite=True)
    ws.apply(table, count=False, search=True, bin_size=32, bins=10, hist=hist, sums=sums)
    dev_mirror(table)
torch.cuda.synchronize()
print("windows", int(sums.to(torch.int64).sum()), "bases", spec.total_bases)
