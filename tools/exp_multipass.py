"""Experiment: key-range multi-pass count/search (every pass re-scans the stream, only keys of one slice are
touched, so the slice can live in L2).  Prints ms per configuration."""
import json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lrbinner_b200.profile import dev_count, dev_search, dev_mirror
from lrbinner_b200.synth import CONFIGS, SynthSpec

dev = torch.device("cuda:0")
torch.cuda.init(); torch.zeros(1, device=dev)
cfg = CONFIGS["cfg2_1M_5kb_ont_k4"]
n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 400000
spec = SynthSpec(n_reads, lengths=cfg["lengths"], errors=cfg["errors"], seed=cfg["seed"])
dr, layout = spec.device_reads(dev)
L = spec.total_bases
table = torch.zeros(2 ** 30, dtype=torch.int32, device=dev)
hist = torch.zeros((spec.n_reads, 10), dtype=torch.int32, device=dev)
sums = torch.zeros(spec.n_reads, dtype=torch.int32, device=dev)
def timeit(fn, reps=2):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
ref = None
for npass in (1, 2, 4, 8, 16, 32, 64):
    step = 2 ** 30 // npass
    def count():
        table.zero_()
        for p in range(npass):
            dev_count(dr, table, key_lo=p * step, key_hi=(p + 1) * step)
    ms = timeit(count)
    chk = int((table.to(torch.int64) & 0xFFFFFFFF).sum().item())
    if ref is None: ref = chk
    assert chk == ref
    def search():
        hist.zero_(); sums.zero_()
        for p in range(npass):
            dev_search(dr, table, 32, 10, hist, sums, key_lo=p * step, key_hi=(p + 1) * step)
    ms2 = timeit(search)
    assert int(sums.to(torch.int64).sum().item()) == ref
    print(json.dumps({"exp": "multipass", "reads": spec.n_reads, "Mbases": L / 1e6, "passes": npass, "count_ms": ms, "search_ms": ms2,
                      "count_Gbases_s": L / ms / 1e6, "search_Gbases_s": L / ms2 / 1e6}), flush=True)
