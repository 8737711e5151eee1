"""Host-side overhead of lrb_profile_host: wall clock of the call vs its device time (CUDA events), per call."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from lrbinner_b200.profile import COMP_WIDTH, Context, pinned_empty
from lrbinner_b200.synth import CONFIGS, SynthSpec
cfg = CONFIGS["cfg2_1M_5kb_ont_k4"]
n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else cfg["n_reads"]
dev = torch.device("cuda", 0)
spec = SynthSpec(n_reads, lengths=cfg["lengths"], errors=cfg["errors"], seed=cfg["seed"])
dr, layout = spec.device_reads(dev)
dr.download_into(layout)
layout.index_valid(threads=os.cpu_count() or 8)
del dr
torch.cuda.empty_cache()
n, k = spec.n_reads, cfg["k"]
ctx = Context(0)
out = {"comp": pinned_empty((n, COMP_WIDTH[k])), "hist": pinned_empty((n, 10)), "sums": pinned_empty((n,))}
for i in range(6):
    t0 = time.perf_counter()
    ctx.profile(layout, k=k, bin_size=32, bins=10, out=out)
    wall = (time.perf_counter() - t0) * 1e3
    inf, tm = ctx.info(), ctx.timings()
    print(f"call {i}: python wall {wall:.2f} ms | C wall {inf['wall_ms']:.2f} plan {inf['host_plan_ms']:.2f} enqueue-done {inf['host_enqueue_ms']:.2f} | device total {tm['total']:.2f}", flush=True)
