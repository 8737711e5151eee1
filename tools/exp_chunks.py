"""Cost of building the partition in C chunks (as the host pipeline does) on device-resident reads: separates the
chunking inefficiency (kernel tails, drains between the per-chunk kernels) from PCIe interference."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from lrbinner_b200.profile import PartitionWorkspace
from lrbinner_b200.synth import CONFIGS, SynthSpec
cfg = CONFIGS["cfg2_1M_5kb_ont_k4"]
dev = torch.device("cuda", 0)
spec = SynthSpec(cfg["n_reads"], lengths=cfg["lengths"], errors=cfg["errors"], seed=cfg["seed"])
dr, layout = spec.device_reads(dev)
ws = PartitionWorkspace(dr)
rb = np.array(layout.read_blk)
nb, n = layout.n_blocks, layout.n_reads
for chunks in (1, 4, 16, 32):
    cuts = [0] + [int(rb[np.searchsorted(rb, nb * i // chunks, side="right") - 1]) for i in range(1, chunks)] + [nb]
    for count in (False, True):
        ms = []
        for rep in range(4):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            ws.begin(True, count=count)
            for lo, hi in zip(cuts[:-1], cuts[1:]):
                ws.add(lo, hi)
            b.record()
            torch.cuda.synchronize()
            ms.append(a.elapsed_time(b))
        print(json.dumps({"chunks": chunks, "second_level": count, "ms": round(float(np.median(ms[1:])), 3)}), flush=True)
