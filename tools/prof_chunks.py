"""Partition build in 16 chunks on device-resident reads, for ncu captures of a per-chunk k2_partition launch."""
import os, sys
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
nb = layout.n_blocks
chunks = int(sys.argv[1]) if len(sys.argv) > 1 else 16
cuts = [0] + [int(rb[np.searchsorted(rb, nb * i // chunks, side="right") - 1]) for i in range(1, chunks)] + [nb]
ws.begin(True, count=True)
for lo, hi in zip(cuts[:-1], cuts[1:]):
    ws.add(lo, hi)
torch.cuda.synchronize()
