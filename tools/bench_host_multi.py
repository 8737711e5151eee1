"""Throughput of the PRODUCT multi-GPU path: one process driving N devices through lrb_ctx_create_multi / lrb_profile_host
(what run_profile(n_gpus=N) runs after parsing) — host buffers in, host buffers out, wall clock of the call.

    python tools/bench_host_multi.py N [shards_of_config2_per_gpu=1]
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from lrbinner_b200.profile import COMP_WIDTH, Context, pinned_empty
from lrbinner_b200.synth import CONFIGS, SynthSpec

n_gpus = int(sys.argv[1]) if len(sys.argv) > 1 else 2
cfg = CONFIGS["cfg2_1M_5kb_ont_k4"]
n_reads = cfg["n_reads"] * n_gpus * (int(sys.argv[2]) if len(sys.argv) > 2 else 1)
dev = torch.device("cuda", 0)
spec = SynthSpec(n_reads, lengths=cfg["lengths"], errors=cfg["errors"], seed=cfg["seed"])
dr, layout = spec.device_reads(dev)          # generated on GPU 0, brought home into the pinned packed layout
dr.download_into(layout)
layout.index_valid(threads=os.cpu_count() or 8)
del dr
torch.cuda.empty_cache()
n, k, L = spec.n_reads, cfg["k"], spec.total_bases
out = {"comp": pinned_empty((n, COMP_WIDTH[k])), "hist": pinned_empty((n, 10)), "sums": pinned_empty((n,))}
res = {}
for devices in ([0], list(range(n_gpus))):
    ctx = Context(devices)
    walls = []
    for i in range(int(sys.argv[3]) if len(sys.argv) > 3 else 5):
        t0 = time.perf_counter()
        ctx.profile(layout, k=k, bin_size=32, bins=10, out=out)
        walls.append((time.perf_counter() - t0) * 1e3)
    info = ctx.info()
    res[len(devices)] = {"wall_ms_every_call": [round(w, 2) for w in walls], "Gbases_per_s": L / min(walls[1:]) / 1e6, "run": info,
                         "digest": [int(out["sums"].astype(np.int64).sum()), int(out["hist"].astype(np.int64).sum()), int(out["comp"].astype(np.int64).sum())]}
    if len(devices) == 1:
        ref = {kk: v.copy() for kk, v in out.items()}
    else:
        res[len(devices)]["rows_equal_single_gpu"] = bool(all(np.array_equal(ref[kk], out[kk]) for kk in out))
    ctx.close()
print(json.dumps({"reads": n, "bases": L, "n_gpus": n_gpus, "one_gpu": res[1], "n_gpu": res[n_gpus]}))
