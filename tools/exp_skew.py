"""Throughput on a key-skewed read set (verdict r01 weak #9): a fraction of the reads are low-complexity (homopolymers,
tandem repeats of a short unit), so whole buckets overflow the second-level lists and are counted by the L2-atomic kernel
(same-address REDs), and the search hammers a few table entries.  Prints one JSON line per set: per-kernel ms per step
from the library's launch timers, the buckets that fell back, Gbases/s.  Device-resident, like bench.py's `value`.

    python tools/exp_skew.py [--reads 200000] [--fractions 0,0.01,0.05,0.2]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from lrbinner_b200 import _lib
from lrbinner_b200.profile import COMP_WIDTH, DeviceReads, PackedReads, PartitionWorkspace, dev_composition, dev_mirror
from lrbinner_b200.synth import SynthSpec

ap = argparse.ArgumentParser()
ap.add_argument("--reads", type=int, default=200000)
ap.add_argument("--fractions", default="0,0.01,0.05,0.2")
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--gc", default="", help="comma-separated GC contents: additional sets drawn from genomes of that base composition "
                                         "(e.g. 0.3,0.2): a skewed KEY distribution over the (bucket, sub-slice) cells without hot keys")
a = ap.parse_args()
dev = torch.device("cuda", 0)
k, P, BS, BC = 4, COMP_WIDTH[4], 32, 10
spec = SynthSpec(a.reads, lengths="gamma5k", errors="ont", seed=22)
base = spec.host_sequences()
rng = np.random.default_rng(3)
def gc_reads(gc, lengths, rng):
    """error-free reads, both strands, from one 20 Mbp genome with P(G) = P(C) = gc / 2"""
    p = [(1 - gc) / 2, gc / 2, (1 - gc) / 2, gc / 2]                 # A C T G in code order
    genome = rng.choice(np.frombuffer(b"ACTG", dtype=np.uint8), size=20_000_000, p=p)
    comp = np.zeros(256, dtype=np.uint8)
    comp[list(b"ACGT")] = list(b"TGCA")
    out = []
    for L in lengths:
        st = int(rng.integers(0, len(genome) - int(L)))
        r = genome[st:st + int(L)]
        if rng.random() < 0.5:
            r = comp[r[::-1]]
        out.append(r.tobytes())
    return out


sets = [("low_complexity", float(x)) for x in a.fractions.split(",") if x != ""] + [("gc", float(x)) for x in a.gc.split(",") if x != ""]
for kind, frac in sets:
    seqs = list(base) if kind == "low_complexity" else gc_reads(frac, spec.lengths, rng)
    n_low = int(frac * len(seqs)) if kind == "low_complexity" else 0
    for i in rng.choice(len(seqs), size=n_low, replace=False):
        L = len(seqs[i])
        if rng.random() < 0.3:
            seqs[i] = bytes([b"ACGT"[int(rng.integers(0, 4))]]) * L                     # homopolymer: one key
        else:
            unit = bytes(rng.choice(list(b"ACGT"), size=int(rng.integers(2, 12))).tolist())  # tandem repeat: <= 11 keys
            seqs[i] = (unit * (L // len(unit) + 1))[:L]
    pr = PackedReads.from_sequences(seqs, threads=os.cpu_count() or 8)
    dr = DeviceReads(pr, dev)
    n, Lb = pr.n_reads, pr.total_bases
    table = torch.zeros(2 ** 30, dtype=torch.int32, device=dev)
    comp = torch.zeros((n, P), dtype=torch.int32, device=dev)
    hist = torch.zeros((n, BC), dtype=torch.int32, device=dev)
    sums = torch.zeros(n, dtype=torch.int32, device=dev)
    ws = PartitionWorkspace(dr)

    def step():
        comp.zero_(); hist.zero_(); sums.zero_()
        dev_composition(dr, k, comp)
        ws.build(True)
        ws.apply(table, count=True, overwrite=True)
        ws.apply(table, count=False, search=True, bin_size=BS, bins=BC, hist=hist, sums=sums)
        dev_mirror(table)

    step(); step()
    torch.cuda.synchronize()
    _lib.lib.lrb_prof_enable(1)
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(a.steps):
        step()
    t1.record()
    torch.cuda.synchronize()
    _lib.lib.lrb_prof_enable(0)
    prof = _lib.prof_report()
    ms = t0.elapsed_time(t1) / a.steps
    # PartMeta (csrc/partition.cu) in u64 units: counts, offsets [64][64], chunk_base[65], needed, overflow, overflow2 u32[64], spill_n
    meta = ws.small.cpu().numpy()
    o2 = meta[2 * 64 * 64 + 65 + 2:2 * 64 * 64 + 65 + 2 + 32].view(np.uint32)
    spill_n = int(meta[2 * 64 * 64 + 65 + 2 + 32])
    V = int(sums.to(torch.int64).sum().item())
    assert int(hist.to(torch.int64).sum().item()) == V
    print(json.dumps({("low_complexity_fraction" if kind == "low_complexity" else "genome_gc_content"): frac, "reads": n, "bases": Lb, "valid_windows": V, "ms_per_step": ms,
                      "Gbases_per_s": Lb / ms / 1e6, "buckets_counted_by_L2_atomics": int(o2[:ws.part.n_buckets].astype(bool).sum()), "entries_through_the_spill_area": spill_n,
                      "kernel_ms_per_step": {kk: round(v[1] / a.steps, 3) for kk, v in prof.items() if v[1] / a.steps >= 0.01}}), flush=True)
    del ws, table, comp, hist, sums, dr, pr
    torch.cuda.empty_cache()
