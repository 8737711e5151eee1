"""One rank of tests/test_gpu_dist.py (launched by torch.distributed.run, one process per GPU).

Every multi-GPU plan of lrbinner_b200/dist.py — including the copy-engine exchange over NVLink peer memory
(PeerExchange, the plan every N > 1 benchmark number comes from) — is run for two consecutive steps on the same
tables and checked ROW FOR ROW against the oracle: row i of the output is the profile of read i
(search-15mers.cpp:26-48), the exchanged table equals the oracle's 4^15-entry table entry for entry.
Exit code 0 = every check passed on this rank.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from oracle import oracle
    from lrbinner_b200 import dist as lrb_dist
    from lrbinner_b200.profile import COMP_WIDTH, DeviceReads, PackedReads
    from lrbinner_b200.synth import SynthSpec

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n_synth = int(os.environ.get("LRB_DIST_TEST_READS", "20000"))

    # the global read set, identical on every rank: the golden community + a synthetic set with N / lowercase / edge lengths
    seqs, _ = oracle.load_reads(os.path.join(ROOT, "tests", "golden", "g06_community.fa"))
    spec = SynthSpec(n_synth, seed=91, n_rate=1e-4, lowercase_frac=0.002, edge_lengths=True, scale=0.02)
    seqs = list(seqs) + spec.host_sequences()
    n = len(seqs)
    pr = PackedReads.from_sequences(seqs, threads=8)
    dr = DeviceReads(pr, dev)
    k, bs, bc = 4, 8, 12
    P = COMP_WIDTH[k]

    # oracle: the global table (every rank builds it: the check needs no communication), rows of this rank's reads
    table_o = oracle.Table()
    for s in seqs:
        table_o.count(s)
    lo, hi = lrb_dist.own_range(n, world, rank)
    want_comp = np.stack([oracle.composition(s, k)[0] for s in seqs[lo:hi]]).astype(np.uint32)
    rows = [table_o.coverage(s, bs, bc) for s in seqs[lo:hi]]
    want_hist = np.stack([r[0] for r in rows]).astype(np.uint32)
    want_sums = np.array([r[1] for r in rows], dtype=np.uint32)

    eng = lrb_dist.CudaEngine(dr, workspace_entries=int(pr.total_bases / world * 1.3) + (1 << 20))
    px = lrb_dist.PeerExchange(dev)          # must work on the box the benchmark runs on: no silent NCCL-only fallback here
    table = px.table
    failures = []
    plans = lrb_dist.PLANS + ("readshard_ar/unpipelined", "readshard_ar/p2p", "readshard_ar/p2p+fed")
    for plan in plans:
        for step in range(2):        # two steps back to back: the next step's table.zero_() must not race the peers' pulls
            feed = None
            if plan.endswith("+fed"):   # the e2e form: reads arrive in chunks, composition + partition chunk by chunk
                rb = np.asarray(pr.read_blk)
                cuts = sorted({lo, hi, *[lo + (hi - lo) * j // 5 for j in range(1, 5)]})
                feed = [(a, b, (lambda: None)) for a, b in zip(cuts[:-1], cuts[1:])]
            res = lrb_dist.profile_distributed(eng, k, bs, bc, plan.split("/")[0], table=table,
                                               pipeline_exchange=not plan.endswith("/unpipelined"), feed=feed,
                                               peer_exchange=px if "/p2p" in plan else None)
            torch.cuda.synchronize()
            assert res["own"] == (lo, hi)
            got = {kk: res[kk].cpu().numpy().view(np.uint32) for kk in ("comp", "hist", "sums")}
            for kk, want in (("comp", want_comp), ("hist", want_hist), ("sums", want_sums)):
                if not np.array_equal(got[kk], want):
                    bad = np.flatnonzero((got[kk].reshape(len(want), -1) != want.reshape(len(want), -1)).any(axis=1))
                    failures.append(f"rank {rank} plan {plan} step {step}: {kk} differs in {len(bad)} rows, first read {lo + int(bad[0])}")
            if plan.startswith("readshard_ar") or plan == "keyshard_ag":   # these leave the whole (mirrored) table on every rank
                t = res["table"].cpu().numpy().view(np.uint32)
                if not np.array_equal(t, table_o.array):
                    d = np.flatnonzero(t != table_o.array)
                    failures.append(f"rank {rank} plan {plan} step {step}: table differs in {len(d)} entries, first key {int(d[0])}")
            dist.barrier()
    table_o.close()
    ok = torch.tensor([0 if failures else 1], device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    for f in failures:
        print("FAIL " + f, flush=True)
    if rank == 0:
        print(f"gpu_dist_worker: world {world}, {n} reads / {pr.total_bases} bases, plans {plans} x 2 steps: "
              + ("ALL ROWS AND TABLES BIT-EXACT vs oracle" if int(ok.item()) else "MISMATCH"), flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(ok.item()) else 1)


if __name__ == "__main__":
    main()
