"""world_size-2 gloo test of the multi-GPU decomposition logic (lrbinner_b200/dist.py) on the CPU.

The per-GPU work is injected: an oracle-backed engine (tests only) stands in for the CUDA kernels, so what
is checked here is the sharding (key ranges, read ranges, padding), the collectives and their order — for
all three plans, against the single-process oracle result."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from conftest import GOLDEN, ROOT

K, BS, BC = 3, 10, 8
KT = 9   # k-mer length of the table in this test: 4^9 entries instead of 4^15, same plumbing (see OracleEngine)


class OracleEngine:
    """Test stand-in for CudaEngine: same interface, results from oracle/ (CPU restatement).  It runs the
    table passes with 9-mers (oracle.SmallTable, the reference's rolling/reset/both-strands rule with k as a
    parameter) so that the collectives move 1 MiB instead of 4 GiB; the k = 15 arithmetic itself is pinned
    elsewhere (tests/test_oracle_pins.py, tests/test_host_cpu.py, tests/test_gpu_parity.py)."""

    def __init__(self, seqs):
        from oracle import oracle
        self.o = oracle
        self.seqs = seqs
        self.n_reads = len(seqs)
        self.table_entries = 4 ** KT
        self.canon_bit = KT          # middle base of a 9-mer = base 4 -> its high bit is bit 9
        self._sparse = {}

    def zeros(self, shape):
        return torch.zeros(shape, dtype=torch.int32)

    def _read_keys(self, i):
        """middle-bit-clear keys of read i with their window multiplicity (= what the count kernel increments)."""
        if i not in self._sparse:
            t = self.o.SmallTable(KT)
            t.count(self.seqs[i])
            keys = np.flatnonzero(t.array)
            keys = keys[(keys & (1 << KT)) == 0]
            self._sparse[i] = (keys, t.array[keys].astype(np.int64))
        return self._sparse[i]

    def composition(self, k, comp, lo, hi):
        for i in range(lo, hi):
            comp[i] += torch.from_numpy(self.o.composition(self.seqs[i], k)[0].astype(np.int32))

    def count(self, table, klo, khi, lo, hi):
        tv = table.numpy().view(np.uint32)
        for i in range(lo, hi):
            keys, mult = self._read_keys(i)
            sel = (keys >= klo) & (keys < khi)
            np.add.at(tv, keys[sel], mult[sel].astype(np.uint32))

    def mirror(self, table):
        tv = table.numpy().view(np.uint32)
        nz = np.flatnonzero(tv)
        for x in nz[(nz & (1 << KT)) == 0]:
            tv[self.o.revcomp(int(x), KT)] = tv[x]

    def search(self, table, bs, bc, hist, sums, lo, hi, klo, khi):
        tv = table.numpy().view(np.uint32)
        for i in range(lo, hi):
            keys, mult = self._read_keys(i)
            sel = (keys >= klo) & (keys < khi)
            for x, m in zip(keys[sel], mult[sel]):
                hist[i, self.o.bucket(int(tv[x]), bs, bc)] += int(m)
                sums[i] += int(m)


    def count_fed(self, table, lo, hi, chunks, per_chunk):
        for j, (a, b) in enumerate(chunks):
            per_chunk(j)
            self.count(table, 0, self.table_entries, a, b)

    # slice-wise search: plan X pipelines the table exchange with it (4 key slices here, 64 buckets in CudaEngine)
    N_SLICES = 4

    def n_slices(self, lo, hi):
        return self.N_SLICES

    def slice_keys(self, i):
        w = self.table_entries // self.N_SLICES
        return i * w, (i + 1) * w

    def search_slice(self, table, bs, bc, hist, sums, lo, hi, i, i_end=None):
        for j in range(i, i + 1 if i_end is None else i_end):
            self.search(table, bs, bc, hist, sums, lo, hi, *self.slice_keys(j))


def _worker(rank, world, port, plan, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    from lrbinner_b200 import dist as lrb_dist
    from oracle import oracle
    dist.init_process_group("gloo", rank=rank, world_size=world)
    seqs, _ = oracle.load_reads(os.path.join(GOLDEN, "g06_community.fa"))
    seqs = seqs[:41]                                   # odd count: exercises the padded last chunk
    eng = OracleEngine(seqs)
    simple = plan.endswith("/unpipelined")
    feed = None
    if plan.endswith("/fed"):      # this rank's reads arrive in three chunks (the e2e pipeline); the waits are no-ops on the CPU
        lo, hi = lrb_dist.own_range(len(seqs), world, rank)
        cuts = sorted({lo, lo + (hi - lo) // 3, lo + (hi - lo) // 2, hi})
        arrived = []
        feed = [(a, b, (lambda j=j: arrived.append(j))) for j, (a, b) in enumerate(zip(cuts[:-1], cuts[1:]))]
    res = lrb_dist.profile_distributed(eng, K, BS, BC, plan.split("/")[0], pipeline_exchange=not simple, feed=feed)
    if feed is not None:
        assert arrived == list(range(len(feed)))
    lo, hi = res["own"]
    q.put((rank, lo, hi, res["comp"].numpy().copy(), res["hist"].numpy().copy(), res["sums"].numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("plan", ["keyshard_rs", "keyshard_ag", "readshard_ar", "readshard_ar/unpipelined", "readshard_ar/fed"])
def test_two_rank_plans_match_single_process_oracle(plan):
    from oracle import oracle
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, plan, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    seqs, _ = oracle.load_reads(os.path.join(GOLDEN, "g06_community.fa"))
    seqs = seqs[:41]
    table = oracle.SmallTable(KT)
    for s in seqs:
        table.count(s)
    covered = []
    for rank, lo, hi, comp, hist, sums in sorted(got):
        covered += list(range(lo, hi))
        assert comp.shape[0] == hist.shape[0] == sums.shape[0] == hi - lo
        for row, i in enumerate(range(lo, hi)):
            assert np.array_equal(comp[row].astype(np.uint32), oracle.composition(seqs[i], K)[0].astype(np.uint32)), (plan, i)
            raw, total, _ = table.coverage(seqs[i], BS, BC)
            assert np.array_equal(hist[row].astype(np.uint32), raw.astype(np.uint32)) and int(sums[row]) == total, (plan, i)
    assert covered == list(range(len(seqs)))


def _digest_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    import bench_multi
    from lrbinner_b200 import dist as lrb_dist
    from oracle import oracle
    dist.init_process_group("gloo", rank=rank, world_size=world)
    seqs, _ = oracle.load_reads(os.path.join(GOLDEN, "g06_community.fa"))
    eng = OracleEngine(seqs[:41])
    out = {}
    for plan in ("keyshard_ag", "readshard_ar"):       # both leave the whole table on every rank; different collectives
        res = lrb_dist.profile_distributed(eng, K, BS, BC, plan, pipeline_exchange=False)
        out[plan] = bench_multi.content_digest(torch, dist, res, BC, res["table"], canon_bit=KT)
    # what the old mass check could not see: a row on the wrong read, a count in the wrong bin, a lost table update
    res["hist"] = res["hist"].clone()
    if rank == 1 and res["hist"].shape[0] > 1:
        res["hist"][[0, 1]] = res["hist"][[1, 0]]                         # two rows swapped (same totals)
    out["rows_swapped"] = bench_multi.content_digest(torch, dist, res, BC, res["table"], canon_bit=KT)
    res["hist"] = res["hist"].clone()
    if rank == 0:
        row = int(torch.nonzero(res["hist"].sum(dim=1))[0])
        col = int(torch.nonzero(res["hist"][row])[0])
        res["hist"][row, col] -= 1
        res["hist"][row, (col + 1) % BC] += 1                             # one window in the wrong bin (same totals)
    out["wrong_bin"] = bench_multi.content_digest(torch, dist, res, BC, res["table"], canon_bit=KT)
    t = res["table"].clone()
    nz = torch.nonzero(t.view(-1, 2, 1 << KT)[:, 0, :].reshape(-1))
    i = int(nz[len(nz) // 2])
    flat = t.view(-1, 2, 1 << KT)
    flat[i >> KT, 0, i & ((1 << KT) - 1)] -= 1                             # one lost update in the canonical half
    out["lost_update"] = {"table": bench_multi.table_digest(torch, t, KT)}
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def test_verify_digests_agree_across_plans_and_catch_what_totals_cannot():
    """bench_multi's `verify` digests (position-weighted sums mod 2^64, all-reduced): equal across plans that use
    different collectives, different as soon as a row moves, a window changes bin or a table update is lost — the
    failures the round-1 mass check (sum of sums, of hist, of comp) was blind to."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_digest_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    a, b = got[0], got[1]
    assert a["keyshard_ag"] == a["readshard_ar"] == b["keyshard_ag"] == b["readshard_ar"]
    ok = a["readshard_ar"]
    assert a["rows_swapped"]["hist"] != ok["hist"] and a["rows_swapped"]["sums"] == ok["sums"]
    assert a["wrong_bin"]["hist"] != ok["hist"] and a["wrong_bin"]["hist"] != a["rows_swapped"]["hist"]
    assert a["lost_update"]["table"] != ok["table"] and b["lost_update"]["table"] != ok["table"]


def test_shard_arithmetic():
    from lrbinner_b200 import dist as d
    for n in (0, 1, 7, 8, 41, 1000):
        for w in (1, 2, 3, 4, 8):
            spans = [d.own_range(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n and all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert all(hi - lo <= d.chunk_size(n, w) for lo, hi in spans)
    for w in (1, 2, 4, 8):
        ks = [d.key_range(w, r) for r in range(w)]
        assert ks[0][0] == 0 and ks[-1][1] == 2 ** 30 == d.TABLE_ENTRIES and all(a[1] == b[0] for a, b in zip(ks, ks[1:]))


def test_peer_exchange_schedule_covers_every_slice_once():
    from lrbinner_b200 import dist as d
    for n_slices in (1, 4, 16, 63, 64):
        for world in (1, 2, 3, 4, 5, 7, 8):
            for group_of in (0, 1, 3):
                pieces, rounds = d.exchange_schedule(n_slices, world, group_of)
                assert pieces[0][0] == 0 and pieces[-1][1] == n_slices and all(a[1] == b[0] for a, b in zip(pieces, pieces[1:]))
                seen = [g for rnd in rounds for g, _ in rnd]
                assert seen == list(range(len(pieces)))                      # every piece in exactly one round, in order
                for rnd in rounds:
                    owners = [o for _, o in rnd]
                    assert len(set(owners)) == len(owners) and all(o == g % world for g, o in rnd)   # one piece per rank and round
                    assert [g for g, _ in rnd] == list(range(rnd[0][0], rnd[0][0] + len(rnd)))      # contiguous: one search launch
    pieces, rounds = d.exchange_schedule(64, 8)
    assert len(rounds) == 8 and all(b - a == 1 for a, b in pieces)            # 8 GPUs: 8 rounds of 8 single-bucket pieces
    pieces, rounds = d.exchange_schedule(64, 2)
    assert len(rounds) == 8 and all(b - a == 4 for a, b in pieces)
