"""CPU tier: host code of liblrb200 (ingest, packing layout, exact %f, writers) and the kernels'
per-lane arithmetic (csrc/lane_core.cuh run by tests/host_emul.cpp) against the oracle and the golden
fixtures.  No compute entry point of the library is called here (there is no GPU)."""
import ctypes as C
import gzip
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import COV_PARAMS, GOLDEN, ROOT, golden_inputs
from oracle import oracle

from lrbinner_b200 import _lib
from lrbinner_b200.profile import COMP_WIDTH, PackedReads, kmer_lut

EMUL_SO = os.path.join(ROOT, "tests", "_build", "libemul.so")


@pytest.fixture(scope="session")
def emul():
    src = os.path.join(ROOT, "tests", "host_emul.cpp")
    deps = [src, os.path.join(ROOT, "lrbinner_b200", "csrc", "lane_core.cuh"), os.path.join(ROOT, "lrbinner_b200", "csrc", "fixed6.h")]
    if not os.path.exists(EMUL_SO) or os.path.getmtime(EMUL_SO) < max(os.path.getmtime(d) for d in deps):
        os.makedirs(os.path.dirname(EMUL_SO), exist_ok=True)
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-Wno-unknown-pragmas", "-fPIC", "-shared", "-o", EMUL_SO, src])
    L = C.CDLL(EMUL_SO)
    L.emul_mirror_check.restype = C.c_uint64
    L.emul_coverage_bin.restype = C.c_uint32
    L.emul_coverage_bin.argtypes = [C.c_uint32, C.c_long, C.c_int]
    L.emul_revcomp15.restype = C.c_uint32
    L.emul_revcomp15.argtypes = [C.c_uint32]
    return L


def _gz(path):
    with gzip.open(path, "rb") as f:
        return f.read()


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


# ---- C ABI surface ---------------------------------------------------------------------------------

def test_every_declared_symbol_is_exported():
    hdr = open(os.path.join(ROOT, "include", "lrbinner_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(lrb_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 30
    for name in sorted(declared):
        assert hasattr(_lib.lib, name), f"{name} declared in include/lrbinner_b200.h but not exported"
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)


def test_no_device_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    rc = _lib.lib.lrb_ctx_create(0, C.byref(h))
    assert rc == _lib.LRB_ECUDA and "no CPU path" in _lib.last_error()
    # the file-level drop-ins report failure (non-zero) instead of silently computing somewhere else
    assert _lib.lib.lrb_count_kmers(os.path.join(GOLDEN, "g03_edge.fa").encode(), b"/tmp/_lrb_never", 3, 1) != 0


# ---- ingest ----------------------------------------------------------------------------------------

@pytest.mark.parametrize("name", golden_inputs())
def test_ingest_matches_oracle_reader(name):
    seqs, _ = oracle.load_reads(os.path.join(GOLDEN, name))
    pr = PackedReads.from_file(os.path.join(GOLDEN, name), threads=3)
    assert pr.n_reads == len(seqs)
    assert list(pr.read_len) == [len(s) for s in seqs]
    _check_packing(pr, seqs)


def _check_packing(pr, seqs):
    assert pr.total_bases == sum(len(s) for s in seqs)
    blk = 0
    codes, valid = pr.codes, pr.valid
    for i, s in enumerate(seqs):
        assert pr.read_blk[i] == blk
        a = np.frombuffer(s, dtype=np.uint8)
        nb = len(s) // 32 + 1
        want_code = np.zeros(nb * 32, dtype=np.uint32)
        want_code[:len(s)] = (a >> 1) & 3
        want_valid = np.zeros(nb * 32, dtype=bool)
        want_valid[:len(s)] = np.isin(a, np.frombuffer(b"ACGT", dtype=np.uint8))
        words = codes[2 * blk: 2 * (blk + nb)]
        got_code = ((words[:, None] >> (30 - 2 * np.arange(16, dtype=np.uint32))[None, :]) & 3).reshape(-1)
        assert np.array_equal(got_code, want_code), i
        got_valid = ((valid[blk: blk + nb][:, None] >> np.arange(32, dtype=np.uint32)[None, :]) & 1).reshape(-1).astype(bool)
        assert np.array_equal(got_valid, want_valid), i
        blk += nb
    assert pr.read_blk[len(seqs)] == blk == pr.n_blocks
    # tiles: <= 256 blocks of one read each, in order, covering every block
    t = 0
    for i in range(len(seqs)):
        b = int(pr.read_blk[i])
        while b < pr.read_blk[i + 1]:
            assert pr.tile_read[t] == i and pr.tile_blk[t] == b
            b += _lib.TILE_BLOCKS
            t += 1
    assert t == pr.n_tiles


def test_ingest_fuzz_against_oracle(tmp_path):
    rng = np.random.default_rng(7)
    alphabet = np.frombuffer(b"ACGTacgtN>@+\n\r \t;-", dtype=np.uint8)
    probs = np.array([8, 8, 8, 8, 1, 1, 1, 1, 1, .6, .6, .6, 4, 1.5, .5, .3, .2, .2])
    probs /= probs.sum()
    for it in range(400):
        n = int(rng.integers(0, 400))
        data = rng.choice(alphabet, size=n, p=probs).tobytes()
        if it % 5 == 0:
            data = b">" + data
        if it % 7 == 0:
            data = data.replace(b"+", b"+\n")
        path = tmp_path / f"f{it}.txt"
        path.write_bytes(data)
        seqs, _ = oracle.parse_reads(data)
        # serial parse, then the speculative multi-range parse with ranges of a few bytes (every guess at a record
        # boundary that can go wrong does go wrong somewhere in these files)
        for threads, chunk in ((1, None), (4, 24), (8, 5), (3, 1)):
            if chunk is None:
                os.environ.pop("LRB_PARSE_CHUNK", None)
            else:
                os.environ["LRB_PARSE_CHUNK"] = str(chunk)
            try:
                pr = PackedReads.from_file(str(path), threads=threads)
            finally:
                os.environ.pop("LRB_PARSE_CHUNK", None)
            assert pr.n_reads == len(seqs), (it, threads, chunk, data)
            assert list(pr.read_len) == [len(s) for s in seqs], (it, threads, chunk, data)
            for i, s in enumerate(seqs):
                assert pr.unpack(i) == bytes(b"ACTG"[(c >> 1) & 3] for c in s), (it, i, threads, chunk)


def test_parallel_parse_of_structured_files(tmp_path):
    """Well-formed-looking FASTA/FASTQ (multi-line records, CRLF, quality lines that begin with '@', '>' or '+',
    empty records) cut into ranges at every granularity: the multi-range parse equals the oracle reader."""
    rng = np.random.default_rng(12)
    qual_alpha = np.frombuffer(b"@>+!I5#", dtype=np.uint8)
    base_alpha = np.frombuffer(b"ACGTNacgt", dtype=np.uint8)
    for it in range(40):
        fastq, crlf, wrap = bool(it % 2), bool(it % 3 == 0), int(rng.integers(0, 3))
        eol = b"\r\n" if crlf else b"\n"
        out = []
        for r in range(int(rng.integers(1, 60))):
            n = int(rng.integers(0, 200))
            seq = rng.choice(base_alpha, size=n).tobytes()
            if fastq:
                q = rng.choice(qual_alpha, size=n).tobytes()
                out += [b"@r%d some comment" % r, seq, b"+", q]
            else:
                w = (0, 60, 7)[wrap]
                lines = [seq[i:i + w] for i in range(0, n, w)] if w and n else [seq]
                out += [b">r%d" % r] + lines
        data = eol.join(out) + (eol if it % 4 else b"")
        path = tmp_path / f"s{it}.txt"
        path.write_bytes(data)
        seqs, _ = oracle.parse_reads(data)
        for threads, chunk in ((1, None), (8, 64), (5, 9), (16, 1)):
            if chunk is not None:
                os.environ["LRB_PARSE_CHUNK"] = str(chunk)
            try:
                pr = PackedReads.from_file(str(path), threads=threads)
            finally:
                os.environ.pop("LRB_PARSE_CHUNK", None)
            assert list(pr.read_len) == [len(x) for x in seqs], (it, threads, chunk)
            for i, x in enumerate(seqs):
                assert pr.unpack(i) == bytes(b"ACTG"[(c >> 1) & 3] for c in x), (it, i, threads, chunk)


def _load_gz_streamed(path, chunk, threads=3, parse_chunk=None):
    env = {"LRB_GZ_CHUNK": str(chunk)}
    if parse_chunk is not None:
        env["LRB_PARSE_CHUNK"] = str(parse_chunk)
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        return PackedReads.from_file(str(path), threads=threads)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.mark.parametrize("name", golden_inputs())
def test_streamed_gzip_ingest_equals_the_whole_image_parse(name, tmp_path):
    """gzip input is inflated chunk by chunk while the previous chunk is parsed; a record cut by the end of a chunk is parsed
    again with more data.  Chunks of 1 .. 4096 bytes put a chunk boundary at every byte of every fixture."""
    src = os.path.join(GOLDEN, name)
    raw = _gz(src) if name.endswith(".gz") else open(src, "rb").read()
    gz = tmp_path / "in.gz"
    with gzip.open(gz, "wb") as f:
        f.write(raw)
    seqs, _ = oracle.parse_reads(raw)
    for chunk, threads, pchunk in ((1, 1, None), (3, 2, 2), (64, 4, 16), (1000, 3, 100), (4096, 8, None), (1 << 28, 4, None)):
        pr = _load_gz_streamed(gz, chunk, threads, pchunk)
        assert pr.n_reads == len(seqs) and list(pr.read_len) == [len(x) for x in seqs], (name, chunk)
        _check_packing(pr, seqs)
        blk, word = pr.exceptions()
        want = _default_valid(pr)[:pr.n_blocks]
        if len(blk):
            want[blk] = word
        assert np.array_equal(want, pr.valid[:pr.n_blocks]), (name, chunk)


def test_streamed_gzip_ingest_fuzz(tmp_path):
    """random byte soup (headers, quality-like lines, CR, NUL-free) through the streamed reader with tiny chunks"""
    rng = np.random.default_rng(77)
    alphabet = np.frombuffer(b"ACGTacgtN>@+\n\r \t;-", dtype=np.uint8)
    probs = np.array([8, 8, 8, 8, 1, 1, 1, 1, 1, .6, .6, .6, 4, 1.5, .5, .3, .2, .2])
    probs /= probs.sum()
    for it in range(150):
        n = int(rng.integers(0, 600))
        data = rng.choice(alphabet, size=n, p=probs).tobytes()
        if it % 5 == 0:
            data = b">" + data
        if it % 7 == 0:
            data = data.replace(b"+", b"+\n")
        if it % 3 == 0:      # well-formed FASTQ records in the soup, so that complete quality blocks meet chunk boundaries
            s = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=int(rng.integers(1, 40))).tobytes()
            data = b"@r\n" + s + b"\n+\n" + b"@" * len(s) + b"\n" + data
        path = tmp_path / f"z{it}.gz"
        with gzip.open(path, "wb") as f:
            f.write(data)
        seqs, _ = oracle.parse_reads(data)
        for chunk, threads, pchunk in ((1, 1, None), (5, 4, 3), (37, 3, 8)):
            pr = _load_gz_streamed(path, chunk, threads, pchunk)
            assert list(pr.read_len) == [len(x) for x in seqs], (it, chunk, data)
            for i, x in enumerate(seqs):
                assert pr.unpack(i) == bytes(b"ACTG"[(c >> 1) & 3] for c in x), (it, i, chunk)


def test_ingest_nul_byte_and_missing_file(tmp_path):
    p = tmp_path / "nul.fa"
    p.write_bytes(b">a\nACGT\0ACGTACGT\n>b\nGGGG\n")
    seqs, _ = oracle.load_reads(str(p))
    assert seqs == [b"ACGT", b"GGGG"]
    pr = PackedReads.from_file(str(p))
    assert list(pr.read_len) == [4, 4]
    pr = PackedReads.from_file(str(tmp_path / "does_not_exist.fa"))
    assert pr.n_reads == 0 and pr.n_blocks == 0 and pr.n_tiles == 0


def test_from_sequences_and_lengths_layout():
    rng = np.random.default_rng(3)
    lens = [0, 1, 31, 32, 33, 8191, 8192, 8193, 20000, 0, 5]
    seqs = [rng.choice(np.frombuffer(b"ACGTN", dtype=np.uint8), size=n).tobytes() for n in lens]
    pr = PackedReads.from_sequences(seqs, threads=4)
    _check_packing(pr, seqs)
    pl = PackedReads.from_lengths(np.array(lens, dtype=np.uint32))
    assert np.array_equal(pl.read_blk, pr.read_blk) and np.array_equal(pl.tile_blk, pr.tile_blk)
    assert np.array_equal(pl.tile_read, pr.tile_read) and not pl.codes.any()


def test_reads_slice_is_a_read_set_of_its_own():
    """lrb_reads_slice (device shards / batches of the host pipeline): same reads, indices rebased, exceptions kept."""
    rng = np.random.default_rng(5)
    seqs = []
    for i in range(300):
        n = int(rng.choice([0, 1, 14, 15, 31, 32, 33, 500, 9000]))
        s = bytearray(rng.choice(list(b"ACGT"), n).astype(np.uint8).tobytes())
        if n > 40 and i % 7 == 0:
            s[n // 2] = ord("N")
        if n > 40 and i % 11 == 0:
            s = bytearray(bytes(s).lower())
        seqs.append(bytes(s))
    pr = PackedReads.from_sequences(seqs, threads=3)
    for lo, hi in [(0, 300), (0, 1), (17, 18), (5, 120), (120, 300), (299, 300), (40, 40)]:
        sl = pr.slice(lo, hi)
        assert sl.n_reads == hi - lo and sl.total_bases == sum(len(s) for s in seqs[lo:hi])
        assert sl.n_blocks == sum(len(s) // 32 + 1 for s in seqs[lo:hi])
        _check_packing(sl, seqs[lo:hi])
        # the slice's exception list rebuilds its validity words exactly
        blk, word = sl.exceptions()
        want = _default_valid(sl)[:sl.n_blocks]
        if len(blk):
            want[blk] = word
        assert np.array_equal(want, sl.valid[:sl.n_blocks])
        # tiles: <= 256 blocks of one read each, covering every block once, in order
        covered = 0
        for t in range(sl.n_tiles):
            r, b = int(sl.tile_read[t]), int(sl.tile_blk[t])
            assert b == covered and int(sl.read_blk[r]) <= b < int(sl.read_blk[r + 1])
            covered = min(b + 256, int(sl.read_blk[r + 1]))
        assert covered == sl.n_blocks


def _default_valid(pr):
    blk, ln = np.array(pr.read_blk, dtype=np.int64), np.array(pr.read_len, dtype=np.int64)
    want = np.zeros(pr.n_blocks + 1, dtype=np.uint32)
    for r in range(pr.n_reads):
        for b in range(blk[r], blk[r + 1]):
            n_in = min(32, max(0, ln[r] - 32 * (b - blk[r])))
            want[b] = 0xFFFFFFFF if n_in == 32 else (1 << n_in) - 1
    return want


def test_validity_exceptions_index():
    """The exception list (what lrb_profile_host ships instead of the bitmap) + the length-implied words == valid."""
    rng = np.random.default_rng(5)
    lens = [0, 1, 31, 32, 33, 64, 700, 8193, 3000, 0, 5, 2500]
    alphabet = np.frombuffer(b"ACGT", dtype=np.uint8)
    seqs = [bytearray(rng.choice(alphabet, size=n).tobytes()) for n in lens]
    seqs[6][10] = ord("N"); seqs[6][699] = ord("a"); seqs[7][8192] = ord("n"); seqs[8][0] = ord("R")
    seqs[11] = bytearray(bytes(seqs[11]).lower())
    for threads in (1, 4):
        pr = PackedReads.from_sequences([bytes(s) for s in seqs], threads=threads)
        blk, word = pr.exceptions()
        want = _default_valid(pr)
        assert np.all(np.diff(blk.astype(np.int64)) > 0)
        assert np.all(want[blk] != word)
        want[blk] = word
        assert np.array_equal(want, pr.valid)
        assert len(blk) == 1 + 1 + 1 + 1 + (2500 // 32 + 1)     # N, a, n, R and every block of the lowercase read
    # a hand-filled layout has no list until it is indexed
    pl = PackedReads.from_lengths(np.array(lens, dtype=np.uint32))
    assert len(pl.exceptions()[0]) == 0
    pl.codes[:] = pr.codes
    pl.valid[:] = pr.valid
    assert pl.index_valid(threads=3) == len(blk)
    b2, w2 = pl.exceptions()
    assert np.array_equal(b2, blk) and np.array_equal(w2, word)


# ---- exact "%f" ------------------------------------------------------------------------------------

def test_fixed6_matches_printf():
    rng = np.random.default_rng(11)
    pairs = [(0, 0), (0, 5), (1, 1), (5, 5), (1, 3), (2, 3), (1, 8), (1, 16), (1, 64), (3, 64), (1, 2 ** 20), (1, 2 ** 31),
             (1, 4294967295), (4294967295, 4294967295), (1, 10000), (1, 10001), (1, 9999), (5, 49999), (5, 50000), (5, 50001),
             (1, 2000000), (3, 2000000), (1, 1000000), (15, 32), (1, 32), (9, 16)]
    dens = rng.integers(1, 200000, size=60000)
    nums = (rng.random(60000) * (dens + 1)).astype(np.int64).clip(0, dens)
    pairs += list(zip(nums.tolist(), dens.tolist()))
    big = rng.integers(1, 2 ** 32, size=20000)
    pairs += [(int(rng.integers(0, d + 1)), int(d)) for d in big]
    # ties: values with exactly 7 significant decimals ending in 5 can only tie if exactly representable
    pairs += [(n, 2 ** s) for s in range(1, 25) for n in (1, 3, 5, 7, 2 ** s - 1) if n <= 2 ** s]
    for num, den in pairs:
        v = float(num) / float(max(den, 1))
        want = int(round(float("%f" % v) * 1e6))
        assert _lib.lib.lrb_fixed6(num, den, 0) == want, (num, den)
        vc = 0.0 if v < 1e-4 else v
        assert _lib.lib.lrb_fixed6(num, den, 1) == int(round(float("%f" % vc) * 1e6)), (num, den)
        assert oracle.format_f(v) == "%f" % v


# ---- kernels' lane arithmetic, emulated on the CPU, vs oracle and golden files ---------------------------

def test_kmer_lut_matches_oracle():
    for k in (3, 4, 5):
        lut, width = kmer_lut(k)
        olut, owidth = oracle.kmer_lut(k)
        assert width == owidth == COMP_WIDTH[k] and np.array_equal(lut, olut)


def test_bucket_rule_and_revcomp(emul):
    rng = np.random.default_rng(5)
    for S in (1, 2, 3, 7, 10, 32, 33, 1000, 65536, 2 ** 31 - 1, 2 ** 32 - 1, 2 ** 33, 2 ** 40):
        for B in (1, 2, 3, 8, 10, 32, 4096):
            cs = [0, 1, 2, 3, S - 1, S, S + 1, 2 * S - 1, 2 * S, 2 * S + 1, 3 * S, B * S - 1, B * S, (B + 1) * S - 1, (B + 1) * S,
                  (B + 1) * S + 1, 2 ** 32 - 1] + rng.integers(0, 2 ** 32, size=40).tolist() + rng.integers(0, 50 * S + 50, size=40).tolist()
            for c in cs:
                if 0 <= c < 2 ** 32:
                    assert emul.emul_coverage_bin(c, S, B) == oracle.bucket(c, S, B), (c, S, B)
    for x in [0, 1, 2 ** 30 - 1, 0x8000, 0x12345678 & (2 ** 30 - 1)] + rng.integers(0, 2 ** 30, size=2000).tolist():
        assert emul.emul_revcomp15(x) == oracle.revcomp(x, 15)


def test_mirror_index_algebra(emul):
    seen = np.zeros(2 ** 27, dtype=np.uint8)
    assert emul.emul_mirror_check(_p(seen)) == 0
    # every bit-15-set index was a destination exactly once, no bit-15-clear index ever was
    bits = np.unpackbits(seen[:2 ** 16], bitorder="little")
    idx = np.arange(bits.size)
    assert np.array_equal(bits.astype(bool), (idx & 0x8000) != 0)
    assert int(np.unpackbits(seen).sum()) == 2 ** 29


def _emulated_profile(emul, pr, seqs, k_list, cov_params):
    n = pr.n_reads
    view = C.byref(pr.view)
    comps = {}
    for k in k_list:
        lut, P = kmer_lut(k)
        out = np.zeros((n, P), dtype=np.uint32)
        assert emul.emul_composition(view, k, _p(lut), _p(out)) == 0
        comps[k] = out
    table = np.zeros(2 ** 30, dtype=np.uint32)
    emul.emul_count(view, _p(table), C.c_uint32(0), C.c_uint32(2 ** 30))
    keys = np.flatnonzero(table).astype(np.uint32)
    assert emul.emul_mirror_keys(_p(table), _p(keys), C.c_uint64(len(keys))) == 0
    covs = {}
    for bs, bc in cov_params:
        hist = np.zeros((n, bc), dtype=np.uint32)
        sums = np.zeros(n, dtype=np.uint32)
        emul.emul_search(view, _p(table), C.c_long(bs), bc, _p(hist), _p(sums), C.c_uint32(0), C.c_uint32(2 ** 30))
        covs[(bs, bc)] = (hist, sums)
    return comps, table, covs


@pytest.mark.parametrize("name", golden_inputs())
def test_emulated_kernels_reproduce_reference_files(emul, name, tmp_path):
    stem = name.split(".")[0]
    src = os.path.join(GOLDEN, name)
    pr = PackedReads.from_file(src, threads=2)
    seqs, _ = oracle.load_reads(src)
    comps, table, covs = _emulated_profile(emul, pr, seqs, (3, 4, 5), COV_PARAMS)
    n = pr.n_reads
    rl = np.array(pr.read_len, dtype=np.uint32)
    for k, out in comps.items():
        # raw counts against the oracle, text against the reference tool's file
        for i, s in enumerate(seqs):
            raw, total, _ = oracle.composition(s, k)
            assert np.array_equal(out[i], raw.astype(np.uint32)), (name, k, i)
        path = str(tmp_path / f"com{k}")
        assert _lib.lib.lrb_write_composition_txt(path.encode(), _p(out), _p(rl), n, k, 2) == 0
        assert open(path, "rb").read() == _gz(os.path.join(GOLDEN, f"{stem}.com_k{k}.txt.gz")), (name, k)
    gold = np.load(os.path.join(GOLDEN, f"{stem}.table.npz"))
    keys = gold["keys"].astype(np.int64)
    assert np.array_equal(table[keys], gold["counts"])
    assert int(np.count_nonzero(table)) == len(keys)
    for (bs, bc), (hist, sums) in covs.items():
        path = str(tmp_path / "cov")
        assert _lib.lib.lrb_write_coverage_txt(path.encode(), _p(hist), _p(sums), n, bc, 2) == 0
        assert open(path, "rb").read() == _gz(os.path.join(GOLDEN, f"{stem}.cov_bs{bs}_bc{bc}.txt.gz")), (name, bs, bc)
    del table


def test_emulated_key_sharded_count_and_search_sum_to_whole(emul):
    src = os.path.join(GOLDEN, "g06_community.fa")
    pr = PackedReads.from_file(src)
    view = C.byref(pr.view)
    n = pr.n_reads
    full = np.zeros(2 ** 30, dtype=np.uint32)
    emul.emul_count(view, _p(full), C.c_uint32(0), C.c_uint32(2 ** 30))
    parts = np.zeros(2 ** 30, dtype=np.uint32)
    bounds = [0, 2 ** 28, 2 ** 29, 3 * 2 ** 28, 2 ** 30]
    hist_sum = np.zeros((n, 10), dtype=np.uint32)
    sums_sum = np.zeros(n, dtype=np.uint32)
    for lo, hi in zip(bounds[:-1], bounds[1:]):
        emul.emul_count(view, _p(parts), C.c_uint32(lo), C.c_uint32(hi))
    nz = np.flatnonzero(full)
    assert np.array_equal(np.flatnonzero(parts), nz) and np.array_equal(parts[nz], full[nz])
    for lo, hi in zip(bounds[:-1], bounds[1:]):
        h = np.zeros((n, 10), dtype=np.uint32)
        s = np.zeros(n, dtype=np.uint32)
        emul.emul_search(view, _p(parts), C.c_long(32), 10, _p(h), _p(s), C.c_uint32(lo), C.c_uint32(hi))
        hist_sum += h
        sums_sum += s
    keys = nz.astype(np.uint32)
    emul.emul_mirror_keys(_p(full), _p(keys), C.c_uint64(len(keys)))
    h = np.zeros((n, 10), dtype=np.uint32)
    s = np.zeros(n, dtype=np.uint32)
    emul.emul_search(view, _p(full), C.c_long(32), 10, _p(h), _p(s), C.c_uint32(0), C.c_uint32(2 ** 30))
    assert np.array_equal(h, hist_sum) and np.array_equal(s, sums_sum)


def test_npy_writers_equal_float_of_text(tmp_path):
    rng = np.random.default_rng(9)
    n, k, bins = 300, 4, 10
    P = COMP_WIDTH[k]
    rl = rng.integers(0, 3000, size=n).astype(np.uint32)
    comp = np.zeros((n, P), dtype=np.uint32)
    for i in range(n):
        tot = max(0, int(rl[i]) - k + 1)
        if tot:
            comp[i] = rng.multinomial(tot, np.ones(P) / P)
    sums = rng.integers(0, 3000, size=n).astype(np.uint32)
    hist = np.zeros((n, bins), dtype=np.uint32)
    for i in range(n):
        if sums[i]:
            hist[i] = rng.multinomial(int(sums[i]), rng.dirichlet(np.ones(bins) * 0.3))
    for kind, args, width in (("composition", (_p(comp), _p(rl), n, k, 2), P), ("coverage", (_p(hist), _p(sums), n, bins, 2), bins)):
        txt, npy = str(tmp_path / f"{kind}.txt"), str(tmp_path / f"{kind}.npy")
        assert getattr(_lib.lib, f"lrb_write_{kind}_txt")(txt.encode(), *args) == 0
        assert getattr(_lib.lib, f"lrb_write_{kind}_npy")(npy.encode(), *args) == 0
        # what pipelines.py:313-324 does with the text file
        want = np.array([np.array(list(map(float, line.strip().split()))) for line in open(txt) if len(line.strip()) > 0])
        got = np.load(npy)
        assert got.dtype == np.float64 and got.shape == (n, width) and np.array_equal(got, want)


def test_writers_are_thread_count_invariant_and_match_printf(tmp_path):
    """Multi-threaded writers (blocks written at their final offset): any thread count gives the same bytes, and the
    bytes are C's "%f" of the double quotient (count-kmers.cpp:110-118, search-15mers.cpp:35-48)."""
    rng = np.random.default_rng(9)
    n, k, P, bins = 5000, 4, 136, 10
    rl = rng.integers(0, 30000, size=n).astype(np.uint32)
    comp = np.zeros((n, P), dtype=np.uint32)
    for i in range(n):
        tot = max(int(rl[i]) - k + 1, 0)
        if tot:
            comp[i] = rng.multinomial(tot, rng.dirichlet(np.ones(P) * 0.2))
    sums = rng.integers(0, 30000, size=n).astype(np.uint32)
    hist = np.zeros((n, bins), dtype=np.uint32)
    for i in range(n):
        if sums[i]:
            hist[i] = rng.multinomial(int(sums[i]), rng.dirichlet(np.ones(bins) * 0.3))
    for kind, args, in (("composition", (_p(comp), _p(rl), n, k)), ("coverage", (_p(hist), _p(sums), n, bins))):
        blobs = {}
        for ext in ("txt", "npy"):
            for th in (1, 3, 8):
                path = str(tmp_path / f"{kind}.{th}.{ext}")
                assert getattr(_lib.lib, f"lrb_write_{kind}_{ext}")(path.encode(), *args, th) == 0
                blobs[(ext, th)] = open(path, "rb").read()
            assert blobs[(ext, 1)] == blobs[(ext, 3)] == blobs[(ext, 8)], (kind, ext)
        lines = blobs[("txt", 1)].split(b"\n")
        assert len(lines) == n + 1 and lines[-1] == b""
        for i in rng.choice(n, size=200, replace=False):
            if kind == "composition":
                tot = max(1.0, float(max(int(rl[i]) - k + 1, 0)))
                want = b"".join(b"%f " % (float(c) / tot) for c in comp[i])
            else:
                vals = [float(c) / float(sums[i]) if sums[i] else 0.0 for c in hist[i]]
                want = b" ".join(b"%f" % (0.0 if v < 1e-4 else v) for v in vals)
            assert lines[i] == want, (kind, i)


def test_table_file_roundtrip_format(tmp_path):
    t = np.zeros(2 ** 30, dtype=np.uint32)
    t[[0, 5, 2 ** 30 - 1]] = [7, 9, 11]
    path = str(tmp_path / "tbl")
    assert _lib.lib.lrb_table_write_file(path.encode(), _p(t)) == 0
    assert os.path.getsize(path) == 8 + 4 * 2 ** 30      # 4 294 967 304 bytes
    assert int(np.fromfile(path, dtype=np.uint64, count=1)[0]) == 2 ** 30
    back = np.empty(2 ** 30, dtype=np.uint32)
    assert _lib.lib.lrb_table_read_file(path.encode(), _p(back)) == 0
    assert back[0] == 7 and back[5] == 9 and back[-1] == 11 and int(np.count_nonzero(back)) == 3
    os.remove(path)
    (tmp_path / "bad").write_bytes(b"\x01\x00\x00\x00\x00\x00\x00\x00abcd")
    assert _lib.lib.lrb_table_read_file(str(tmp_path / "bad").encode(), _p(back)) == _lib.LRB_EFORMAT
