"""Synthetic long-read sets for the benchmark and the tests (input generation only).

Follows SURVEY.md section 8(d): a community of 64 random genomes (1-6 Mbp, log-normal abundances),
reads drawn with probability ~ abundance x length, uniform start, random strand, ONT-like
(4 % sub / 3 % ins / 3 % del) or HiFi-like (0.5 %) errors.  Read METADATA (genome, start, strand,
length) comes from numpy.random.default_rng(seed) here; the BASES are expanded from it by the
integer-only generator in csrc/synth_core.h, identically on the host (lrb_synth_host, fed to the oracle
in tests) and on the device (lrb_dev_synth, straight into the packed stream for the large configs).
"""
import ctypes as C

import numpy as np

from ._lib import SynthParams, check, lib
from .profile import PackedReads, _ptr

# BASELINE.json configs: name -> (reads, length model, error model, k, seed)
CONFIGS = {
    "cfg1_100k_5kb_k3": dict(n_reads=100_000, lengths="gamma5k", errors="ont", k=3, seed=11),
    "cfg2_1M_5kb_ont_k4": dict(n_reads=1_000_000, lengths="gamma5k", errors="ont", k=4, seed=22),
    "cfg3_2M_5kb_k3": dict(n_reads=2_000_000, lengths="gamma5k", errors="ont", k=3, seed=33),
    "cfg4_500k_15kb_hifi_k5": dict(n_reads=500_000, lengths="hifi15k", errors="hifi", k=5, seed=44),
    "cfg5_longtail_k3": dict(n_reads=1_100_000, lengths="longtail", errors="ont", k=3, seed=55),
}

ERRORS = {"ont": (0.04, 0.03, 0.03), "hifi": (0.002, 0.0015, 0.0015), "none": (0.0, 0.0, 0.0)}


def draw_lengths(rng, n, model):
    if model == "gamma5k":
        x = rng.gamma(2.0, 2500.0, size=n)
        return np.clip(x, 500, 60000).astype(np.uint32)
    if model == "hifi15k":
        return np.clip(rng.normal(15000.0, 3000.0, size=n), 1000, None).astype(np.uint32)
    if model == "longtail":
        return np.clip(rng.lognormal(np.log(6000.0), 1.0, size=n), 1000, 100000).astype(np.uint32)
    raise ValueError(model)


class SynthSpec:
    """Everything needed to expand a read set: params struct, genome lengths, per-read metadata, lengths."""

    def __init__(self, n_reads, lengths="gamma5k", errors="ont", seed=1, n_genomes=64, n_rate=0.0, lowercase_frac=0.0,
                 edge_lengths=False, scale=1.0, shard=0, read_base=0):
        rng = np.random.default_rng(seed)
        self.glen = rng.integers(int(1_000_000 * scale), int(6_000_000 * scale) + 1, size=n_genomes).astype(np.uint32)
        abundance = rng.lognormal(0.0, 1.0, size=n_genomes)
        if shard:   # same community (genomes, abundances), an independent draw of reads
            rng = np.random.default_rng([seed, shard])
        w = abundance * self.glen
        genome = rng.choice(n_genomes, size=n_reads, p=w / w.sum()).astype(np.uint32)
        self.lengths = draw_lengths(rng, n_reads, lengths) if isinstance(lengths, str) else np.asarray(lengths, dtype=np.uint32)
        if edge_lengths and n_reads >= 64:   # empty / shorter-than-k / shorter-than-15 / block-boundary reads
            edge = np.array([0, 1, 2, 3, 4, 5, 13, 14, 15, 16, 31, 32, 33, 63, 64, 65, 8191, 8192, 8193, 16384], dtype=np.uint32)
            pos = rng.choice(n_reads, size=len(edge), replace=False)
            self.lengths[pos] = edge
        start = (rng.random(n_reads) * self.glen[genome]).astype(np.uint32)
        flags = (rng.random(n_reads) < 0.5).astype(np.uint32)
        if lowercase_frac > 0:
            flags |= ((rng.random(n_reads) < lowercase_frac).astype(np.uint32) << 1)
        self.meta = np.zeros((n_reads, 4), dtype=np.uint32)
        self.meta[:, 0], self.meta[:, 1], self.meta[:, 2] = genome, start, flags
        sub, ins, dele = ERRORS[errors] if isinstance(errors, str) else errors
        f = lambda p: int(round(p * 2 ** 32))
        self.params = SynthParams(seed=int(seed) * 0x9E3779B1 + 12345, n_genomes=n_genomes, sub_thr=f(sub), ins_thr=f(ins),
                                  del_thr=f(dele), n_thr=f(n_rate), read_base=int(read_base))
        self.n_reads = n_reads
        self.total_bases = int(self.lengths.astype(np.uint64).sum())

    def host_ascii(self):
        """-> (bases u8[L], offsets u64[N+1]) generated on the CPU (tests: the oracle's input)."""
        offsets = np.zeros(self.n_reads + 1, dtype=np.uint64)
        offsets[1:] = np.cumsum(self.lengths.astype(np.uint64))
        bases = np.zeros(int(offsets[-1]) + 1, dtype=np.uint8)
        check(lib.lrb_synth_host(C.byref(self.params), _ptr(self.glen), _ptr(self.meta), _ptr(self.lengths), self.n_reads,
                                 _ptr(offsets), _ptr(bases)))
        return bases, offsets

    def host_sequences(self):
        bases, offsets = self.host_ascii()
        raw = bases.tobytes()
        return [raw[int(offsets[i]):int(offsets[i + 1])] for i in range(self.n_reads)]

    def host_packed(self, threads=8):
        bases, offsets = self.host_ascii()
        return PackedReads.from_ascii(bases, offsets, threads)

    def device_reads(self, device):
        """Generate straight into HBM: returns (DeviceReads, layout PackedReads with empty host codes)."""
        import torch
        from .profile import DeviceReads
        layout = PackedReads.from_lengths(self.lengths)
        dr = DeviceReads(layout, device, upload=False)
        glen = torch.from_numpy(self.glen.view(np.int32)).to(dr.device)
        meta = torch.from_numpy(self.meta.view(np.int32).reshape(-1)).to(dr.device)
        with torch.cuda.device(dr.device):
            check(lib.lrb_dev_synth(C.byref(dr.view), C.byref(self.params), C.c_void_p(glen.data_ptr()),
                                    C.c_void_p(meta.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
            torch.cuda.synchronize()
        return dr, layout


def write_fasta(path, seqs, width=0):
    with open(path, "wb") as f:
        for i, s in enumerate(seqs):
            f.write(b">r%d\n" % i)
            if width and len(s) > width:
                for j in range(0, len(s), width):
                    f.write(s[j:j + width] + b"\n")
            else:
                f.write(s + b"\n")
