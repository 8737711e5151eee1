"""ctypes view of oracle/_build/liboracle.so (the plain-C restatement in oracle/lrb_oracle.c).

TEST INFRASTRUCTURE ONLY. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module, and only as the checker. The product package
(lrbinner_b200/) never imports it.

Parity status: PINNED against the unmodified reference tools (oracle/_ref, built by oracle/Makefile
from /root/reference) — see tests/test_oracle_pins.py and tests/golden/.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
REF_DIR = os.path.join(_HERE, "_ref")
TABLE_SIZE = 1 << 30


def build(force=False):
    """Compile liboracle.so (and oracle/_ref when /root/reference is mounted)."""
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_HERE, "lrb_oracle.c")):
        subprocess.check_call(["make", "-C", _HERE, "_build/liboracle.so"], stdout=subprocess.DEVNULL)
    if os.path.isdir("/root/reference/mbcclr_utils") and (force or not ref_available()):
        subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)


def ref_available():
    return all(os.access(os.path.join(REF_DIR, t), os.X_OK) for t in ("count-kmers", "count-15mers", "search-15mers"))


class _Reads(C.Structure):
    _fields_ = [("n", C.c_size_t), ("seq", C.POINTER(C.c_char_p)), ("len", C.POINTER(C.c_size_t)),
                ("name", C.POINTER(C.c_char_p))]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.orc_revcomp.restype = C.c_uint64
        L.orc_revcomp.argtypes = [C.c_uint64, C.c_int]
        L.orc_kmer_lut.restype = C.c_int
        L.orc_kmer_lut.argtypes = [C.c_int, C.c_void_p]
        L.orc_composition.restype = None
        L.orc_composition.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_count_15mers.restype = None
        L.orc_count_15mers.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p]
        L.orc_coverage.restype = None
        L.orc_coverage.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p, C.c_long, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_count_kmers_k.restype = None
        L.orc_count_kmers_k.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_void_p]
        L.orc_coverage_k.restype = None
        L.orc_coverage_k.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_void_p, C.c_long, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_window_keys.restype = C.c_size_t
        L.orc_window_keys.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_void_p]
        L.orc_bucket.restype = C.c_int
        L.orc_bucket.argtypes = [C.c_uint32, C.c_long, C.c_int]
        L.orc_reads_load.restype = C.c_int
        L.orc_reads_load.argtypes = [C.c_char_p, C.POINTER(_Reads)]
        L.orc_reads_parse.restype = C.c_int
        L.orc_reads_parse.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(_Reads)]
        L.orc_reads_free.restype = None
        L.orc_reads_free.argtypes = [C.POINTER(_Reads)]
        L.orc_count_kmers_file.restype = C.c_int
        L.orc_count_kmers_file.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
        L.orc_count_15mers_file.restype = C.c_int
        L.orc_count_15mers_file.argtypes = [C.c_char_p, C.c_char_p]
        L.orc_search_15mers_file.restype = C.c_int
        L.orc_search_15mers_file.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_long, C.c_int]
        L.orc_table_alloc.restype = C.c_void_p
        L.orc_table_free.argtypes = [C.c_void_p]
        L.orc_table_write.restype = C.c_int
        L.orc_table_write.argtypes = [C.c_char_p, C.c_void_p, C.c_uint64]
        L.orc_format_f.restype = C.c_int
        L.orc_format_f.argtypes = [C.c_double, C.c_char_p]
        _lib = L
    return _lib


def revcomp(x, k):
    return int(lib().orc_revcomp(x, k))


def kmer_lut(k):
    lut = np.zeros(4 ** k, dtype=np.uint32)
    width = lib().orc_kmer_lut(k, lut.ctypes.data)
    return lut, width


def _as_bytes(seq):
    return seq if isinstance(seq, (bytes, bytearray)) else seq.encode("latin-1")


def composition(seq, k):
    """-> (raw u64[width], total, profile f64[width]) for one read (count-kmers.cpp:66-95)."""
    lut, width = kmer_lut(k)
    s = _as_bytes(seq)
    raw = np.zeros(width, dtype=np.uint64)
    prof = np.zeros(width, dtype=np.float64)
    total = C.c_uint64(0)
    lib().orc_composition(s, len(s), k, lut.ctypes.data, width, raw.ctypes.data, C.byref(total), prof.ctypes.data)
    return raw, total.value, prof


class Table:
    """4^15-entry u32 table (lazy pages), kmer_utils.h:114-156 semantics."""

    def __init__(self):
        self._p = lib().orc_table_alloc()
        if not self._p:
            raise MemoryError("oracle table")
        self.array = np.ctypeslib.as_array((C.c_uint32 * TABLE_SIZE).from_address(self._p))

    def count(self, seq):
        s = _as_bytes(seq)
        lib().orc_count_15mers(s, len(s), self._p)

    def coverage(self, seq, bin_size, bins):
        """-> (raw u64[bins], sum, vec f64[bins]) (kmer_utils.h:24-87)."""
        s = _as_bytes(seq)
        raw = np.zeros(bins, dtype=np.uint64)
        vec = np.zeros(bins, dtype=np.float64)
        total = C.c_uint64(0)
        lib().orc_coverage(s, len(s), self._p, bin_size, bins, raw.ctypes.data, C.byref(total), vec.ctypes.data)
        return raw, total.value, vec

    def write(self, path):
        return lib().orc_table_write(path.encode(), self._p, TABLE_SIZE)

    def close(self):
        if self._p:
            self.array = None
            lib().orc_table_free(self._p)
            self._p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class SmallTable:
    """4^k-entry table for odd k < 15 (tests of the multi-GPU plumbing only; the reference fixes k = 15)."""

    def __init__(self, k):
        self.k = k
        self.array = np.zeros(4 ** k, dtype=np.uint32)

    def count(self, seq):
        s = _as_bytes(seq)
        lib().orc_count_kmers_k(s, len(s), self.k, self.array.ctypes.data)

    def coverage(self, seq, bin_size, bins):
        s = _as_bytes(seq)
        raw = np.zeros(bins, dtype=np.uint64)
        vec = np.zeros(bins, dtype=np.float64)
        total = C.c_uint64(0)
        lib().orc_coverage_k(s, len(s), self.k, self.array.ctypes.data, bin_size, bins, raw.ctypes.data, C.byref(total),
                             vec.ctypes.data)
        return raw, total.value, vec


def window_keys(seq, k=15):
    """Forward table index of every valid k-mer window of the read, in order (kmer_utils.h:36-72)."""
    s = _as_bytes(seq)
    out = np.zeros(max(len(s), 1), dtype=np.uint32)
    n = lib().orc_window_keys(s, len(s), k, out.ctypes.data)
    return out[:n]


def bucket(count, bin_size, bins):
    return lib().orc_bucket(count, bin_size, bins)


def _reads_out(r):
    seqs = [C.string_at(r.seq[i], r.len[i]) for i in range(r.n)]
    names = [r.name[i] for i in range(r.n)]
    lib().orc_reads_free(C.byref(r))
    return seqs, names


def load_reads(path):
    """FASTA/FASTQ(+gz) -> (list of sequences as bytes, list of names) with kseq/io_utils semantics."""
    r = _Reads()
    lib().orc_reads_load(path.encode(), C.byref(r))
    return _reads_out(r)


def parse_reads(data):
    r = _Reads()
    lib().orc_reads_parse(data, len(data), C.byref(r))
    return _reads_out(r)


def count_kmers_file(reads, out_txt, k):
    return lib().orc_count_kmers_file(reads.encode(), out_txt.encode(), k)


def count_15mers_file(reads, out_table):
    return lib().orc_count_15mers_file(reads.encode(), out_table.encode())


def search_15mers_file(table, reads, out_txt, bin_size, bins):
    return lib().orc_search_15mers_file(table.encode(), reads.encode(), out_txt.encode(), bin_size, bins)


def format_f(v):
    buf = C.create_string_buffer(40)
    lib().orc_format_f(v, buf)
    return buf.value.decode()


# ---- the unmodified reference tools (oracle/_ref), argv contracts from count-kmers.cpp:195-198,
# ---- count-15mers.cpp:101-103, search-15mers.cpp:124-136

def ref_count_kmers(reads, out_txt, k, threads=1):
    subprocess.check_call([os.path.join(REF_DIR, "count-kmers"), reads, out_txt, str(k), str(threads)],
                          stdout=subprocess.DEVNULL)


def ref_count_15mers(reads, out_table, threads=1):
    subprocess.check_call([os.path.join(REF_DIR, "count-15mers"), reads, out_table, str(threads)],
                          stdout=subprocess.DEVNULL)


def ref_search_15mers(table, reads, out_txt, bin_size, bins, threads=1):
    subprocess.check_call([os.path.join(REF_DIR, "search-15mers"), table, reads, out_txt, str(bin_size), str(bins),
                           str(threads)], stdout=subprocess.DEVNULL)


def table_file_sparse(path):
    """15mers-counts file -> (keys u32[], counts u32[]) of the non-zero entries."""
    size = int(np.fromfile(path, dtype=np.uint64, count=1)[0])
    assert size == TABLE_SIZE, size
    mm = np.memmap(path, dtype=np.uint32, mode="r", offset=8, shape=(size,))
    keys = []
    step = 1 << 26
    for lo in range(0, size, step):
        nz = np.flatnonzero(mm[lo:lo + step])
        if nz.size:
            keys.append(nz.astype(np.uint64) + lo)
    keys = np.concatenate(keys).astype(np.uint32) if keys else np.zeros(0, np.uint32)
    vals = np.asarray(mm[keys]) if keys.size else np.zeros(0, np.uint32)
    del mm
    return keys, vals
