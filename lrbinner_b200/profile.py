"""Buffer-level Python view of the profile stage (thin wrappers over the C ABI).

  PackedReads  — a read set in the packed host layout (lrb_reads*), from a FASTA/FASTQ file, from
                 ASCII bases, or layout-only from lengths.
  Context      — one GPU: seam-to-seam `profile()` over host buffers (lrb_profile_host).
  DeviceReads / dev_* — torch-tensor plumbing for the device-pointer level (lrb_dev_*): used by the
                 device-resident benchmark and the multi-GPU driver (lrbinner_b200/dist.py), where
                 torch.distributed needs tensors over the same memory.

The arithmetic lives in the CUDA kernels; nothing here computes profiles on the CPU.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import ReadsView, check, lib

COMP_WIDTH = {3: 32, 4: 136, 5: 512}
LRB_MAX_CHUNKS = 64


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class PackedReads:
    """Owns an lrb_reads* (host memory, page-locked when a GPU is present)."""

    def __init__(self, handle):
        self._h = handle
        self.view = ReadsView()
        check(lib.lrb_reads_view_get(self._h, C.byref(self.view)))

    # -- constructors ---------------------------------------------------------------------------
    @classmethod
    def from_file(cls, path, threads=8):
        h = C.c_void_p()
        check(lib.lrb_reads_from_file(str(path).encode(), int(threads), C.byref(h)))
        return cls(h)

    @classmethod
    def from_sequences(cls, seqs, threads=8):
        """seqs: iterable of bytes/str (one per read)."""
        bs = [s if isinstance(s, (bytes, bytearray)) else s.encode("latin-1") for s in seqs]
        offsets = np.zeros(len(bs) + 1, dtype=np.uint64)
        if bs:
            offsets[1:] = np.cumsum([len(b) for b in bs], dtype=np.uint64)
        bases = np.frombuffer(b"".join(bs) + b"\0", dtype=np.uint8)
        return cls.from_ascii(bases, offsets, threads)

    @classmethod
    def from_ascii(cls, bases, offsets, threads=8):
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        h = C.c_void_p()
        check(lib.lrb_reads_from_ascii(_ptr(bases), _ptr(offsets), len(offsets) - 1, int(threads), C.byref(h)))
        return cls(h)

    @classmethod
    def from_lengths(cls, lengths):
        lengths = np.ascontiguousarray(lengths, dtype=np.uint32)
        h = C.c_void_p()
        check(lib.lrb_reads_from_lengths(_ptr(lengths), len(lengths), C.byref(h)))
        return cls(h)

    def slice(self, read_lo, read_hi):
        """Reads [read_lo, read_hi) as a read set of its own (shares the packed stream with self, which it keeps alive)."""
        h = C.c_void_p()
        check(lib.lrb_reads_slice(self._h, int(read_lo), int(read_hi), C.byref(h)))
        s = PackedReads(h)
        s._parent = self
        return s

    # -- accessors ------------------------------------------------------------------------------
    n_reads = property(lambda self: int(self.view.n_reads))
    n_blocks = property(lambda self: int(self.view.n_blocks))
    n_tiles = property(lambda self: int(self.view.n_tiles))
    total_bases = property(lambda self: int(self.view.total_bases))

    def _arr(self, ptr, n):
        if n == 0:
            return np.zeros(0, dtype=np.uint32)
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint32)), shape=(n,))

    codes = property(lambda self: self._arr(self.view.codes, 2 * self.n_blocks + 2))
    valid = property(lambda self: self._arr(self.view.valid, self.n_blocks + 1))
    read_len = property(lambda self: self._arr(self.view.read_len, self.n_reads))
    read_blk = property(lambda self: self._arr(self.view.read_blk, self.n_reads + 1))
    tile_read = property(lambda self: self._arr(self.view.tile_read, self.n_tiles))
    tile_blk = property(lambda self: self._arr(self.view.tile_blk, self.n_tiles))

    def index_valid(self, threads=8):
        """(Re)build the validity-exception list from the current `valid` words; returns its length."""
        n = C.c_uint64(0)
        check(lib.lrb_reads_index_valid(self._h, int(threads), C.byref(n)))
        return int(n.value)

    def exceptions(self):
        """(blocks, words) of the validity-exception list (copies; empty when none was built)."""
        blk, word, n = C.c_void_p(), C.c_void_p(), C.c_uint64(0)
        check(lib.lrb_reads_exceptions(self._h, C.byref(blk), C.byref(word), C.byref(n)))
        if not n.value:
            return np.zeros(0, dtype=np.uint32), np.zeros(0, dtype=np.uint32)
        return (np.array(self._arr(blk, int(n.value)), copy=True), np.array(self._arr(word, int(n.value)), copy=True))

    def unpack(self, i):
        n = int(self.read_len[i])
        buf = C.create_string_buffer(max(n, 1))
        check(lib.lrb_reads_unpack(self._h, i, buf, n))
        return buf.raw[:n]

    def close(self):
        if self._h:
            lib.lrb_reads_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Context:
    """One GPU, or several driven from this process (device = an ordinal or a list of ordinals; the first is the
    primary and ends up with the complete table).  profile() is the seam-to-seam call: host buffers in, host buffers out."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        if isinstance(device, (list, tuple)):
            ids = (C.c_int * len(device))(*[int(d) for d in device])
            check(lib.lrb_ctx_create_multi(ids, len(device), C.byref(self._h)))
        else:
            check(lib.lrb_ctx_create(int(device), C.byref(self._h)))
        self.device = device

    def profile(self, reads, k=None, bin_size=None, bins=None, want_table=False, use_loaded_table=False, out=None,
                keep_table=False):
        """Returns dict(comp=[N,P] u32, hist=[N,bins] u32, sums=[N] u32, table=[2^30] u32) for the requested parts.
        `out` may carry preallocated (ideally page-locked) arrays under the same keys."""
        n = reads.n_reads
        out = dict(out or {})
        comp = hist = sums = table = None
        if k is not None:
            comp = out.get("comp")
            if comp is None:
                comp = np.zeros((n, COMP_WIDTH[k]), dtype=np.uint32)
        if bins is not None:
            hist = out.get("hist")
            if hist is None:
                hist = np.zeros((n, max(int(bins), 1)), dtype=np.uint32)
            sums = out.get("sums")
            if sums is None:
                sums = np.zeros(n, dtype=np.uint32)
        if want_table:
            table = out.get("table")
            if table is None:
                table = np.empty(_lib.TABLE_ENTRIES, dtype=np.uint32)
        check(lib.lrb_profile_host(self._h, reads._h, int(k) if k is not None else 0,
                                   int(bin_size) if bin_size is not None else 1, int(bins) if bins is not None else 1,
                                   _ptr(comp) if comp is not None else None, _ptr(hist) if hist is not None else None,
                                   _ptr(sums) if sums is not None else None, _ptr(table) if table is not None else None,
                                   (_lib.PROFILE_USE_LOADED_TABLE if use_loaded_table else 0) |
                                   (_lib.PROFILE_KEEP_TABLE if keep_table else 0)))
        return {"comp": comp, "hist": hist, "sums": sums, "table": table}

    def info(self):
        """How the last profile() ran: devices, batches, table path, wall / exchange milliseconds."""
        ri = _lib.RunInfo()
        check(lib.lrb_ctx_last_info(self._h, C.byref(ri)))
        return {f: getattr(ri, f) for f, _ in ri._fields_}

    def timings(self):
        ms = (C.c_float * 7)()
        check(lib.lrb_ctx_last_timings(self._h, ms))
        # phases overlap (3-stream pipeline): h2d runs beside composition+partition; "total" is the whole call
        return dict(zip(("h2d", "composition_partition", "table_passes", "mirror", "search_direct", "d2h_tail", "total"),
                        [float(x) for x in ms]))

    def table_load(self, path):
        check(lib.lrb_ctx_table_load(self._h, str(path).encode()))

    def table_save(self, path):
        check(lib.lrb_ctx_table_save(self._h, str(path).encode()))

    def close(self):
        if self._h:
            lib.lrb_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def pinned_empty(shape, dtype=np.uint32):
    """numpy array over cudaHostAlloc memory (freed when the array's base object dies)."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    p = lib.lrb_pinned_alloc(max(n, 16))
    if not p:
        raise _lib.LrbError(_lib.LRB_ECUDA, _lib.last_error())

    class _Owner:
        def __init__(self, p):
            self.p = p

        def __del__(self):
            lib.lrb_pinned_free(self.p)

    owner = _Owner(p)
    buf = (C.c_uint8 * max(n, 16)).from_address(p)
    buf._owner = owner
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    return arr


# ---- device-pointer level over torch tensors -------------------------------------------------------

class DeviceReads:
    """The packed read set resident in HBM as torch tensors (int32 views of the u32 words)."""

    def __init__(self, reads, device, upload=True):
        import torch
        self.torch = torch
        self.device = torch.device(device)
        self.n_reads, self.n_blocks, self.n_tiles, self.total_bases = reads.n_reads, reads.n_blocks, reads.n_tiles, reads.total_bases

        def up(a, n, fill=True):
            t = torch.zeros(max(n, 1), dtype=torch.int32, device=self.device)
            if fill and len(a):
                t[:len(a)].copy_(torch.from_numpy(a.view(np.int32)))
            return t

        self.codes = up(reads.codes, 2 * self.n_blocks + 2, upload)
        self.valid = up(reads.valid, self.n_blocks + 1, upload)
        self.read_len = up(reads.read_len, self.n_reads)
        self.read_blk = up(reads.read_blk, self.n_reads + 1)
        self.tile_read = up(reads.tile_read, self.n_tiles)
        self.tile_blk = up(reads.tile_blk, self.n_tiles)
        # int64 copy: np.searchsorted then takes any integer type without converting the whole array on every call
        # (an O(n_tiles) conversion per lookup cost 13 ms per chunk of the 4-GPU e2e step)
        self.tile_read_host = np.asarray(reads.tile_read, dtype=np.int64).copy()
        self.view = ReadsView(self.n_reads, self.n_blocks, self.n_tiles, self.total_bases, self.codes.data_ptr(),
                              self.valid.data_ptr(), self.read_len.data_ptr(), self.read_blk.data_ptr(),
                              self.tile_read.data_ptr(), self.tile_blk.data_ptr())

    def download_into(self, reads):
        """Copy the device stream back into the host buffers of `reads` (same layout)."""
        torch = self.torch
        torch.from_numpy(reads.codes.view(np.int32)).copy_(self.codes[:2 * self.n_blocks + 2])
        torch.from_numpy(reads.valid.view(np.int32)).copy_(self.valid[:self.n_blocks + 1])

    def tile_range_for_reads(self, read_lo, read_hi):
        lo = int(np.searchsorted(self.tile_read_host, np.int64(read_lo), side="left"))
        hi = int(np.searchsorted(self.tile_read_host, np.int64(read_hi), side="left"))
        return lo, hi


def _stream():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def dev_composition(dr, k, counts, tile_lo=0, tile_hi=None):
    check(lib.lrb_dev_composition(C.byref(dr.view), k, C.c_void_p(counts.data_ptr()), tile_lo,
                                  dr.n_tiles if tile_hi is None else tile_hi, _stream()))


def dev_count(dr, table, blk_lo=0, blk_hi=None, key_lo=0, key_hi=_lib.TABLE_ENTRIES):
    check(lib.lrb_dev_count(C.byref(dr.view), C.c_void_p(table.data_ptr()), blk_lo, dr.n_blocks if blk_hi is None else blk_hi,
                            key_lo, min(key_hi, _lib.TABLE_ENTRIES), _stream()))


def dev_fill_valid(dr, exc_blk=None, exc_word=None):
    """Rebuild dr.valid on the device from the read lengths + the exception list (torch int32 tensors or None)."""
    n = 0 if exc_blk is None else int(exc_blk.numel())
    check(lib.lrb_dev_fill_valid(C.byref(dr.view), C.c_void_p(exc_blk.data_ptr()) if n else None,
                                 C.c_void_p(exc_word.data_ptr()) if n else None, n, _stream()))


def dev_mirror(table):
    check(lib.lrb_dev_mirror(C.c_void_p(table.data_ptr()), _stream()))


def dev_search(dr, table, bin_size, bins, hist, sums, tile_lo=0, tile_hi=None, key_lo=0, key_hi=_lib.TABLE_ENTRIES):
    check(lib.lrb_dev_search(C.byref(dr.view), C.c_void_p(table.data_ptr()), bin_size, bins, C.c_void_p(hist.data_ptr()),
                             C.c_void_p(sums.data_ptr()), tile_lo, dr.n_tiles if tile_hi is None else tile_hi, key_lo,
                             min(key_hi, _lib.TABLE_ENTRIES), _stream()))


class PartitionWorkspace:
    """Scratch of the L2-resident table passes: 4-byte list entries for up to `capacity` windows, the per-step
    tables, the per-block read index, and the lrb_partition descriptor the C ABI fills."""

    def __init__(self, dr, capacity=None, max_chunks=LRB_MAX_CHUNKS, sub_capacity=None):
        torch = dr.torch
        self.dr = dr
        self.capacity = int(capacity if capacity is not None else max(dr.n_blocks * 32, 1))
        self.keys = torch.empty(self.capacity, dtype=torch.int32, device=dr.device)
        # second-level lists of the shared-memory count: 2 B per window, fixed-size segments -> 2x headroom (3x for small sets)
        self.sub_capacity = int(sub_capacity if sub_capacity is not None
                                else (2 if self.capacity >= (1 << 28) else 3) * self.capacity + (1 << 22))
        self.sub = torch.empty(self.sub_capacity, dtype=torch.int16, device=dr.device) if self.sub_capacity else None
        self.small = torch.zeros(_lib.PART_SMALL_U64, dtype=torch.int64, device=dr.device)
        self.step_capacity = int(lib.lrb_partition_step_capacity(dr.n_blocks, max_chunks))
        self.steps = torch.empty(int(lib.lrb_partition_steps_words(self.step_capacity)), dtype=torch.int32, device=dr.device)
        self.blk_read = torch.empty(max(dr.n_blocks, 1), dtype=torch.int32, device=dr.device)
        check(lib.lrb_dev_fill_blk_read(C.byref(dr.view), C.c_void_p(self.blk_read.data_ptr()), _stream()))
        self.part = _lib.Partition(keys=self.keys.data_ptr(), small=self.small.data_ptr(), steps=self.steps.data_ptr(),
                                   sub=self.sub.data_ptr() if self.sub is not None else None, capacity=self.capacity,
                                   sub_capacity=self.sub_capacity, step_capacity=self.step_capacity)

    def begin(self, with_rids=True, key_lo=0, key_hi=_lib.TABLE_ENTRIES, log2_bucket_keys=24, count=True):
        """count=False: the partition will only be searched — skip building the second-level (count) lists."""
        self.part.sub = self.sub.data_ptr() if (count and self.sub is not None) else None
        check(lib.lrb_dev_partition_begin(C.byref(self.part), 1 if with_rids else 0, key_lo, min(key_hi, _lib.TABLE_ENTRIES),
                                          log2_bucket_keys, _stream()))

    def add(self, blk_lo=0, blk_hi=None):
        check(lib.lrb_dev_partition_add(C.byref(self.dr.view), C.c_void_p(self.blk_read.data_ptr()), blk_lo,
                                        self.dr.n_blocks if blk_hi is None else blk_hi, C.byref(self.part), _stream()))

    def build(self, with_rids=True, blk_lo=0, blk_hi=None, key_lo=0, key_hi=_lib.TABLE_ENTRIES, log2_bucket_keys=24, grow=False,
              count=True):
        self.begin(with_rids, key_lo, key_hi, log2_bucket_keys, count)
        self.add(blk_lo, blk_hi)
        if grow:      # verify (synchronises); if the lists did not fit, the device told us the size: grow once and redo
            needed = C.c_uint64(0)
            rc = lib.lrb_dev_partition_check(C.byref(self.part), C.byref(needed), _stream())
            if rc == _lib.LRB_ENOMEM:
                torch = self.dr.torch
                self.keys = None
                torch.cuda.empty_cache()
                self.capacity = int(needed.value) + int(needed.value) // 64 + 1024
                self.keys = torch.empty(self.capacity, dtype=torch.int32, device=self.dr.device)
                self.part.keys, self.part.capacity = self.keys.data_ptr(), self.capacity
                if self.sub is not None and self.sub_capacity < 2 * self.capacity:
                    self.sub = None
                    torch.cuda.empty_cache()
                    self.sub_capacity = 2 * self.capacity + (1 << 22)
                    self.sub = torch.empty(self.sub_capacity, dtype=torch.int16, device=self.dr.device)
                    self.part.sub_capacity = self.sub_capacity
                self.begin(with_rids, key_lo, key_hi, log2_bucket_keys, count)
                self.add(blk_lo, blk_hi)
                rc = lib.lrb_dev_partition_check(C.byref(self.part), C.byref(needed), _stream())
            check(rc)

    def check(self):
        """Synchronises; raises LrbError(LRB_ENOMEM) if the lists overflowed `capacity`. Returns entries needed."""
        needed = C.c_uint64(0)
        check(lib.lrb_dev_partition_check(C.byref(self.part), C.byref(needed), _stream()))
        return int(needed.value)

    def apply(self, table, count=True, search=False, bin_size=1, bins=1, hist=None, sums=None, smem_count=True,
              bucket_lo=0, bucket_hi=None, overwrite=False):
        """smem_count: count through the second-level, shared-memory path (falls back per bucket on key skew).
        bucket_lo/bucket_hi: only these buckets (sums are rewritten by the call that includes the last one).
        overwrite: the count WRITES the table slices of the applied buckets (no memset by the caller; this partition must
        hold all windows of those keys)."""
        mode = (1 if count else 0) | (2 if search else 0) | (4 if (count and smem_count) else 0) | (8 if (count and overwrite) else 0)
        check(lib.lrb_dev_partition_apply_range(C.byref(self.part), mode, bucket_lo, 64 if bucket_hi is None else bucket_hi,
                                                C.c_void_p(table.data_ptr()), bin_size, bins,
                                                C.c_void_p(hist.data_ptr()) if hist is not None else None,
                                                C.c_void_p(sums.data_ptr()) if sums is not None else None, _stream()))


def dev_table15_partitioned(dr, ws, table, do_count=True, bin_size=1, bins=1, hist=None, sums=None, blk_lo=0, blk_hi=None,
                            key_lo=0, key_hi=_lib.TABLE_ENTRIES, log2_bucket_keys=24, smem_count=True):
    """One-shot: partition the windows of [blk_lo, blk_hi) x [key_lo, key_hi), then count and/or search per bucket."""
    search = hist is not None
    ws.build(search, blk_lo, blk_hi, key_lo, key_hi, log2_bucket_keys, count=do_count and smem_count)
    ws.apply(table, do_count, search, bin_size, bins, hist, sums, smem_count=smem_count)


def dev_format_composition(counts, read_len, n_reads, k, text):
    check(lib.lrb_dev_format_composition(C.c_void_p(counts.data_ptr()), C.c_void_p(read_len.data_ptr()), n_reads, k,
                                         C.c_void_p(text.data_ptr()), _stream()))


def dev_format_coverage(hist, sums, n_reads, bins, text):
    check(lib.lrb_dev_format_coverage(C.c_void_p(hist.data_ptr()), C.c_void_p(sums.data_ptr()), n_reads, bins,
                                      C.c_void_p(text.data_ptr()), _stream()))


def kmer_lut(k):
    lut = np.zeros(4 ** k, dtype=np.uint16)
    width = lib.lrb_kmer_lut(k, _ptr(lut))
    return lut, width
