"""Python face of include/hlacount.h (ctypes). Mirrors the C ABI one to one; numpy arrays are host buffers,
torch CUDA tensors (or raw device pointers) are device buffers."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, lib

NOT_A_CELL = 0xFFFFFFFF
CODE_NONE = 0xFF


def _vp(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return C.c_void_p(a.ctypes.data)
    if hasattr(a, "data_ptr"):
        return C.c_void_p(a.data_ptr())
    return C.c_void_p(int(a))


def pack_reads(seqs, words_per_read=None):
    """DnaString::from_acgt_bytes packing through the library's own packer (hlac_pack_read)."""
    n = len(seqs)
    mx = max((len(s) for s in seqs), default=0)
    wpr = words_per_read or max(1, (mx + 31) // 32)
    out = np.zeros((n, wpr), dtype=np.uint64)
    lens = np.zeros(n, dtype=np.uint16)
    L = lib()
    for i, s in enumerate(seqs):
        b = s.encode() if isinstance(s, str) else s
        check(L.hlac_pack_read(b, C.c_uint32(len(b)), C.c_void_p(out[i].ctypes.data), C.c_uint32(wpr)))
        lens[i] = len(b)
    return out, lens


class Index:
    """hlac_index: a Pseudoaligner<Kmer24> on one device (device=-1: host-only, inspection)."""

    def __init__(self, handle):
        self.h = handle

    @classmethod
    def from_idx(cls, path, device=0):
        h = C.c_void_p()
        check(lib().hlac_index_from_idx_file(str(path).encode(), C.c_int(device), C.byref(h)))
        return cls(h)

    @classmethod
    def from_fasta(cls, path, allele_status=None, device=0):
        h = C.c_void_p()
        check(lib().hlac_index_from_fasta(str(path).encode(), str(allele_status).encode() if allele_status else None,
                                          C.c_int(device), C.byref(h)))
        return cls(h)

    @classmethod
    def from_sequences(cls, seqs, names, device=0):
        """seqs: list of ASCII strings; packed here with the library packer."""
        words, starts, lens = [], [], []
        for s in seqs:
            w, _ = pack_reads([s])
            starts.append(sum(len(x) for x in words))
            words.append(w[0])
            lens.append(len(s))
        packed = np.concatenate(words) if words else np.zeros(1, np.uint64)
        starts = np.asarray(starts, dtype=np.uint64)
        lens = np.asarray(lens, dtype=np.uint32)
        arr = (C.c_char_p * len(names))(*[n.encode() for n in names])
        h = C.c_void_p()
        check(lib().hlac_index_from_sequences(_vp(packed), _vp(starts), _vp(lens), arr, C.c_uint32(len(names)),
                                              C.c_int(device), C.byref(h)))
        return cls(h)

    def close(self):
        if getattr(self, "h", None):
            lib().hlac_index_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def info(self):
        i = _lib.IndexInfo()
        check(lib().hlac_index_get_info(self.h, C.byref(i)))
        return {k: getattr(i, k) for k, _ in _lib.IndexInfo._fields_}

    @property
    def tx_names(self):
        return [lib().hlac_index_tx_name(self.h, C.c_uint32(i)).decode() for i in range(self.info["n_tx"])]

    def nodes(self):
        """[(seq, exts, colours)] — same shape as the oracle's Index.nodes()"""
        info = self.info
        out = []
        buf = C.create_string_buffer(1 << 20)
        cols = np.zeros(max(info["n_tx"], 1), dtype=np.uint32)
        ln, ex, nc = C.c_uint32(), C.c_uint32(), C.c_uint32()
        for i in range(info["n_nodes"]):
            check(lib().hlac_index_node(self.h, C.c_uint32(i), buf, C.c_uint32(len(buf)), C.byref(ln), C.byref(ex),
                                        _vp(cols), C.c_uint32(len(cols)), C.byref(nc)))
            out.append((buf.raw[:ln.value].decode(), ex.value, tuple(int(x) for x in cols[:nc.value])))
        return out


def make_batch(seq, lens, cell, umi, flags):
    """Builds an hlac_batch over host numpy arrays or device tensors; returns (struct, keepalive)."""
    n = len(lens)
    if isinstance(seq, np.ndarray):
        seq = np.ascontiguousarray(seq, dtype=np.uint64)
        lens = np.ascontiguousarray(lens, dtype=np.uint16)
        cell = np.ascontiguousarray(cell, dtype=np.uint32)
        umi = np.ascontiguousarray(umi, dtype=np.uint32)
        flags = np.ascontiguousarray(flags, dtype=np.uint8)
        wpr = seq.size // max(n, 1)
    else:
        wpr = seq.numel() // max(n, 1)
    b = _lib.Batch(_vp(seq), _vp(lens), _vp(cell), _vp(umi), _vp(flags), n, wpr)
    return b, (seq, lens, cell, umi, flags)


class CountSession:
    """hlac_count: one map_and_count_pseudo run (src/mapping.rs:640-821)."""

    def __init__(self, cds, genomic, n_cells, primary_only=False):
        self.h = C.c_void_p()
        self.cds, self.gen = cds, genomic
        check(lib().hlac_count_new(cds.h, genomic.h, C.c_uint32(n_cells), C.c_int(int(primary_only)), C.byref(self.h)))
        self.n_cells = n_cells

    def close(self):
        if getattr(self, "h", None):
            lib().hlac_count_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def nrows(self):
        return lib().hlac_count_nrows(self.h)

    @property
    def row_names(self):
        return [lib().hlac_count_row_name(self.h, C.c_uint32(i)).decode() for i in range(self.nrows)]

    def reserve(self, n):
        check(lib().hlac_count_reserve(self.h, C.c_uint64(n)))

    def record_codes(self, on=True):
        check(lib().hlac_count_record_codes(self.h, C.c_int(int(on))))

    def push(self, seq, lens, cell, umi, flags):
        b, keep = make_batch(seq, lens, cell, umi, flags)
        self._keep = keep
        if isinstance(keep[0], np.ndarray):
            check(lib().hlac_count_push(self.h, C.byref(b)))
        else:
            check(lib().hlac_count_push_device(self.h, C.byref(b)))

    def sync(self):
        check(lib().hlac_count_sync(self.h))

    def last_codes(self, n):
        out = np.zeros(n, dtype=np.uint8)
        check(lib().hlac_count_last_codes(self.h, _vp(out), C.c_uint64(n)))
        return out

    def reduce(self):
        """-> (device pointer of the dense u32 [n_cells][nrows] matrix, element count)"""
        p = C.c_void_p()
        n = C.c_uint64()
        check(lib().hlac_count_reduce(self.h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def entries(self):
        coo = _lib.Coo()
        m = _lib.Metrics()
        check(lib().hlac_count_entries(self.h, C.byref(coo), C.byref(m)))
        ent = [(coo.row[i], coo.col[i], coo.val[i]) for i in range(coo.n)]
        lib().hlac_coo_free(C.byref(coo))
        return ent, [getattr(m, k) for k, _ in _lib.Metrics._fields_]

    def finish(self):
        self.reduce()
        return self.entries()

    def reset(self):
        check(lib().hlac_count_reset(self.h))

    def timing(self):
        ms = (C.c_float * 4)()
        n = C.c_uint64()
        check(lib().hlac_count_timing(self.h, ms, C.byref(n)))
        return dict(map_ms=ms[0], sort_ms=ms[1], consensus_ms=ms[2], h2d_ms=ms[3], launches=n.value)

    def probe_stats(self):
        a, b = C.c_uint64(), C.c_uint64()
        check(lib().hlac_count_probe_stats(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value
