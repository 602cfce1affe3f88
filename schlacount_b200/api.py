"""Python face of include/hlacount.h (ctypes). Mirrors the C ABI one to one; numpy arrays are host buffers,
torch CUDA tensors (or raw device pointers) are device buffers."""
import ctypes as C
import os

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


def set_seed_stride(stride):
    """hlac_set_seed_stride: 1 (default) or 3; sessions adopt the value current when they are created"""
    check(lib().hlac_set_seed_stride(C.c_int(int(stride))))


def get_seed_stride():
    return int(lib().hlac_get_seed_stride())


class Comm:
    """hlac_comm: one rank of a multi-GPU run (NCCL communicator owned by the library)."""

    ID_BYTES = 128

    def __init__(self, handle):
        self.h = handle

    @staticmethod
    def unique_id():
        buf = (C.c_uint8 * Comm.ID_BYTES)()
        check(lib().hlac_comm_unique_id(buf))
        return bytes(buf)

    @classmethod
    def init_rank(cls, uid, rank, nranks, device):
        h = C.c_void_p()
        buf = (C.c_uint8 * Comm.ID_BYTES).from_buffer_copy(uid)
        check(lib().hlac_comm_init_rank(buf, C.c_int(rank), C.c_int(nranks), C.c_int(device), C.byref(h)))
        return cls(h)

    @classmethod
    def init_all(cls, devices):
        n = len(devices)
        arr = (C.c_int * n)(*devices)
        hs = (C.c_void_p * n)()
        check(lib().hlac_comm_init_all(arr, C.c_int(n), hs))
        return [cls(C.c_void_p(h)) for h in hs]

    @classmethod
    def from_torch_distributed(cls, device):
        """bootstrap over an initialised torch.distributed group: rank 0's id is broadcast as a byte tensor"""
        import torch
        import torch.distributed as dist
        rank, world = dist.get_rank(), dist.get_world_size()
        on_gpu = dist.get_backend() == "nccl"
        t = torch.zeros(Comm.ID_BYTES, dtype=torch.uint8)
        if rank == 0:
            t = torch.frombuffer(bytearray(cls.unique_id()), dtype=torch.uint8).clone()
        if on_gpu:
            t = t.to(torch.device("cuda", device))
        dist.broadcast(t, 0)
        return cls.init_rank(bytes(t.cpu().numpy().tobytes()), rank, world, device)

    @property
    def info(self):
        r, n, d, v = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        check(lib().hlac_comm_info(self.h, C.byref(r), C.byref(n), C.byref(d), C.byref(v)))
        return dict(rank=r.value, nranks=n.value, device=d.value, nccl_version=v.value)

    def close(self):
        if getattr(self, "h", None):
            lib().hlac_comm_free(self.h)
            self.h = None


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

    def write_idx(self, path):
        check(lib().hlac_index_write_idx(self.h, os.fsencode(path)))

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
    b = _lib.Batch(_vp(seq), _vp(lens), _vp(cell), _vp(umi), _vp(flags), n, wpr, None)
    return b, (seq, lens, cell, umi, flags)


PACKED_NOT_A_CELL = 0x7FFFFF


def pack_records(seq, lens, cell, umi, flags):
    """hlac_pack_records: the five SoA arrays -> one u64 [n, wpr + 1] array of packed records (sequence words, then
    umi << 32 | primary << 31 | len << 23 | cell)."""
    b, keep = make_batch(seq, lens, cell, umi, flags)
    rec = np.zeros((int(b.n), int(b.words_per_read) + 1), dtype=np.uint64)
    check(lib().hlac_pack_records(C.byref(b), _vp(rec)))
    return rec


def pack_wire(seq, lens, cell, umi, flags, read_len, cell_bits, umi_bits):
    """hlac_pack_wire: the five SoA arrays -> u32 [n, words] wire records (bases, primary, cell, UMI key as one bit stream)"""
    b, keep = make_batch(seq, lens, cell, umi, flags)
    words = int(lib().hlac_wire_words(read_len, cell_bits, umi_bits))
    rec = np.zeros((int(b.n), words), dtype=np.uint32)
    check(lib().hlac_pack_wire(C.byref(b), C.c_uint32(read_len), C.c_uint32(cell_bits), C.c_uint32(umi_bits), _vp(rec)))
    return rec


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

    def set_stream(self, stream_ptr):
        check(lib().hlac_count_set_stream(self.h, C.c_void_p(stream_ptr)))

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

    def push_packed(self, rec, n=None, words_per_read=None):
        """packed records (pack_records): a numpy u64 [n, wpr + 1] array = host buffer, a torch CUDA tensor = device buffer"""
        pb = _lib.PackedBatch()
        if isinstance(rec, np.ndarray):
            rec = np.ascontiguousarray(rec, dtype=np.uint64)
            pb.n, pb.words_per_read = rec.shape[0], rec.shape[1] - 1
            pb.rec = rec.ctypes.data
            check(lib().hlac_count_push_packed(self.h, C.byref(pb)))
        else:
            pb.n = rec.shape[0] if n is None else n
            pb.words_per_read = (rec.shape[1] - 1) if words_per_read is None else words_per_read
            pb.rec = rec.data_ptr()
            check(lib().hlac_count_push_packed_device(self.h, C.byref(pb)))
        self._keep = rec

    def push_wire(self, rec, read_len, cell_bits, umi_bits, n=None):
        """wire records (pack_wire): a numpy u32 [n, words] array = host buffer, a torch CUDA tensor = device buffer"""
        wb = _lib.WireBatch()
        wb.read_len, wb.cell_bits, wb.umi_bits = read_len, cell_bits, umi_bits
        wb.words_per_record = int(lib().hlac_wire_words(read_len, cell_bits, umi_bits))
        if isinstance(rec, np.ndarray):
            rec = np.ascontiguousarray(rec, dtype=np.uint32)
            wb.n = rec.shape[0] if n is None else n
            wb.rec = rec.ctypes.data
            check(lib().hlac_count_push_wire(self.h, C.byref(wb)))
        else:
            wb.n = rec.shape[0] if n is None else n
            wb.rec = rec.data_ptr()
            check(lib().hlac_count_push_wire_device(self.h, C.byref(wb)))
        self._keep = rec

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
        n = int(coo.n)
        if n:
            rows = np.ctypeslib.as_array(coo.row, shape=(n,)).tolist()
            cols = np.ctypeslib.as_array(coo.col, shape=(n,)).tolist()
            vals = np.ctypeslib.as_array(coo.val, shape=(n,)).tolist()
            ent = list(zip(rows, cols, vals))
        else:
            ent = []
        lib().hlac_coo_free(C.byref(coo))
        return ent, [getattr(m, k) for k, _ in _lib.Metrics._fields_]

    def entries_arrays(self, copy=True):
        """Like entries() but as three numpy arrays (row, col, val) — no Python-object conversion. copy=False returns
        views of the session's pinned result buffer (valid until the next entries/reset on this session)."""
        coo = _lib.Coo()
        m = _lib.Metrics()
        check(lib().hlac_count_entries(self.h, C.byref(coo), C.byref(m)))
        n = int(coo.n)
        if n:
            out = tuple(np.ctypeslib.as_array(p, shape=(n,)) for p in (coo.row, coo.col, coo.val))
            if copy:
                out = tuple(a.copy() for a in out)
        else:
            out = tuple(np.zeros(0, np.uint32) for _ in range(3))
        lib().hlac_coo_free(C.byref(coo))
        return out, [getattr(m, k) for k, _ in _lib.Metrics._fields_]

    def attach_comm(self, comm, root=-1):
        """hlac_count_attach_comm: finish() then merges over the communicator (root < 0: all-reduce)"""
        check(lib().hlac_count_attach_comm(self.h, comm.h if comm is not None else None, C.c_int(root)))
        self._comm = comm

    def finish(self):
        """hlac_count_finish -> ([(row, col, val)], metrics)"""
        (r, c, v), m = self.finish_arrays()
        return list(zip(r.tolist(), c.tolist(), v.tolist())), m

    def finish_arrays(self, copy=True):
        """hlac_count_finish as three numpy arrays; copy=False returns views of the session's pinned result buffer"""
        coo = _lib.Coo()
        m = _lib.Metrics()
        check(lib().hlac_count_finish(self.h, C.byref(coo), C.byref(m)))
        n = int(coo.n)
        if n:
            out = tuple(np.ctypeslib.as_array(p, shape=(n,)) for p in (coo.row, coo.col, coo.val))
            if copy:
                out = tuple(a.copy() for a in out)
        else:
            out = tuple(np.zeros(0, np.uint32) for _ in range(3))
        lib().hlac_coo_free(C.byref(coo))
        return out, [getattr(m, k) for k, _ in _lib.Metrics._fields_]

    def reset(self):
        check(lib().hlac_count_reset(self.h))

    def timing(self):
        ms = (C.c_float * 4)()
        n = C.c_uint64()
        check(lib().hlac_count_timing(self.h, ms, C.byref(n)))
        return dict(map_ms=ms[0], group_ms=ms[1], coo_ms=ms[2], h2d_ms=ms[3], launches=n.value)

    def shape(self):
        buf = C.create_string_buffer(256)
        check(lib().hlac_count_shape(self.h, buf, C.c_size_t(256)))
        return buf.value.decode()

    def group_stats(self):
        k, b, f = C.c_uint64(), C.c_uint32(), C.c_uint64()
        check(lib().hlac_count_group_stats(self.h, C.byref(k), C.byref(b), C.byref(f)))
        return dict(keys=k.value, buckets=b.value, fallbacks=f.value)

    def probe_stats(self):
        a, b = C.c_uint64(), C.c_uint64()
        check(lib().hlac_count_probe_stats(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value


def _tables_to_py(t):
    def csr(n, off, tx, cnt):
        out = {}
        for c in range(n):
            out[tuple(tx[q] for q in range(off[c], off[c + 1]))] = cnt[c]
        return out
    reads_list = [(tuple(t.reads_tx[q] for q in range(t.reads_off[c], t.reads_off[c + 1])), t.reads_cnt[c])
                  for c in range(t.n_read_classes)] if t.reads_off else []
    return dict(nitems=t.nitems, nreads=t.nreads, reads_classes=reads_list,
                counts_reads=dict(reads_list), counts_umi=csr(t.n_umi_classes, t.umi_off, t.umi_tx, t.umi_cnt),
                reads_explained=[t.reads_explained[i] for i in range(t.nitems)])


class GenoSession:
    """hlac_geno: one mapping_wrapper run (src/mapping.rs:310-405) and its EqClassDb."""

    def __init__(self, index, record_reads=False, lean=False):
        self.h = C.c_void_p()
        self.index = index
        check(lib().hlac_geno_new(index.h, C.byref(self.h)))
        self.tables = None
        self.lean = lean
        if lean:
            check(lib().hlac_geno_set_lean(self.h, C.c_int(1)))
        if record_reads:
            check(lib().hlac_geno_record_reads(self.h, C.c_int(1)))

    def close(self):
        if getattr(self, "tables", None) is not None:
            lib().hlac_class_tables_free(C.byref(self.tables))
            self.tables = None
        if getattr(self, "h", None):
            lib().hlac_geno_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def push(self, seq, lens, cell, umi, forward_only=False, ordinal=None):
        """ordinal (optional u64 per read): global position in BAM order when this session only sees a shard."""
        n = len(lens)
        flags = np.zeros(n, np.uint8) if isinstance(seq, np.ndarray) else None
        if isinstance(seq, np.ndarray):
            b, keep = make_batch(seq, lens, cell, umi, flags)
            if ordinal is not None:
                ordinal = np.ascontiguousarray(ordinal, dtype=np.uint64)
                b.ordinal = _vp(ordinal)
            check(lib().hlac_geno_push(self.h, C.byref(b), C.c_int(int(forward_only))))
        else:
            wpr = seq.numel() // max(n, 1)
            b = _lib.Batch(_vp(seq), _vp(lens), _vp(cell), _vp(umi), None, n, wpr, _vp(ordinal))
            check(lib().hlac_geno_push_device(self.h, C.byref(b), C.c_int(int(forward_only))))
        self._n_last = n

    # ---- multi-GPU merge ----
    def attach_comm(self, comm):
        """hlac_geno_attach_comm: finish() / finish_begin() then merge with the other ranks inside the library (collective)"""
        check(lib().hlac_geno_attach_comm(self.h, comm.h if comm is not None else None))
        self._comm = comm

    def export_paths(self):
        """-> dict of numpy arrays (hash_a, hash_b, first_ord, count, off, colours): picklable, for an all-gather"""
        p = _lib.Paths()
        check(lib().hlac_geno_export_paths(self.h, C.byref(p)))
        n = int(p.n_paths)
        tot = int(p.off[n]) if n else 0
        out = dict(hash_a=np.ctypeslib.as_array(p.hash_a, shape=(max(n, 1),))[:n].copy(),
                   hash_b=np.ctypeslib.as_array(p.hash_b, shape=(max(n, 1),))[:n].copy(),
                   first_ord=np.ctypeslib.as_array(p.first_ord, shape=(max(n, 1),))[:n].copy(),
                   count=np.ctypeslib.as_array(p.count, shape=(max(n, 1),))[:n].copy(),
                   off=np.ctypeslib.as_array(p.off, shape=(n + 1,)).copy(),
                   colours=np.ctypeslib.as_array(p.colours, shape=(max(tot, 1),))[:tot].copy())
        lib().hlac_paths_free(C.byref(p))
        return out

    def import_merged_paths(self, parts):
        """parts: list of export_paths() dicts from every rank (same order on every rank). Builds the union and adopts it."""
        arr = (_lib.Paths * len(parts))()
        keep = []
        for k, d in enumerate(parts):
            a = {x: np.ascontiguousarray(d[x]) for x in d}
            for x in ("hash_a", "hash_b", "first_ord", "off"):
                a[x] = a[x].astype(np.uint64, copy=False)
            for x in ("count", "colours"):
                a[x] = a[x].astype(np.uint32, copy=False)
                if a[x].size == 0:
                    a[x] = np.zeros(1, np.uint32)
            keep.append(a)
            arr[k].n_paths = len(d["hash_a"])
            for x, t in (("hash_a", C.c_uint64), ("hash_b", C.c_uint64), ("first_ord", C.c_uint64), ("off", C.c_uint64),
                         ("count", C.c_uint32), ("colours", C.c_uint32)):
                setattr(arr[k], x, a[x].ctypes.data_as(C.POINTER(t)))
        merged = _lib.Paths()
        check(lib().hlac_paths_merge(arr, C.c_uint32(len(parts)), C.byref(merged)))
        try:
            check(lib().hlac_geno_import_paths(self.h, C.byref(merged)))
            return int(merged.n_paths)
        finally:
            lib().hlac_paths_free(C.byref(merged))

    def finish_begin(self):
        """-> (device pointer of the u32 per-class UMI histogram, n_classes): all-reduce it across ranks, then finish_end()"""
        p = C.c_void_p()
        n = C.c_uint64()
        check(lib().hlac_geno_finish_begin(self.h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def finish_end(self, raw=False):
        if self.tables is not None:
            lib().hlac_class_tables_free(C.byref(self.tables))
        self.tables = _lib.ClassTables()
        check(lib().hlac_geno_finish_end(self.h, C.byref(self.tables)))
        if raw:
            t = self.tables
            return dict(nitems=t.nitems, nreads=t.nreads, n_read_classes=t.n_read_classes, n_umi_classes=t.n_umi_classes)
        return _tables_to_py(self.tables)

    def last_cov(self):
        out = np.zeros(self._n_last, dtype=np.uint32)
        check(lib().hlac_geno_last_cov(self.h, _vp(out), C.c_uint64(self._n_last)))
        return out

    def finish(self, raw=False):
        """-> dict(nitems, nreads, counts_reads, counts_umi, reads_explained, reads_classes [first-seen order]);
        raw=True skips the (slow) conversion of the class tables to Python objects and returns only the sizes."""
        if self.tables is not None:
            lib().hlac_class_tables_free(C.byref(self.tables))
        self.tables = _lib.ClassTables()
        check(lib().hlac_geno_finish(self.h, C.byref(self.tables)))
        if raw:
            t = self.tables
            return dict(nitems=t.nitems, nreads=t.nreads, n_read_classes=t.n_read_classes, n_umi_classes=t.n_umi_classes,
                        reads_nnz=t.reads_off[t.n_read_classes] if (t.n_read_classes and t.reads_off) else 0,
                        umi_nnz=t.umi_off[t.n_umi_classes] if t.n_umi_classes else 0)
        return _tables_to_py(self.tables)

    def read_class_ids(self, n):
        out = np.zeros(n, dtype=np.uint32)
        check(lib().hlac_geno_read_class_ids(self.h, _vp(out), C.c_uint64(n)))
        return out

    def eqclassdb(self, read_bc, read_umi, barcodes, umis):
        """After finish() (record_reads=True, not lean): the reference's EqClassDb of this run. read_bc / read_umi index,
        per pushed read in push order, into the `barcodes` / `umis` lists of bytes."""
        read_bc, read_umi = (np.ascontiguousarray(a, dtype=np.uint32) for a in (read_bc, read_umi))

        def csr(items):
            off = np.zeros(len(items) + 1, dtype=np.uint64)
            for i, x in enumerate(items):
                off[i + 1] = off[i] + len(x)
            return off, np.frombuffer(b"".join(items) or b"\0", dtype=np.uint8).copy()
        boff, bby = csr(barcodes)
        uoff, uby = csr(umis)
        d = _lib.EqClassDbC()
        check(lib().hlac_geno_eqclassdb(self.h, C.byref(self.tables), _vp(read_bc), _vp(read_umi), C.c_uint64(len(read_bc)),
                                        _vp(boff), _vp(bby), C.c_uint32(len(barcodes)), _vp(uoff), _vp(uby),
                                        C.c_uint32(len(umis)), C.byref(d)))
        try:
            return EqClassDb._from_c(d)
        finally:
            lib().hlac_eqclassdb_free(C.byref(d))

    def stats(self):
        v = [C.c_uint64() for _ in range(5)]
        check(lib().hlac_geno_stats(self.h, *[C.byref(x) for x in v]))
        return dict(zip(("some_aln", "long_aln", "n_paths", "bitmap_tests", "dict_probes"), [x.value for x in v]))

    def timing(self):
        ms = (C.c_float * 4)()
        n = C.c_uint64()
        check(lib().hlac_geno_timing(self.h, ms, C.byref(n)))
        return dict(map_ms=ms[0], intern_ms=ms[1], finish_ms=ms[2], h2d_ms=ms[3], launches=n.value)

    def resolve_stats(self):
        ms = C.c_float()
        v = [C.c_uint64() for _ in range(4)]
        check(lib().hlac_geno_resolve_stats(self.h, C.byref(ms), *[C.byref(x) for x in v]))
        return dict(ms=ms.value, paths=v[0].value, merges=v[1].value, element_steps=v[2].value, bytes=v[3].value)

    def tables_csr(self):
        """the finished tables as numpy CSR arrays: dict(reads=(off, tx, cnt) or None in lean mode, umi=(off, tx, cnt),
        reads_explained, nreads)"""
        t = self.tables

        def csr(n, off, tx, cnt):
            n = int(n)
            if not off:
                return None
            o = np.ctypeslib.as_array(off, shape=(n + 1,)).copy()
            tot = int(o[n]) if n else 0
            return (o, np.ctypeslib.as_array(tx, shape=(max(tot, 1),))[:tot].copy(),
                    np.ctypeslib.as_array(cnt, shape=(max(n, 1),))[:n].copy())
        return dict(reads=csr(t.n_read_classes, t.reads_off, t.reads_tx, t.reads_cnt),
                    umi=csr(t.n_umi_classes, t.umi_off, t.umi_tx, t.umi_cnt),
                    reads_explained=np.ctypeslib.as_array(t.reads_explained, shape=(max(t.nitems, 1),))[:t.nitems].copy(),
                    nreads=int(t.nreads))

    def squarem(self, device=0):
        theta = np.zeros(self.tables.nitems, dtype=np.float64)
        it = C.c_uint32()
        check(lib().hlac_squarem(C.byref(self.tables), C.c_int(device), _vp(theta), C.byref(it)))
        return theta, it.value

    def call(self, theta, on_device=None):
        """the caller (em.rs:361-434). on_device=True (default in lean mode) evaluates pair overlaps on the class vectors
        still resident on the GPU (hlac_geno_call); otherwise from the host tables (hlac_call_genotypes)."""
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        c = _lib.Calls()
        if on_device if on_device is not None else self.lean:
            check(lib().hlac_geno_call(self.h, C.byref(self.tables), _vp(theta), C.byref(c)))
        else:
            check(lib().hlac_call_genotypes(self.index.h, C.byref(self.tables), _vp(theta), C.byref(c)))
        out = dict(called=[c.called[i] for i in range(c.n_called)],
                   weights=[(c.w_tx[i], c.w_em[i], c.w_frac[i]) for i in range(c.n_weights)],
                   pairs=[(c.p_a[i], c.p_b[i], c.p_frac[i]) for i in range(c.n_pairs)])
        lib().hlac_calls_free(C.byref(c))
        return out


def device_cache_trim():
    """hlac_device_cache_trim: returns the library's cached (unused) device blocks to the driver -> (bytes released, bytes live)"""
    rel, live = C.c_uint64(), C.c_uint64()
    check(lib().hlac_device_cache_trim(C.byref(rel), C.byref(live)))
    return rel.value, live.value


def squarem(nitems, classes, device=0):
    """hlac_squarem on an explicit {class tuple: count} table (classes iterated in sorted order) -> (theta, iters)"""
    keys = sorted(classes)
    off = np.zeros(len(keys) + 1, dtype=np.uint64)
    tx = []
    for i, k in enumerate(keys):
        tx.extend(k)
        off[i + 1] = len(tx)
    tx = np.asarray(tx if tx else [0], dtype=np.uint32)
    cnt = np.asarray([classes[k] for k in keys] if keys else [0], dtype=np.uint32)
    t = _lib.ClassTables()
    t.nitems = nitems
    t.n_umi_classes = len(keys)
    t.umi_off = off.ctypes.data_as(C.POINTER(C.c_uint64))
    t.umi_tx = tx.ctypes.data_as(C.POINTER(C.c_uint32))
    t.umi_cnt = cnt.ctypes.data_as(C.POINTER(C.c_uint32))
    theta = np.zeros(nitems, dtype=np.float64)
    it = C.c_uint32()
    check(lib().hlac_squarem(C.byref(t), C.c_int(device), _vp(theta), C.byref(it)))
    return theta, it.value


class EqClassDb:
    """The reference's pseudoaligner.counts.bin (bincode EqClassDb, mapping.rs:84-90) as plain Python/numpy:
    barcodes / umis = list of bytes by id, classes = list of tuples by eq_class_id, bc / umi / cls = u32 arrays (one
    triple per counted read)."""

    def __init__(self, barcodes, umis, classes, bc, umi, cls):
        self.barcodes, self.umis, self.classes = list(barcodes), list(umis), [tuple(c) for c in classes]
        self.bc, self.umi, self.cls = (np.ascontiguousarray(a, dtype=np.uint32) for a in (bc, umi, cls))

    @staticmethod
    def _strings(n, off, data):
        raw = bytes(bytearray(data[0:off[n]])) if n and off[n] else b""
        return [raw[off[i]:off[i + 1]] for i in range(n)]

    @classmethod
    def _from_c(cls, d):
        n = d.n_counts
        classes = [tuple(d.class_tx[q] for q in range(d.class_off[c], d.class_off[c + 1])) for c in range(d.n_classes)]
        arr = lambda p: np.ctypeslib.as_array(p, shape=(n,)).copy() if n else np.zeros(0, np.uint32)
        return cls(cls._strings(d.n_barcodes, d.barcode_off, d.barcode_bytes), cls._strings(d.n_umis, d.umi_off, d.umi_bytes),
                   classes, arr(d.bc), arr(d.umi), arr(d.cls))

    @classmethod
    def read(cls, path):
        d = _lib.EqClassDbC()
        check(lib().hlac_eqclassdb_read(os.fsencode(path), C.byref(d)))
        try:
            return cls._from_c(d)
        finally:
            lib().hlac_eqclassdb_free(C.byref(d))

    def _to_c(self):
        keep = []

        def csr_bytes(items):
            off = np.zeros(len(items) + 1, dtype=np.uint64)
            for i, x in enumerate(items):
                off[i + 1] = off[i] + len(x)
            data = np.frombuffer(b"".join(items) or b"\0", dtype=np.uint8).copy()
            keep.extend([off, data])
            return off.ctypes.data_as(C.POINTER(C.c_uint64)), data.ctypes.data_as(C.POINTER(C.c_uint8))
        d = _lib.EqClassDbC()
        d.n_barcodes, d.n_umis, d.n_classes, d.n_counts = len(self.barcodes), len(self.umis), len(self.classes), len(self.cls)
        d.barcode_off, d.barcode_bytes = csr_bytes(self.barcodes)
        d.umi_off, d.umi_bytes = csr_bytes(self.umis)
        coff = np.zeros(len(self.classes) + 1, dtype=np.uint64)
        for i, c in enumerate(self.classes):
            coff[i + 1] = coff[i] + len(c)
        ctx = np.asarray([t for c in self.classes for t in c] or [0], dtype=np.uint32)
        pads = [a if len(a) else np.zeros(1, np.uint32) for a in (self.bc, self.umi, self.cls)]
        keep.extend([coff, ctx] + pads)
        d.class_off, d.class_tx = coff.ctypes.data_as(C.POINTER(C.c_uint64)), ctx.ctypes.data_as(C.POINTER(C.c_uint32))
        d.bc, d.umi, d.cls = (a.ctypes.data_as(C.POINTER(C.c_uint32)) for a in pads)
        return d, keep

    def write(self, path):
        d, keep = self._to_c()
        check(lib().hlac_eqclassdb_write(os.fsencode(path), C.byref(d)))

    def counts(self, nitems, device=0):
        """EqClassDb::eq_class_counts(nitems) on the device -> the same dict GenoSession.finish() returns"""
        d, keep = self._to_c()
        t = _lib.ClassTables()
        check(lib().hlac_eqclassdb_counts(C.byref(d), C.c_uint32(nitems), C.c_int(device), C.byref(t)))
        try:
            return _tables_to_py(t)
        finally:
            lib().hlac_class_tables_free(C.byref(t))
