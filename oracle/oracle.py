"""ctypes face of the CPU oracle (oracle/hla_oracle.cpp). TEST INFRASTRUCTURE ONLY — see that file's header.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "liborc.so")
    src = os.path.join(_HERE, "hla_oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liborc.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        L.orc_last_error.restype = C.c_char_p
        for f in ("orc_index_from_idx", "orc_index_from_fasta", "orc_geno_new"):
            getattr(L, f).restype = C.c_void_p
        L.orc_index_from_idx.argtypes = [C.c_char_p]
        L.orc_index_from_fasta.argtypes = [C.c_char_p, C.c_int, C.c_char_p]
        L.orc_index_free.argtypes = [C.c_void_p]
        for f in ("orc_index_n_nodes", "orc_index_n_tx", "orc_index_n_colours"):
            getattr(L, f).restype = C.c_uint32
            getattr(L, f).argtypes = [C.c_void_p]
        L.orc_index_n_bases.restype = C.c_uint64
        L.orc_index_n_bases.argtypes = [C.c_void_p]
        L.orc_index_tx_name.restype = C.c_char_p
        L.orc_index_tx_name.argtypes = [C.c_void_p, C.c_uint32]
        L.orc_index_node_seq.restype = C.c_uint32
        L.orc_index_node_seq.argtypes = [C.c_void_p, C.c_uint32, C.c_char_p, C.c_uint32]
        L.orc_index_node_exts.restype = C.c_uint32
        L.orc_index_node_exts.argtypes = [C.c_void_p, C.c_uint32]
        L.orc_index_node_colours.restype = C.c_uint32
        L.orc_index_node_colours.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32]
        L.orc_pack_ascii.argtypes = [C.c_char_p, C.c_uint32, C.c_void_p, C.c_uint32]
        L.orc_map_read.restype = C.c_int
        L.orc_map_read.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_void_p, C.c_uint32, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
        L.orc_rows.restype = C.c_uint32
        L.orc_rows.argtypes = [C.c_void_p, C.c_char_p, C.c_uint32]
        L.orc_count_run.restype = C.c_int64
        L.orc_count_run.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64,
                                    C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_geno_free.argtypes = [C.c_void_p]
        L.orc_geno_push.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_uint64, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_geno_finish.argtypes = [C.c_void_p, C.c_uint32]
        L.orc_geno_nreads.restype = C.c_uint32
        L.orc_geno_nreads.argtypes = [C.c_void_p]
        L.orc_geno_n_read_classes.restype = C.c_uint32
        L.orc_geno_n_read_classes.argtypes = [C.c_void_p]
        L.orc_geno_reads_explained.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_geno_table.restype = C.c_uint64
        L.orc_geno_table.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64,
                                     C.c_void_p]
        L.orc_squarem.argtypes = [C.c_uint32, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_geno_call.restype = C.c_uint32
        L.orc_geno_call.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
        L.orc_geno_weights_rows.restype = C.c_uint32
        L.orc_geno_weights_rows.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
        L.orc_geno_pairs_rows.restype = C.c_uint32
        L.orc_geno_pairs_rows.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
        L.orc_parse_allele.restype = C.c_int
        L.orc_parse_allele.argtypes = [C.c_char_p, C.c_char_p, C.c_uint32, C.c_void_p]
        _LIB = L
    return _LIB


def set_stride(s):
    """seed-scan stride of the oracle's map_read (1..3 reproduce the goldens; default 1)"""
    lib().orc_set_stride(int(s))


def get_stride():
    return lib().orc_get_stride()


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Index:
    def __init__(self, handle):
        if not handle:
            raise RuntimeError("oracle: " + lib().orc_last_error().decode())
        self.h = handle

    @classmethod
    def from_idx(cls, path):
        return cls(lib().orc_index_from_idx(str(path).encode()))

    @classmethod
    def from_fasta(cls, path, allele_status=None):
        return cls(lib().orc_index_from_fasta(str(path).encode(), 1 if allele_status else 0,
                                              str(allele_status).encode() if allele_status else None))

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_index_free(self.h)
            self.h = None

    @property
    def n_nodes(self):
        return lib().orc_index_n_nodes(self.h)

    @property
    def n_tx(self):
        return lib().orc_index_n_tx(self.h)

    @property
    def n_colours(self):
        return lib().orc_index_n_colours(self.h)

    @property
    def n_bases(self):
        return lib().orc_index_n_bases(self.h)

    @property
    def tx_names(self):
        return [lib().orc_index_tx_name(self.h, i).decode() for i in range(self.n_tx)]

    def nodes(self):
        """[(seq:str, exts:int, colours:tuple)] for every node."""
        out = []
        buf = C.create_string_buffer(1 << 20)
        cb = np.zeros(max(self.n_tx, 1), dtype=np.uint32)
        for i in range(self.n_nodes):
            n = lib().orc_index_node_seq(self.h, i, buf, len(buf))
            k = lib().orc_index_node_colours(self.h, i, _p(cb), len(cb))
            out.append((buf.raw[:n].decode(), lib().orc_index_node_exts(self.h, i), tuple(int(x) for x in cb[:k])))
        return out

    def rows(self):
        buf = C.create_string_buffer(1 << 16)
        n = lib().orc_rows(self.h, buf, len(buf))
        names = buf.value.decode().split("\n")[:-1]
        assert len(names) == n
        return names

    def map_read(self, words, L, rc=False):
        """-> None or (class tuple, cov, node tuple)"""
        words = np.ascontiguousarray(words, dtype=np.uint64)
        cls = np.zeros(max(self.n_tx, 1), dtype=np.uint32)
        nodes = np.zeros(512, dtype=np.uint32)
        ncls = C.c_uint32()
        cov = C.c_uint32()
        nn = C.c_uint32()
        ok = lib().orc_map_read(self.h, _p(words), L, int(rc), _p(cls), len(cls), C.byref(ncls), C.byref(cov),
                                _p(nodes), len(nodes), C.byref(nn))
        if not ok:
            return None
        return tuple(int(x) for x in cls[:ncls.value]), cov.value, tuple(int(x) for x in nodes[:nn.value])


def pack_reads(seqs, words_per_read=None):
    """ASCII reads -> (uint64 [n, wpr] 2-bit MSB-first, uint16 len) with from_acgt_bytes semantics (N -> A)."""
    n = len(seqs)
    mx = max((len(s) for s in seqs), default=0)
    wpr = words_per_read or max(1, (mx + 31) // 32)
    out = np.zeros((n, wpr), dtype=np.uint64)
    lens = np.zeros(n, dtype=np.uint16)
    L = lib()
    for i, s in enumerate(seqs):
        b = s.encode() if isinstance(s, str) else s
        L.orc_pack_ascii(b, len(b), C.c_void_p(out[i].ctypes.data), wpr)
        lens[i] = len(b)
    return out, lens


def count_run(cds, gen, seq, lens, cell, umi, flags, primary_only=False, want_codes=False):
    """Counting path. -> dict(entries=[(row,col,val)], metrics=[6], codes, cov)"""
    n = len(lens)
    seq = np.ascontiguousarray(seq, dtype=np.uint64)
    wpr = seq.shape[1] if seq.ndim == 2 else (seq.size // max(n, 1))
    lens = np.ascontiguousarray(lens, dtype=np.uint16)
    cell = np.ascontiguousarray(cell, dtype=np.uint32)
    umi = np.ascontiguousarray(umi, dtype=np.uint32)
    flags = np.ascontiguousarray(flags, dtype=np.uint8)
    nrows = len(cds.rows())
    ncell = len(np.unique(cell))
    cap = max(1, nrows * (ncell + 1))
    r = np.zeros(cap, dtype=np.uint32)
    c = np.zeros(cap, dtype=np.uint32)
    v = np.zeros(cap, dtype=np.uint32)
    m = np.zeros(6, dtype=np.uint64)
    codes = np.zeros(n, dtype=np.uint8) if want_codes else None
    cov = np.zeros(n, dtype=np.uint32) if want_codes else None
    ne = lib().orc_count_run(cds.h, gen.h, _p(seq), wpr, _p(lens), _p(cell), _p(umi), _p(flags), n, int(primary_only),
                             _p(r), _p(c), _p(v), cap, _p(m), _p(codes), _p(cov))
    if ne < 0:
        raise RuntimeError("oracle: " + lib().orc_last_error().decode())
    return dict(entries=list(zip(r[:ne].tolist(), c[:ne].tolist(), v[:ne].tolist())), metrics=m.tolist(),
                codes=codes, cov=cov, nrows=nrows)


class Geno:
    """Genotyping path: push batches in BAM order, finish, call."""

    def __init__(self, index):
        self.ix = index
        self.g = lib().orc_geno_new()

    def __del__(self):
        if getattr(self, "g", None):
            lib().orc_geno_free(self.g)
            self.g = None

    def push(self, seq, lens, cell, umi, forward_only=False):
        n = len(lens)
        seq = np.ascontiguousarray(seq, dtype=np.uint64)
        wpr = seq.shape[1] if seq.ndim == 2 else (seq.size // max(n, 1))
        lens = np.ascontiguousarray(lens, dtype=np.uint16)
        cell = np.ascontiguousarray(cell, dtype=np.uint32)
        umi = np.ascontiguousarray(umi, dtype=np.uint32)
        cid = np.zeros(n, dtype=np.uint32)
        cov = np.zeros(n, dtype=np.uint32)
        lib().orc_geno_push(self.g, self.ix.h, _p(seq), wpr, _p(lens), _p(cell), _p(umi), n, int(forward_only),
                            _p(cid), _p(cov))
        return cid, cov

    def finish(self):
        lib().orc_geno_finish(self.g, self.ix.n_tx)

    @property
    def nreads(self):
        return lib().orc_geno_nreads(self.g)

    @property
    def n_read_classes(self):
        return lib().orc_geno_n_read_classes(self.g)

    def reads_explained(self):
        out = np.zeros(self.ix.n_tx, dtype=np.uint64)
        lib().orc_geno_reads_explained(self.g, _p(out))
        return out

    def table(self, which):
        """which: 'reads' or 'umi' -> {class tuple: count}"""
        w = 0 if which == "reads" else 1
        ntx = C.c_uint64()
        nc = lib().orc_geno_table(self.g, w, None, None, None, 0, 0, C.byref(ntx))
        off = np.zeros(nc + 1, dtype=np.uint64)
        tx = np.zeros(max(ntx.value, 1), dtype=np.uint32)
        cnt = np.zeros(max(nc, 1), dtype=np.uint32)
        lib().orc_geno_table(self.g, w, _p(off), _p(tx), _p(cnt), nc, ntx.value, None)
        return {tuple(int(t) for t in tx[int(off[i]):int(off[i + 1])]): int(cnt[i]) for i in range(nc)}

    def table_raw(self, which):
        """which: 'reads' or 'umi' -> CSR numpy arrays (off, tx, cnt) in the oracle's own (map) order"""
        w = 0 if which == "reads" else 1
        ntx = C.c_uint64()
        nc = lib().orc_geno_table(self.g, w, None, None, None, 0, 0, C.byref(ntx))
        off = np.zeros(nc + 1, dtype=np.uint64)
        tx = np.zeros(max(ntx.value, 1), dtype=np.uint32)
        cnt = np.zeros(max(nc, 1), dtype=np.uint32)
        lib().orc_geno_table(self.g, w, _p(off), _p(tx), _p(cnt), nc, ntx.value, None)
        return off, tx[:ntx.value], cnt[:nc]

    def call(self):
        """-> dict(theta, iters, called tx ids, weights rows, pairs rows)"""
        theta = np.zeros(self.ix.n_tx, dtype=np.float64)
        iters = C.c_uint32()
        called = np.zeros(self.ix.n_tx, dtype=np.uint32)
        n = lib().orc_geno_call(self.g, self.ix.h, _p(theta), C.byref(iters), _p(called), len(called))
        cap = 1024
        tx = np.zeros(cap, dtype=np.uint32)
        w = np.zeros(cap, dtype=np.float64)
        fr = np.zeros(cap, dtype=np.float64)
        nw = lib().orc_geno_weights_rows(self.g, _p(tx), _p(w), _p(fr), cap)
        a = np.zeros(cap, dtype=np.uint32)
        b = np.zeros(cap, dtype=np.uint32)
        pf = np.zeros(cap, dtype=np.float64)
        npair = lib().orc_geno_pairs_rows(self.g, _p(a), _p(b), _p(pf), cap)
        return dict(theta=theta, iters=iters.value, called=[int(x) for x in called[:n]],
                    weights=[(int(tx[i]), float(w[i]), float(fr[i])) for i in range(nw)],
                    pairs=[(int(a[i]), int(b[i]), float(pf[i])) for i in range(npair)])


def squarem(nitems, classes):
    """classes: {tuple of tx ids: count} (iterated in sorted order) -> (theta, iters)"""
    keys = sorted(classes)
    off = np.zeros(len(keys) + 1, dtype=np.uint64)
    tx = []
    cnt = np.zeros(max(len(keys), 1), dtype=np.uint32)
    for i, k in enumerate(keys):
        tx.extend(k)
        off[i + 1] = len(tx)
        cnt[i] = classes[k]
    tx = np.asarray(tx if tx else [0], dtype=np.uint32)
    theta = np.zeros(nitems, dtype=np.float64)
    iters = C.c_uint32()
    lib().orc_squarem(nitems, len(keys), _p(off), _p(tx), _p(cnt), _p(theta), C.byref(iters))
    return theta, iters.value


def squarem_ordered(nitems, keys, counts):
    """squarem with the classes iterated in exactly the given order (the reference's HashMap order is unspecified)."""
    off = np.zeros(len(keys) + 1, dtype=np.uint64)
    tx = []
    for i, k in enumerate(keys):
        tx.extend(k)
        off[i + 1] = len(tx)
    tx = np.asarray(tx if tx else [0], dtype=np.uint32)
    cnt = np.asarray(list(counts) if len(keys) else [0], dtype=np.uint32)
    theta = np.zeros(nitems, dtype=np.float64)
    iters = C.c_uint32()
    lib().orc_squarem(nitems, len(keys), _p(off), _p(tx), _p(cnt), _p(theta), C.byref(iters))
    return theta, iters.value


def parse_allele(s):
    gene = C.create_string_buffer(64)
    f = np.zeros(4, dtype=np.int32)
    ok = lib().orc_parse_allele(s.encode(), gene, 64, _p(f))
    if not ok:
        return None
    return gene.value.decode(), [int(x) if x >= 0 else None for x in f]
