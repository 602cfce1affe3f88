"""Shared test helpers: fixture paths, reference-semantics host bits restated in Python for the tests."""
import gzip
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
FIX = os.path.join(ROOT, "tests", "golden", "ref_fixtures")
DEFAULT_REGION = ("6", 28510120, 33480577)  # src/main.rs:118

NOT_A_CELL = 0xFFFFFFFF


def load_barcodes(path):
    """src/main.rs:406-424: column = line index (last occurrence wins), gz sniffed (src/io.rs:9-27)."""
    raw = open(path, "rb").read()
    if raw[:2] == b"\x1f\x8b":
        raw = gzip.decompress(raw)
    lines = raw.split(b"\n")
    if lines and lines[-1] == b"":
        lines.pop()  # BufRead::lines does not yield a final empty piece after the trailing newline
    return {l.rstrip(b"\r") if l.endswith(b"\r") else l: i for i, l in enumerate(lines)}


def umi_key(umi, table):
    """Injective u32 key for a UMI byte string: 2-bit packed with a length sentinel when it is <=15 ACGT bases,
    otherwise an interned id with bit 31 set (mirrors the product host packer, include/hlacount.h)."""
    if len(umi) <= 15 and all(c in b"ACGT" for c in umi):
        v = 1
        for c in umi:
            v = (v << 2) | b"ACGT".index(c)
        return v
    if umi not in table:
        table[umi] = 0x80000000 | len(table)
    return table[umi]


def read_mtx(path):
    rows = [l.split() for l in open(path) if not l.startswith("%")]
    nr, nc, nnz = map(int, rows[0])
    ent = sorted((int(a) - 1, int(b) - 1, int(c)) for a, b, c in rows[1:])
    assert len(ent) == nnz
    return nr, nc, ent


def fixture_reads():
    """All CB+UB records of test.bam in the default region, file order: (seq, cb, ub, primary)."""
    import bamlite
    return list(bamlite.fetch(os.path.join(FIX, "test.bam"), *DEFAULT_REGION))


def make_batch(reads, barcodes=None, pack=None):
    """-> (seq words [n,wpr] u64, len u16, cell u32, umi u32, flags u8). With `barcodes` (dict cb->column) cell is
    the column or NOT_A_CELL; without, cell is a first-seen barcode id (genotyping uses every barcode)."""
    if pack is None:
        from oracle import oracle as orc
        pack = orc.pack_reads
    seq, lens = pack([r[0] for r in reads])
    cell = np.zeros(len(reads), dtype=np.uint32)
    umi = np.zeros(len(reads), dtype=np.uint32)
    flags = np.zeros(len(reads), dtype=np.uint8)
    seen, ut = {}, {}
    for i, (_, cb, ub, primary) in enumerate(reads):
        if barcodes is not None:
            cell[i] = barcodes.get(cb, NOT_A_CELL)
        else:
            cell[i] = seen.setdefault(cb, len(seen))
        umi[i] = umi_key(ub, ut)
        flags[i] = 1 if primary else 0
    return seq, lens, cell, umi, flags


# ---- the reference's pseudoaligner.counts.bin: bincode 1.x of EqClassDb (src/mapping.rs:84-90), restated for the tests
def write_eqclassdb_bincode(path, barcodes, umis, classes, triples, rng=None):
    """barcodes / umis: {bytes: id}; classes: {tuple of tx ids: id}; triples: iterable of (barcode_id, umi_id, class_id).
    HashMap entries are written in a shuffled order when `rng` is given (the reference's iteration order is arbitrary)."""
    import struct

    def items(d):
        it = list(d.items())
        if rng is not None:
            rng.shuffle(it)
        return it
    out = bytearray()
    for d in (barcodes, umis):
        out += struct.pack("<Q", len(d))
        for k, v in items(d):
            out += struct.pack("<Q", len(k)) + bytes(k) + struct.pack("<I", v)
    out += struct.pack("<Q", len(classes))
    for k, v in items(classes):
        out += struct.pack("<Q", len(k)) + struct.pack("<%dI" % len(k), *k) + struct.pack("<I", v)
    triples = list(triples)
    out += struct.pack("<Q", len(triples))
    for t in triples:
        out += struct.pack("<3I", *t)
    open(path, "wb").write(bytes(out))


def read_eqclassdb_bincode(path):
    """-> (barcodes {bytes: id}, umis {bytes: id}, classes {tuple: id}, triples [(bc, umi, cls)]); asserts EOF."""
    import struct
    raw = open(path, "rb").read()
    pos = 0

    def take(fmt):
        nonlocal pos
        v = struct.unpack_from("<" + fmt, raw, pos)
        pos += struct.calcsize("<" + fmt)
        return v
    maps = []
    for _ in range(2):
        d = {}
        for _ in range(take("Q")[0]):
            n = take("Q")[0]
            k = raw[pos:pos + n]
            pos += n
            d[k] = take("I")[0]
        maps.append(d)
    classes = {}
    for _ in range(take("Q")[0]):
        n = take("Q")[0]
        k = take("%dI" % n)
        classes[tuple(k)] = take("I")[0]
    triples = [take("3I") for _ in range(take("Q")[0])]
    assert pos == len(raw), "trailing bytes"
    return maps[0], maps[1], classes, triples


# ---- exact comparison of large class tables (CSR) without building Python tuples ---------------------------------------
def csr_canonical(off, tx, cnt):
    """(off, tx, cnt) of a {class vector: count} table -> (fingerprint rows sorted, flat tx in that class order): two tables
    are equal as maps from ORDERED vectors to counts iff both parts are equal."""
    off = np.asarray(off, dtype=np.int64)
    tx = np.asarray(tx, dtype=np.uint64)
    cnt = np.asarray(cnt, dtype=np.uint64)
    n = len(off) - 1
    if n == 0:
        return np.zeros((0, 3), np.uint64), np.zeros(0, np.uint64)
    lens = np.diff(off)
    assert (lens > 0).all()
    pos = (np.arange(len(tx), dtype=np.int64) - np.repeat(off[:-1], lens)).astype(np.uint64)
    w = (pos * np.uint64(0x9E3779B97F4A7C15) + np.uint64(0xD1B54A32D192ED03)) | np.uint64(1)
    with np.errstate(over="ignore"):
        h = np.add.reduceat((tx + np.uint64(1)) * w, off[:-1])
    order = np.lexsort((cnt, lens, h))
    rows = np.stack([h[order], lens[order].astype(np.uint64), cnt[order]], 1)
    lo = lens[order]
    starts = off[:-1][order]
    idx = np.repeat(starts - np.concatenate([[0], np.cumsum(lo)[:-1]]), lo) + np.arange(int(lo.sum()), dtype=np.int64)
    return rows, tx[idx]


def csr_tables_equal(a, b):
    ra, fa = csr_canonical(*a)
    rb, fb = csr_canonical(*b)
    return ra.shape == rb.shape and bool((ra == rb).all()) and np.array_equal(fa, fb)


def em_loglik(off, tx, cnt, theta):
    """EqClassCounts::likelihood (src/em.rs:119-137): sum over classes of count * ln(sum of theta over the class)"""
    off = np.asarray(off, dtype=np.int64)
    s = np.add.reduceat(np.asarray(theta, dtype=np.float64)[np.asarray(tx, dtype=np.int64)], off[:-1])
    return float((np.asarray(cnt, dtype=np.float64) * np.log(s)).sum())
