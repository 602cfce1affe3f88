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
