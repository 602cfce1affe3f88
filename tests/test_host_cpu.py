"""CPU-side checks of the product library: it loads, exports every symbol include/hlacount.h declares, its host logic
(index loader/builder, packer, BAM reader) agrees with the oracle / an independent decoder, and compute entry points
refuse to run without a device (no CPU fallback). No GPU needed."""
import ctypes as C
import os

import numpy as np
import pytest

import helpers as H


@pytest.fixture(scope="module")
def sb():
    import schlacount_b200
    schlacount_b200.build()
    return schlacount_b200


def test_exports_every_declared_symbol(sb):
    from schlacount_b200 import _lib
    names = _lib.declared_symbols()
    assert len(names) >= 45
    L = _lib.lib()
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_index_build_and_idx_loader_match_fixture(sb, orc):
    idx = os.path.join(H.FIX, "fake_db", "hla_nuc.fasta.idx")
    o = orc.Index.from_idx(idx)
    a = sb.Index.from_idx(idx, device=-1)
    b = sb.Index.from_fasta(os.path.join(H.FIX, "fake_db", "hla_nuc.fasta"),
                            os.path.join(H.FIX, "fake_db", "Allele_status.txt"), device=-1)
    c = sb.Index.from_fasta(os.path.join(H.FIX, "cds_ABC.fa"), device=-1)
    assert a.nodes() == o.nodes()
    assert sorted(b.nodes()) == sorted(o.nodes()) == sorted(c.nodes())
    assert a.tx_names == b.tx_names == c.tx_names == o.tx_names
    info = a.info
    assert (info["n_nodes"], info["n_bases"], info["n_kmers"], info["n_colours"], info["n_tx"]) == (204, 8729, 4037, 39, 6)
    g = sb.Index.from_fasta(os.path.join(H.FIX, "genomic_ABC.fa"), device=-1)
    og = orc.Index.from_fasta(os.path.join(H.FIX, "genomic_ABC.fa"))
    assert sorted(g.nodes()) == sorted(og.nodes()) and g.info["n_kmers"] == 14237


def test_index_from_sequences_and_errors(sb, orc, tmp_path):
    from tools import synth
    cds, _ = synth.fixture_alleles()
    names = [h.split(" ")[1] for h, _ in cds]
    a = sb.Index.from_sequences([s for _, s in cds], names, device=-1)
    # from_sequences keeps the given order (no read_hla_cds sort): compare as node/colour-name sets
    o = orc.Index.from_fasta(os.path.join(H.FIX, "cds_ABC.fa"))
    def named(ix, nodes):
        tn = ix.tx_names
        return sorted((s, e, tuple(sorted(tn[t] for t in c))) for s, e, c in nodes)
    assert named(a, a.nodes()) == named(o, o.nodes())
    with pytest.raises(sb.HlacError):
        sb.Index.from_idx(tmp_path / "missing.idx", device=-1)
    bad = tmp_path / "bad.idx"
    bad.write_bytes(open(os.path.join(H.FIX, "fake_db", "hla_nuc.fasta.idx"), "rb").read()[:1000])
    with pytest.raises(sb.HlacError):
        sb.Index.from_idx(bad, device=-1)


def test_packer_matches_oracle(sb, orc, fixture_reads):
    seqs = [r[0] for r in fixture_reads[:200]] + ["", "ACGTN", "acgtnRYK" * 9, "T" * 97]
    a, la = sb.pack_reads(seqs)
    b, lb = orc.pack_reads(seqs)
    assert np.array_equal(a, b) and np.array_equal(la, lb)


def test_umi_key(sb):
    L = sb.lib()
    k = C.c_uint32()
    assert L.hlac_umi_key(b"ACGT", C.c_uint32(4), C.byref(k)) == 0 and k.value == (1 << 8) | 0b00011011
    assert k.value == H.umi_key(b"ACGT", {})
    assert L.hlac_umi_key(b"ACGN", C.c_uint32(4), C.byref(k)) != 0
    assert L.hlac_umi_key(b"A" * 16, C.c_uint32(16), C.byref(k)) != 0


def test_bam_reader_matches_independent_decoder(sb, fixture_reads):
    L = sb.lib()
    n, h = C.c_uint64(), C.c_uint64()
    bam = os.path.join(H.FIX, "test.bam").encode()
    assert L.hlac_bam_scan(bam, b"6:28510120-33480577", C.byref(n), C.byref(h)) == 0

    def checksum(reads):
        x = 0xcbf29ce484222325
        code = {"A": 0, "C": 1, "G": 2, "T": 3}
        for seq, cb, ub, primary in reads:
            for b in [code.get(c, 0) for c in seq] + [0xFF] + list(cb) + [0xFE] + list(ub) + [1 if primary else 0]:
                x = ((x ^ b) * 0x100000001b3) & 0xFFFFFFFFFFFFFFFF
        return x
    assert n.value == len(fixture_reads) == 2864
    assert h.value == checksum(fixture_reads)
    # a narrower region must select exactly the overlapping records, in order
    import bamlite
    sub = list(bamlite.fetch(os.path.join(H.FIX, "test.bam"), "6", 31268749, 31272092))
    assert 0 < len(sub) < 2864
    assert L.hlac_bam_scan(bam, b"6:31268749-31272092", C.byref(n), C.byref(h)) == 0
    assert n.value == len(sub) and h.value == checksum(sub)
    # unmapped tail of the fixture is empty
    assert L.hlac_bam_scan(bam, None, C.byref(n), C.byref(h)) == 0 and n.value == 0
    assert L.hlac_bam_scan(bam, b"chr6:1-2", C.byref(n), C.byref(h)) != 0   # '6' vs 'chr6' (mapping.rs:54)


def _batch_checksum(reads, pack):
    """what hlac_bam_batch_checksum hashes, from an independent decode: the batch H.make_batch builds"""
    seq, lens, cell, umi, flags = H.make_batch(reads, None, pack=pack)
    x = 0xcbf29ce484222325
    for i in range(len(lens)):
        nw = (int(lens[i]) + 31) // 32
        data = seq[i, :nw].astype("<u8").tobytes() + int(lens[i]).to_bytes(2, "little") + int(cell[i]).to_bytes(4, "little") \
            + int(umi[i]).to_bytes(4, "little") + bytes([int(flags[i])])
        for b in data:
            x = ((x ^ b) * 0x100000001b3) & 0xFFFFFFFFFFFFFFFF
    return x


def test_bam_batch_path_matches_independent_decoder(sb, fixture_reads, tmp_path):
    """The wrappers' ingest (records framed sequentially, parsed / packed on host threads, appended in file order) must
    hand the sessions exactly the batch an independent decode of the BAM gives: 2-bit words, lengths, first-seen barcode
    ids, UMI keys, flags - on the reference's fixture and on a synthetic BAM with dropped records and an unmapped tail."""
    import bamlite
    L = sb.lib()
    n, h = C.c_uint64(), C.c_uint64()
    bam = os.path.join(H.FIX, "test.bam").encode()
    for threads in ("1", "3"):
        os.environ["HLAC_BAM_THREADS"] = threads   # read when the reader / pool is created
        assert L.hlac_bam_batch_checksum(bam, b"6:28510120-33480577", C.byref(n), C.byref(h)) == 0, L.hlac_last_error()
        assert n.value == 2864 and h.value == _batch_checksum(fixture_reads, sb.pack_reads)
    os.environ.pop("HLAC_BAM_THREADS")
    sub = list(bamlite.fetch(os.path.join(H.FIX, "test.bam"), "6", 31268749, 31272092))
    assert L.hlac_bam_batch_checksum(bam, b"6:31268749-31272092", C.byref(n), C.byref(h)) == 0
    assert n.value == len(sub) and h.value == _batch_checksum(sub, sb.pack_reads)
    import subprocess
    import sys
    syn = str(tmp_path / "syn.bam")
    subprocess.check_call([sys.executable, os.path.join(H.ROOT, "tools", "make_bam.py"), syn, "6000", "700", "9000"], stdout=subprocess.DEVNULL)   # 9000 records (2.7 MB) on contig "1" come first and must be skipped
    mapped = list(bamlite.fetch(syn, "6", 28510120, 33480577))
    assert 5000 < len(mapped) < 6000                      # every 97th record has no UB and is dropped
    assert L.hlac_bam_batch_checksum(syn.encode(), b"6:28510120-33480577", C.byref(n), C.byref(h)) == 0, L.hlac_last_error()
    assert n.value == len(mapped) and h.value == _batch_checksum(mapped, sb.pack_reads)
    tail = list(bamlite.fetch_unmapped(syn))
    assert 600 < len(tail) < 700
    assert L.hlac_bam_batch_checksum(syn.encode(), None, C.byref(n), C.byref(h)) == 0
    assert n.value == len(tail) and h.value == _batch_checksum(tail, sb.pack_reads)


def test_no_cpu_fallback(sb):
    """Compute entry points must fail loudly without a CUDA device."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    cds = sb.Index.from_fasta(os.path.join(H.FIX, "cds_ABC.fa"), device=-1)
    with pytest.raises(sb.HlacError):
        sb.CountSession(cds, cds, 4)
    with pytest.raises(sb.HlacError):
        sb.GenoSession(cds)
    if not has_gpu:
        with pytest.raises(sb.HlacError):
            sb.Index.from_fasta(os.path.join(H.FIX, "cds_ABC.fa"), device=0)
        with pytest.raises(sb.HlacError):
            sb.squarem(4, {(0,): 1})


def test_locus_parsing_follows_the_reference_regex(sb):
    """src/locus.rs:29-48: ^(.*):([0-9,]+)(-|..)([0-9,]+)$ — commas allowed, '-' or ANY two characters as separator."""
    L = sb.lib()
    n, h = C.c_uint64(), C.c_uint64()
    bam = os.path.join(H.FIX, "test.bam").encode()
    for good in (b"6:28510120-33480577", b"6:28,510,120-33,480,577", b"6:28510120..33480577", b"6:28510120xy33480577"):
        assert L.hlac_bam_scan(bam, good, C.byref(n), C.byref(h)) == 0 and n.value == 2864, good
    for bad in (b"6", b"6:100", b"6:-200", b"6:100-", b"6:100-2x0", b"6:99999999999-5"):
        assert L.hlac_bam_scan(bam, bad, C.byref(n), C.byref(h)) != 0, bad
        assert b"invalid locus" in L.hlac_last_error()


def test_ctypes_layouts_match_the_library(sb):
    """The ctypes mirrors of the ABI structs must have the sizes the compiled header has (guards against drift)."""
    from schlacount_b200 import _lib
    out = (C.c_uint32 * 10)()
    assert sb.lib().hlac_abi_sizes(out, C.c_uint32(10)) == 0
    mine = [C.sizeof(x) for x in (_lib.Batch, _lib.Metrics, _lib.Coo, _lib.IndexInfo, _lib.ClassTables, _lib.Paths, _lib.Calls,
                                  _lib.EqClassDbC, _lib.PackedBatch, _lib.WireBatch)]
    assert list(out) == mine
    old = (C.c_uint32 * 7)()   # a binding built against the 7-value form still works
    assert sb.lib().hlac_abi_sizes(old, C.c_uint32(7)) == 0 and list(old) == mine[:7]


def test_eqclassdb_bincode_host(sb, tmp_path):
    """pseudoaligner.counts.bin in the reference's format (bincode EqClassDb, src/mapping.rs:84-90,403; src/em.rs:347):
    host reader/writer against an independent Python restatement of the format, incl. arbitrary HashMap order."""
    import random
    from schlacount_b200.api import EqClassDb
    bcs = {b"AAACCTGAGACCACGA-1": 1, b"TTTGTCATCTTGTATC-1": 0, b"": 2}
    umis = {b"ACGTACGTAC": 0, b"NNNNNNNNNN": 1}
    classes = {(3, 1, 2): 0, (5,): 1, (): 2, tuple(range(300)): 3}
    triples = [(1, 0, 0), (1, 0, 1), (0, 1, 3), (2, 1, 2), (1, 0, 0)]
    p1 = str(tmp_path / "a.bin")
    H.write_eqclassdb_bincode(p1, bcs, umis, classes, triples, rng=random.Random(5))
    db = EqClassDb.read(p1)
    assert db.barcodes == [b"TTTGTCATCTTGTATC-1", b"AAACCTGAGACCACGA-1", b""] and db.umis == [b"ACGTACGTAC", b"NNNNNNNNNN"]
    assert db.classes == [(3, 1, 2), (5,), (), tuple(range(300))]
    assert list(zip(db.bc.tolist(), db.umi.tolist(), db.cls.tolist())) == triples
    p2 = str(tmp_path / "b.bin")
    db.write(p2)
    assert H.read_eqclassdb_bincode(p2) == (bcs, umis, classes, triples)
    # an empty database
    H.write_eqclassdb_bincode(p1, {}, {}, {}, [])
    e = EqClassDb.read(p1)
    assert e.barcodes == [] and e.classes == [] and len(e.cls) == 0
    e.write(p2)
    assert open(p2, "rb").read() == open(p1, "rb").read()
    # malformed files are errors, not crashes
    raw = open(str(tmp_path / "b.bin"), "rb").read()
    H.write_eqclassdb_bincode(p1, bcs, umis, classes, triples)
    raw = open(p1, "rb").read()
    for bad in (raw[:-1], raw + b"\0", raw[:40], b"\xff" * 8 + raw[8:]):
        open(p2, "wb").write(bad)
        with pytest.raises(sb.HlacError):
            EqClassDb.read(p2)
    H.write_eqclassdb_bincode(p2, bcs, umis, classes, [(9, 0, 0)])      # barcode id out of range
    with pytest.raises(sb.HlacError):
        EqClassDb.read(p2)
    H.write_eqclassdb_bincode(p2, bcs, umis, {(1,): 0, (2,): 0}, [])    # duplicate class id
    with pytest.raises(sb.HlacError):
        EqClassDb.read(p2)
    # eq_class_counts is device work: no CPU fallback
    with pytest.raises(sb.HlacError, match="no CPU fallback"):
        db.counts(400, device=-1)


def test_idx_writer_reproduces_fixture(sb, tmp_path):
    """make_hla_index (src/hla.rs:230-264) end to end on the host: the fixture database's own FASTA + Allele_status.txt
    -> read_hla_cds -> build_index -> write_obj must give the committed test/fake_db/hla_nuc.fasta.idx byte for byte
    (node order, colour ids, exts, and the three boomphf 0.5.2 tables), and load -> write is the identity."""
    db = os.path.join(H.FIX, "fake_db")
    want = open(os.path.join(db, "hla_nuc.fasta.idx"), "rb").read()
    built = sb.Index.from_fasta(os.path.join(db, "hla_nuc.fasta"), os.path.join(db, "Allele_status.txt"), device=-1)
    p = str(tmp_path / "built.idx")
    built.write_idx(p)
    assert open(p, "rb").read() == want
    loaded = sb.Index.from_idx(os.path.join(db, "hla_nuc.fasta.idx"), device=-1)
    loaded.write_idx(p)
    assert open(p, "rb").read() == want
    # a different allele set: the written file must load back to the same graph
    other = sb.Index.from_fasta(os.path.join(H.FIX, "genomic_ABC.fa"), device=-1)
    other.write_idx(p)
    back = sb.Index.from_idx(p, device=-1)
    assert back.info == other.info and back.tx_names == other.tx_names
    key = lambda n: (n[0], n[1], tuple(n[2]))
    assert sorted(map(key, back.nodes())) == sorted(map(key, other.nodes()))


def test_pack_records_layout(sb):
    """hlac_pack_records: sequence words, then umi << 32 | primary << 31 | len << 23 | cell (include/hlacount.h)."""
    rng = np.random.default_rng(3)
    n, wpr = 50, 3
    seq = rng.integers(0, 1 << 63, size=(n, wpr), dtype=np.uint64)
    lens = rng.integers(0, 97, size=n).astype(np.uint16)
    cell = rng.integers(0, 5000, size=n).astype(np.uint32)
    cell[::7] = sb.NOT_A_CELL
    umi = rng.integers(0, 1 << 32, size=n, dtype=np.uint64).astype(np.uint32)
    flags = rng.integers(0, 2, size=n).astype(np.uint8)
    rec = sb.pack_records(seq, lens, cell, umi, flags)
    assert rec.shape == (n, wpr + 1) and np.array_equal(rec[:, :wpr], seq)
    c23 = np.where(cell == sb.NOT_A_CELL, sb.PACKED_NOT_A_CELL, cell).astype(np.uint64)
    meta = (umi.astype(np.uint64) << np.uint64(32)) | (flags.astype(np.uint64) << np.uint64(31)) | (lens.astype(np.uint64) << np.uint64(23)) | c23
    assert np.array_equal(rec[:, wpr], meta)
    assert sb.pack_records(seq[:0], lens[:0], cell[:0], umi[:0], flags[:0]).shape[0] == 0
    with pytest.raises(sb.HlacError):   # 8-bit length field
        sb.pack_records(np.zeros((1, 7), np.uint64), np.array([256], np.uint16), cell[:1], umi[:1], flags[:1])
    with pytest.raises(sb.HlacError):   # 23-bit cell field
        sb.pack_records(seq[:1], lens[:1], np.array([1 << 23], np.uint32), umi[:1], flags[:1])


def test_pack_wire_layout(sb):
    """hlac_pack_wire: a record is one big-endian bit stream - 2 * read_len bits of bases (the batch's own 2-bit codes), then
    primary, cell (all ones = not a cell), UMI key - checked against a bit-by-bit restatement."""
    rng = np.random.default_rng(9)
    for L, cb, ub in ((91, 14, 21), (91, 17, 21), (64, 5, 32), (33, 27, 9), (150, 20, 25)):
        n = 40
        seqs = ["".join("ACGT"[int(x)] for x in rng.integers(0, 4, size=L)) for _ in range(n)]
        seq, lens = sb.pack_reads(seqs)
        cell = rng.integers(0, (1 << cb) - 1, size=n).astype(np.uint32)
        cell[::7] = 0xFFFFFFFF
        umi = rng.integers(0, 1 << ub, size=n, dtype=np.uint64).astype(np.uint32)
        flags = rng.integers(0, 2, size=n).astype(np.uint8)
        rec = sb.pack_wire(seq, lens, cell, umi, flags, L, cb, ub)
        W = (2 * L + 1 + cb + ub + 31) // 32
        assert rec.shape == (n, W) and W == sb.lib().hlac_wire_words(L, cb, ub)
        for i in range(n):
            bits = "".join(format(int(w), "032b") for w in rec[i])
            want = "".join(format("ACGT".index(c), "02b") for c in seqs[i]) + str(int(flags[i] & 1)) + \
                format((1 << cb) - 1 if cell[i] == 0xFFFFFFFF else int(cell[i]), "0%db" % cb) + format(int(umi[i]), "0%db" % ub)
            assert bits[:len(want)] == want and set(bits[len(want):]) <= {"0"}
    assert sb.lib().hlac_wire_words(91, 14, 21) == 7   # 28 bytes per 91-bp read
    with pytest.raises(sb.HlacError):   # ragged batch
        sb.pack_wire(seq, np.where(np.arange(n) == 3, L - 1, lens).astype(np.uint16), cell, umi, flags, L, cb, ub)
    with pytest.raises(sb.HlacError):   # a cell equal to the not-a-cell sentinel
        sb.pack_wire(seq, lens, np.full(n, (1 << cb) - 1, np.uint32), umi, flags, L, cb, ub)


def test_bam_reader_rejects_damaged_files(sb, tmp_path):
    """Truncated or corrupted BAM / BAI input must surface as an error status from the ABI (serial and batch paths),
    never as a crash or a silently shorter result."""
    L = sb.lib()
    n, h = C.c_uint64(), C.c_uint64()
    src = os.path.join(H.FIX, "test.bam")
    raw, bai = open(src, "rb").read(), open(src + ".bai", "rb").read()
    region = b"6:28510120-33480577"

    def write(name, bam_bytes, bai_bytes=bai):
        p = str(tmp_path / name)
        open(p, "wb").write(bam_bytes)
        if bai_bytes is not None:
            open(p + ".bai", "wb").write(bai_bytes)
        return p.encode()
    for fn in (L.hlac_bam_scan, L.hlac_bam_batch_checksum):
        assert fn(write("ok.bam", raw), region, C.byref(n), C.byref(h)) == 0 and n.value == 2864
        assert fn(write("cut.bam", raw[:len(raw) // 2 + 7]), region, C.byref(n), C.byref(h)) != 0       # cut inside a BGZF block
        flipped = bytearray(raw); flipped[len(raw) // 3] ^= 0x5A                                         # damaged deflate stream: inflate or the block CRC32 catches it
        assert fn(write("flip.bam", bytes(flipped)), region, C.byref(n), C.byref(h)) != 0
        assert fn(write("noidx.bam", raw, None), region, C.byref(n), C.byref(h)) != 0 and b".bai" in L.hlac_last_error()
        assert fn(write("badidx.bam", raw, bai[:40]), region, C.byref(n), C.byref(h)) != 0
        assert fn(write("notbam.bam", b"\x1f\x8b" + b"\0" * 100), region, C.byref(n), C.byref(h)) != 0
        assert fn(write("ok.bam", raw), b"6:notanumber", C.byref(n), C.byref(h)) != 0                    # locus parse error (locus.rs)


def test_bam_aux_arrays_and_overruns(sb, tmp_path):
    """BAM aux fields the reader has to step over (SAM spec 4.2.4): 'B' arrays of every element width and htslib's 'd'
    must be skipped exactly; an array whose count is negative or runs past the record is a format error, not a mis-parse
    that picks up bogus CB/UB tags further on."""
    import struct
    from tools import make_bam as mb
    L = sb.lib()
    seq = "ACGT" * 22 + "ACG"

    def write(name, extras):
        p = str(tmp_path / name)
        w = mb.BgzfWriter(p)
        hdr = b"@HD\tVN:1.4\tSO:coordinate\n@SQ\tSN:6\tLN:170805979\n"
        w.write(b"BAM\x01" + struct.pack("<i", len(hdr)) + hdr + struct.pack("<i", 1) + struct.pack("<i", 2) + b"6\0" + struct.pack("<i", 170805979))
        first = w.tell()
        for i, ex in enumerate(extras):
            w.write(mb.record(0, 30_000_000 + 100 * i, 0, seq, b"r%d" % i, b"AAACCCGGGTTTACGT-1", b"ACGTACGTA" + b"ACGT"[i % 4:i % 4 + 1], extra_aux=ex))
        end = w.tell()
        w.close()
        n_intv = (30_000_000 + 100 * len(extras) >> 14) + 1
        with open(p + ".bai", "wb") as f:
            f.write(b"BAI\x01" + struct.pack("<i", 1) + struct.pack("<i", 1) + struct.pack("<Ii", 0, 1) + struct.pack("<QQ", first, end) + struct.pack("<i", n_intv))
            f.write(struct.pack("<%dQ" % n_intv, *([first] * n_intv)) + struct.pack("<Q", 0))
        return p.encode()
    good = [b"ZAB" + b"c" + struct.pack("<i", 5) + bytes([1, 2, 3, 0x43, 0x42]),              # bytes that look like a tag inside the array
            b"ZBB" + b"S" + struct.pack("<i", 3) + struct.pack("<3H", 1, 2, 3),
            b"ZCB" + b"f" + struct.pack("<i", 2) + struct.pack("<2f", 1.5, 2.5) + b"ZDd" + struct.pack("<d", 3.25),
            b"ZEB" + b"I" + struct.pack("<i", 0)]
    n, h = C.c_uint64(), C.c_uint64()
    region = b"6:28510120-33480577"
    for fn in (L.hlac_bam_scan, L.hlac_bam_batch_checksum):
        assert fn(write("good.bam", good), region, C.byref(n), C.byref(h)) == 0, L.hlac_last_error()
        assert n.value == len(good)
        for bad in (b"ZAB" + b"I" + struct.pack("<i", -1), b"ZAB" + b"I" + struct.pack("<i", 1 << 28), b"ZAB" + b"x" + struct.pack("<i", 1) + b"\0",
                    b"ZAB" + b"c"):
            assert fn(write("bad.bam", [good[0], bad]), region, C.byref(n), C.byref(h)) != 0
            assert b"aux" in L.hlac_last_error()


def test_index_build_equals_oracle_on_synthetic_db(sb, orc, tmp_path):
    """build_index on a 600-allele synthetic IMGT-like database: the product's sort-based host builder and
    the oracle's restatement must produce the same graph (node sequences, extension masks, colour lists) and tx order, and
    the written .idx must load back to it."""
    from tools import synth
    db = str(tmp_path / "db")
    os.makedirs(db)
    synth.make_db(db, 200, 31337)
    fa, st = os.path.join(db, "hla_nuc.fasta"), os.path.join(db, "Allele_status.txt")
    mine, ref = sb.Index.from_fasta(fa, st, device=-1), orc.Index.from_fasta(fa, st)
    assert mine.tx_names == ref.tx_names and mine.info["n_tx"] == 600
    key = lambda n: (n[0], n[1], tuple(n[2]))
    want = sorted(map(key, ref.nodes()))
    assert sorted(map(key, mine.nodes())) == want and len(want) == mine.info["n_nodes"] > 1000
    p = str(tmp_path / "x.idx")
    mine.write_idx(p)
    assert sorted(map(key, sb.Index.from_idx(p, device=-1).nodes())) == want
    assert sorted(map(key, orc.Index.from_idx(p).nodes())) == want      # the oracle's own bincode reader accepts the file


def test_header_is_plain_c(tmp_path):
    """include/hlacount.h is the drop-in boundary: it must be consumable by a C compiler (and so by bindgen / cgo-style
    tools), not only by the C++ that implements it."""
    import subprocess
    src = tmp_path / "use_header.c"
    src.write_text('#include "hlacount.h"\nint main(void) { hlac_batch b; hlac_packed_batch p; hlac_eqclassdb d; (void)b; (void)p; (void)d;\n'
                   '  return (int)(HLAC_PACKED_META(1u, 1, 91, 5) >> 63); }\n')
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(H.ROOT, "include"),
                           "-fsyntax-only", str(src)])


def test_inflate_decoder_against_zlib(tmp_path):
    """csrc/inflate_fast.cpp (the BAM reader's own DEFLATE decoder) against zlib: tests/native/inflate_check.cpp deflates a dozen
    kinds of data at every level / strategy / memory level and with mid-stream flushes, and takes the BGZF blocks of the
    fixture BAMs; every stream must decode to the identical bytes (a few may be declined, never wrong), damaged and
    truncated streams must not write outside their buffers. Built with ASan + UBSan when the toolchain has them."""
    import subprocess
    exe = str(tmp_path / "inflate_check")
    base = ["g++", "-O1", "-g", "-std=c++17", "-Wall", os.path.join(H.ROOT, "tests", "native", "inflate_check.cpp"),
            os.path.join(H.ROOT, "schlacount_b200", "csrc", "inflate_fast.cpp"), "-lz", "-o", exe]
    if subprocess.call(base[:5] + ["-fsanitize=address,undefined"] + base[5:], stderr=subprocess.DEVNULL) != 0:
        subprocess.check_call(base)
    bams = [os.path.join(H.FIX, f) for f in sorted(os.listdir(H.FIX)) if f.endswith(".bam")]
    assert bams
    out = subprocess.run([exe] + bams, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.startswith("ok:") and " 0 declined" in out.stdout, out.stdout + out.stderr


def test_bam_reader_same_records_with_either_inflate(sb, monkeypatch):
    """the reader's result does not depend on which decoder inflates the blocks (HLAC_ZLIB_INFLATE=1 selects zlib; read once
    per process, so the comparison runs in two subprocesses)"""
    import subprocess
    import sys
    code = ("import ctypes as C, sys; sys.path.insert(0, %r); from schlacount_b200._lib import lib; L = lib(); n, h = C.c_uint64(), C.c_uint64();"
            "assert L.hlac_bam_scan(%r, b'6:28510120-33480577', C.byref(n), C.byref(h)) == 0; print(n.value, h.value)"
            % (H.ROOT, os.path.join(H.FIX, "test.bam").encode()))
    outs = []
    for zl in ("", "1"):
        env = dict(os.environ)
        env.pop("HLAC_ZLIB_INFLATE", None)
        if zl:
            env["HLAC_ZLIB_INFLATE"] = zl
        outs.append(subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300).stdout.strip())
    assert outs[0] == outs[1] and outs[0].split()[0] == "2864", outs


def _caller_restated(names, theta, reads_explained, counts_reads, nreads):
    """src/em.rs:361-434 restated directly from the Rust, independently of csrc/caller.hpp and of the oracle's call_genotypes
    (a third implementation: the other two were written by the same hand). -> (called tx ids sorted, weights rows, pairs rows)"""
    import itertools
    wn = [(i, w, names[i], names[i].split("*")[0], reads_explained[i]) for i, w in enumerate(theta)]
    wn.sort(key=lambda x: -x[1])              # sort_by(-wa partial_cmp -wb): stable
    wn.sort(key=lambda x: x[3].encode())      # sort_by_key(gene): stable wrt weight
    called, weights_rows, pairs_rows = set(), [], []
    for _, grp in itertools.groupby(wn, key=lambda x: x[3]):
        ws = list(grp)
        written = 0
        for (i, w, _, _, exp) in ws:
            if exp > 100:                     # MIN_READS_CALL
                weights_rows.append((i, w, float(exp) / float(nreads)))
                written += 1
                if written == 10:             # WEIGHTS_TO_OUTPUT
                    break
            else:
                break
        if written == 0:
            continue
        pairs, stop = [], False
        for i in range(written - 1):
            for j in range(i + 1, written):
                a, b = ws[i][0], ws[j][0]
                exp = sum(c for cls, c in counts_reads.items() if a in cls or b in cls)   # pair_reads_explained (em.rs:36-44)
                if exp > 100:
                    pairs.append((a, b, float(exp) / float(nreads), float(max(ws[i][4], ws[j][4])) / float(nreads)))
                else:
                    stop = True               # break 'outer
                    break
            if stop:
                break
        if pairs:
            pairs.sort(key=lambda p: -p[2])
            pairs_rows += [(a, b, f) for a, b, f, _ in pairs[:5]]   # PAIRS_TO_OUTPUT
            if pairs[0][2] * 0.15 > pairs[0][2] - pairs[0][3]:      # HOMOZYGOUS_TH
                called |= {pairs[0][0], pairs[0][1]}
            else:
                called.add(ws[0][0])
        else:
            called.add(ws[0][0])
    return sorted(called), weights_rows, pairs_rows


def test_caller_against_an_independent_restatement(sb):
    """hlac_call_genotypes (csrc/caller.hpp) on random class tables vs a restatement of em.rs:361-434 written from the Rust
    alone: weights.tsv rows, pairs.tsv rows and the called set must be identical (ties, the 100-read cut-offs, the
    break-outer on the first failing pair, the homozygosity rule and the 10 / 5 row limits all occur)."""
    from schlacount_b200 import _lib
    rng = np.random.default_rng(12)
    genes = ["A", "B", "C", "DRB1", "DQA1"]
    names = ["%s*%02d:%02d" % (g, 1 + k // 4, 1 + k % 4) for g in genes for k in range(14)]
    n = len(names)
    seqs = ["".join("ACGT"[int(x)] for x in rng.integers(0, 4, size=40)) for _ in names]
    ix = sb.Index.from_sequences(seqs, names, device=-1)   # host-only handle: the caller needs the tx names only
    assert ix.tx_names == names
    n_pairs_seen = n_both = n_single = 0
    for trial in range(60):
        classes = {}
        for _ in range(int(rng.integers(5, 60))):
            g = int(rng.integers(0, len(genes)))
            members = rng.choice(14, size=int(rng.integers(1, 6)), replace=False) + 14 * g
            if rng.random() < 0.1:
                members = np.concatenate([members, rng.choice(n, size=2, replace=False)])   # a class across genes
            key = tuple(int(x) for x in dict.fromkeys(rng.permutation(members).tolist()))   # unsorted, like a read-level class
            classes[key] = classes.get(key, 0) + int(rng.integers(1, 120 if trial % 3 else 400))
        nreads = sum(classes.values())
        expl = [sum(c for k, c in classes.items() if t in k) for t in range(n)]
        theta = rng.random(n)
        theta[rng.random(n) < 0.3] = 0.0
        if trial % 5 == 0:
            theta[:8] = theta[0]                                                   # ties in the weights
        theta = theta / theta.sum()
        keys = list(classes)
        off = np.zeros(len(keys) + 1, np.uint64)
        for i, k in enumerate(keys):
            off[i + 1] = off[i] + len(k)
        tx = np.asarray([t for k in keys for t in k], np.uint32)
        cnt = np.asarray([classes[k] for k in keys], np.uint32)
        ex = np.asarray(expl, np.uint64)
        t = _lib.ClassTables()
        t.nitems, t.nreads, t.n_read_classes = n, nreads, len(keys)
        t.reads_off, t.reads_tx, t.reads_cnt = (off.ctypes.data_as(C.POINTER(C.c_uint64)), tx.ctypes.data_as(C.POINTER(C.c_uint32)),
                                                cnt.ctypes.data_as(C.POINTER(C.c_uint32)))
        t.reads_explained = ex.ctypes.data_as(C.POINTER(C.c_uint64))
        th = np.ascontiguousarray(theta, np.float64)
        c = _lib.Calls()
        assert sb.lib().hlac_call_genotypes(ix.h, C.byref(t), th.ctypes.data_as(C.c_void_p), C.byref(c)) == 0, sb.lib().hlac_last_error()
        got = ([c.called[i] for i in range(c.n_called)], [(c.w_tx[i], c.w_em[i], c.w_frac[i]) for i in range(c.n_weights)],
               [(c.p_a[i], c.p_b[i], c.p_frac[i]) for i in range(c.n_pairs)])
        sb.lib().hlac_calls_free(C.byref(c))
        want = _caller_restated(names, theta.tolist(), expl, classes, nreads)
        assert got[0] == want[0] and got[1] == want[1] and got[2] == want[2], (trial, got, want)
        n_pairs_seen += len(want[2])
        n_both += sum(1 for _ in want[0])
    assert n_pairs_seen > 50 and n_both > 100
