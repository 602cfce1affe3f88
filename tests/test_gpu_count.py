"""GPU parity for the counting path: CUDA kernels (through the C ABI) vs the CPU oracle and the reference golden.
Replaces map_and_count_pseudo's hot loop + count() (src/mapping.rs:740-909)."""
import os

import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sb():
    import schlacount_b200
    return schlacount_b200


@pytest.fixture(scope="module")
def abc(sb, orc):
    cf, gf = os.path.join(H.FIX, "cds_ABC.fa"), os.path.join(H.FIX, "genomic_ABC.fa")
    return dict(gc=sb.Index.from_fasta(cf), gg=sb.Index.from_fasta(gf), oc=orc.Index.from_fasta(cf),
                og=orc.Index.from_fasta(gf))


def _run_gpu(sb, abc, batch, n_cells, primary_only=False, splits=1):
    s = sb.CountSession(abc["gc"], abc["gg"], n_cells, primary_only)
    s.record_codes(True)
    seq, lens, cell, umi, flags = batch
    n = len(lens)
    codes = []
    bounds = np.linspace(0, n, splits + 1).astype(int)
    for a, b in zip(bounds[:-1], bounds[1:]):
        s.push(seq[a:b], lens[a:b], cell[a:b], umi[a:b], flags[a:b])
        if b > a:
            codes.append(s.last_codes(b - a))
    ent, met = s.finish()
    return ent, met, (np.concatenate(codes) if codes else np.zeros(0, np.uint8)), s


def test_golden_allele_fasta_gpu(sb, orc, abc, fixture_reads):
    """src/main.rs:432-466 through the CUDA path: count matrix == test/test_allele_fasta.mtx, per-read codes == oracle."""
    bc = H.load_barcodes(os.path.join(H.FIX, "barcodes1.tsv"))
    batch = H.make_batch(fixture_reads, bc, pack=sb.pack_reads)
    ent, met, codes, s = _run_gpu(sb, abc, batch, len(bc))
    nr, nc, gold = H.read_mtx(os.path.join(H.FIX, "test_allele_fasta.mtx"))
    assert s.nrows == nr and sorted(ent) == gold
    ref = orc.count_run(abc["oc"], abc["og"], *batch, want_codes=True)
    assert ent == ref["entries"] and met == ref["metrics"]
    assert np.array_equal(codes, ref["codes"])
    assert s.row_names == abc["oc"].rows()


def test_fixture_all_barcodes_gpu(sb, orc, abc, fixture_reads):
    bc = H.load_barcodes(os.path.join(H.FIX, "barcodes7.tsv"))
    batch = H.make_batch(fixture_reads, bc, pack=sb.pack_reads)
    for primary_only in (False, True):
        ent, met, codes, _ = _run_gpu(sb, abc, batch, len(bc), primary_only, splits=3)
        ref = orc.count_run(abc["oc"], abc["og"], *batch, primary_only=primary_only, want_codes=True)
        assert ent == ref["entries"] and met == ref["metrics"] and np.array_equal(codes, ref["codes"])


@pytest.mark.parametrize("n,cells,seed", [(200_000, 1000, 7), (50_000, 3, 8)])
def test_synthetic_parity(sb, orc, abc, n, cells, seed):
    from tools import synth
    cds, gen = synth.fixture_alleles()
    r = synth.make_reads(n, [s for _, s in cds], [s for _, s in gen], cells, seed)
    batch = (r["seq"], r["len"], r["cell"], r["umi"], r["flags"])
    ent, met, codes, s = _run_gpu(sb, abc, batch, cells, splits=2)
    ref = orc.count_run(abc["oc"], abc["og"], *batch, want_codes=True)
    assert np.array_equal(codes, ref["codes"])
    assert met == ref["metrics"]
    assert ent == ref["entries"]
    assert sum(v for _, _, v in ent) > 0


def test_edge_cases(sb, orc, abc):
    """empty batch, reads shorter than K, ragged lengths in one batch, N bases, everything filtered."""
    cds, gen = __import__("tools.synth", fromlist=["x"]).fixture_alleles()
    a = cds[0][1]
    g = gen[3][1]
    rc = lambda s: s[::-1].translate(str.maketrans("ACGT", "TGCA"))
    seqs = [a[:23], a[:24], a[100:250], rc(a[300:420]), g[1500:1640], "ACGT" * 30, a[10:101].replace("G", "N", 2),
            a[500:530] + "T" + a[531:600], "", a[700:760] + g[2000:2040]]
    reads = [(s, b"AAACCTGAGACCACGA-1" if i % 3 else b"XX", b"ACGTACGTAC" if i % 2 else b"TTTTTTTTTT", i != 4)
             for i, s in enumerate(seqs)]
    bc = {b"AAACCTGAGACCACGA-1": 0, b"XX": 1}
    batch = H.make_batch(reads, bc, pack=sb.pack_reads)
    assert batch[0].shape[1] == 5
    ent, met, codes, s = _run_gpu(sb, abc, batch, 2)
    ref = orc.count_run(abc["oc"], abc["og"], *batch, want_codes=True)
    assert np.array_equal(codes, ref["codes"]) and met == ref["metrics"] and ent == ref["entries"]
    # empty session / empty batch
    s2 = sb.CountSession(abc["gc"], abc["gg"], 2)
    s2.push(batch[0][:0], batch[1][:0], batch[2][:0], batch[3][:0], batch[4][:0])
    ent2, met2 = s2.finish()
    assert ent2 == [] and met2 == [0] * 6
    # nothing in the barcode list
    none = (batch[0], batch[1], np.full(len(seqs), H.NOT_A_CELL, np.uint32), batch[3], batch[4])
    ent3, met3, codes3, _ = _run_gpu(sb, abc, none, 2)
    assert ent3 == [] and met3[2] == len(seqs) and (codes3 == 255).all()


def test_reset_and_repeat(sb, abc, fixture_reads):
    bc = H.load_barcodes(os.path.join(H.FIX, "barcodes7.tsv"))
    batch = H.make_batch(fixture_reads, bc, pack=sb.pack_reads)
    s = sb.CountSession(abc["gc"], abc["gg"], len(bc))
    s.push(*batch)
    e1, m1 = s.finish()
    s.reset()
    s.push(*batch)
    e2, m2 = s.finish()
    assert e1 == e2 and m1 == m2
    t = s.timing()
    assert t["launches"] >= 2 and t["map_ms"] > 0


def test_device_memory_cache(sb, abc, fixture_reads):
    """Device buffers are recycled through the library's cache of freed blocks (cuda_util.hpp): a closed session's memory is
    handed to the next one, hlac_device_cache_trim gives it back to the driver, and results do not depend on either."""
    bc = H.load_barcodes(os.path.join(H.FIX, "barcodes7.tsv"))
    batch = H.make_batch(fixture_reads, bc, pack=sb.pack_reads)
    s = sb.CountSession(abc["gc"], abc["gg"], len(bc))
    s.push(*batch)
    e1, m1 = s.finish()
    s.close()
    released, live = sb.device_cache_trim()
    assert released > 0 and live > 0          # the session's blocks were cached; the two indexes are still alive
    assert sb.device_cache_trim()[0] == 0     # nothing left to release
    s = sb.CountSession(abc["gc"], abc["gg"], len(bc))   # fresh blocks from the driver
    s.push(*batch)
    e2, m2 = s.finish()
    s.close()
    s = sb.CountSession(abc["gc"], abc["gg"], len(bc))   # recycled blocks
    s.push(*batch)
    e3, m3 = s.finish()
    assert e1 == e2 == e3 and m1 == m2 == m3


def test_abi_argument_errors(sb, abc):
    """Errors surface as negative status + message, never as a crash or a silent fallback."""
    import ctypes as C
    L = sb.lib()
    assert L.hlac_count_new(None, None, C.c_uint32(1), C.c_int(0), C.byref(C.c_void_p())) != 0
    assert b"null" in L.hlac_last_error()
    with pytest.raises(sb.HlacError):
        sb.CountSession(abc["gc"], abc["gg"], 0)
    s = sb.CountSession(abc["gc"], abc["gg"], 4)
    with pytest.raises(sb.HlacError):        # entries before reduce
        s.entries()
    with pytest.raises(sb.HlacError):        # codes were not being recorded
        s.last_codes(1)
    # a cell index >= n_cells is a caller error and must be reported, not written out of bounds
    seq, lens = sb.pack_reads(["ACGT" * 23])
    s.push(seq, lens, np.array([7], np.uint32), np.array([1], np.uint32), np.array([1], np.uint8))
    with pytest.raises(sb.HlacError, match="cell index"):
        s.finish()


def test_long_reads_and_interned_umis(sb, orc, abc):
    """150/250 bp reads (wpr 5 and 8, smaller shared-memory tiles) and UMI keys with bit 31 set."""
    from tools import synth
    cds, gen = synth.fixture_alleles()
    rng = np.random.default_rng(5)
    rc = lambda s: s[::-1].translate(str.maketrans("ACGT", "TGCA"))
    for L in (150, 250):
        reads = []
        for i in range(3000):
            src = (cds if rng.random() < 0.7 else gen)[int(rng.integers(0, 6))][1]
            st = int(rng.integers(0, len(src) - L))
            s = list(src[st:st + L])
            for p in rng.integers(0, L, size=int(rng.integers(0, 4))):
                s[p] = "ACGT"[int(rng.integers(0, 4))]
            s = "".join(s)
            reads.append((rc(s) if rng.random() < 0.5 else s, b"c%d" % int(rng.integers(0, 9)),
                          b"NNACGTNN%d" % int(rng.integers(0, 40)), True))
        bc = {b"c%d" % i: i for i in range(9)}
        batch = H.make_batch(reads, bc, pack=sb.pack_reads)
        assert (batch[3] >> 31).all()
        ent, met, codes, _ = _run_gpu(sb, abc, batch, 9)
        ref = orc.count_run(abc["oc"], abc["og"], *batch, want_codes=True)
        assert np.array_equal(codes, ref["codes"]) and met == ref["metrics"] and ent == ref["entries"]
        assert len(ent) > 20


def test_row_layout_many_genes(sb, orc, tmp_path):
    """mapping.rs:668-728 beyond the A/B/C fixture: 8 genes (config 4 shape), a gene with a single allele (1 row), a gene
    with three alleles (the third is in the index but has no row), 100 k barcodes; reads from every allele."""
    from tools import synth
    db = str(tmp_path / "db")
    names = synth.make_db(db, 3, 77, extra_genes=("DRB1", "DPA1", "DPB1", "DQA1", "DQB1"))
    nuc = dict((h.split(" ")[1], (h, s)) for h, s in synth.read_fasta(os.path.join(db, "hla_nuc.fasta")))
    gen = dict((h.split(" ")[1], (h, s)) for h, s in synth.read_fasta(os.path.join(db, "hla_gen.fasta")))
    keep = []
    for g in ("A", "B", "C", "DRB1", "DPA1", "DPB1", "DQA1", "DQB1"):
        al = [n for n in names if n.startswith(g + "*")]
        keep += al[:1] if g == "C" else al[:3] if g == "B" else al[:2]
    cf, gf = str(tmp_path / "cds.fa"), str(tmp_path / "gen.fa")
    synth.write_fasta(cf, [nuc[a] for a in keep])
    synth.write_fasta(gf, [gen[a] for a in keep])
    gc, gg = sb.Index.from_fasta(cf), sb.Index.from_fasta(gf)
    oc, og = orc.Index.from_fasta(cf), orc.Index.from_fasta(gf)
    rows = oc.rows()
    assert len(rows) == 3 * 7 + 1 and rows.count("C") == 0 and "B" in rows
    n_cells = 100_000
    r = synth.make_reads(150_000, [nuc[a][1] for a in keep], [gen[a][1] for a in keep], n_cells, 78)
    batch = (r["seq"], r["len"], r["cell"], r["umi"], r["flags"])
    s = sb.CountSession(gc, gg, n_cells)
    assert s.row_names == rows and s.nrows == 22
    s.record_codes(True)
    s.push(*batch)
    codes = s.last_codes(len(r["len"]))
    ent, met = s.finish()
    ref = orc.count_run(oc, og, *batch, want_codes=True)
    assert np.array_equal(codes, ref["codes"]) and met == ref["metrics"] and ent == ref["entries"]
    used = set(np.unique(codes).tolist()) - {255}
    assert len(used) >= 15 and max(used) <= 23    # alleles and gene-level rows of many genes were hit


@pytest.mark.parametrize("cells", [64, 2, 1, 3000])
def test_deep_umis(sb, orc, abc, cells):
    """count() (src/mapping.rs:823-909) with many reads per UMI: few distinct UMIs per cell, so that majority votes, ties,
    the 10 % allele rule and UMIs without any winning gene all occur; also reduce -> push -> reduce on one session."""
    from tools import synth
    cds, gen = synth.fixture_alleles()
    r = synth.make_reads(40000, [s for _, s in cds], [s for _, s in gen], cells, 99 + cells, err=0.01)
    rng = np.random.default_rng(5)
    r["umi"] = (r["umi"] & 0x3F) | 0x100000
    r["cell"][rng.random(len(r["cell"])) < 0.02] = H.NOT_A_CELL
    batch = (r["seq"], r["len"], r["cell"], r["umi"], r["flags"])
    ent, met, codes, s = _run_gpu(sb, abc, batch, cells, splits=3)
    ref = orc.count_run(abc["oc"], abc["og"], *batch, want_codes=True)
    assert np.array_equal(codes, ref["codes"]) and met == ref["metrics"]
    assert ent == ref["entries"] and (len(ent) > 0 or cells < 64)
    # reduce again after another push: counts accumulate over the whole session
    s.push(*batch)
    ent2, met2 = s.finish()
    both = tuple(np.concatenate([x, x]) for x in batch)
    ref2 = orc.count_run(abc["oc"], abc["og"], *both)
    assert ent2 == ref2["entries"] and met2 == ref2["metrics"]


def test_packed_records_equal_soa(sb, orc, abc, fixture_reads):
    """hlac_count_push_packed (one 32-byte record per read, one H2D copy per push) must give exactly what the five-array
    push gives: fixture reads with --primary-alignments semantics, synthetic reads in three pushes, host and device buffers."""
    import torch
    from tools import synth
    bc = H.load_barcodes(os.path.join(H.FIX, "barcodes7.tsv"))
    batch = H.make_batch(fixture_reads, bc, pack=sb.pack_reads)
    cds, gen = synth.fixture_alleles()
    r = synth.make_reads(60000, [s for _, s in cds], [s for _, s in gen], 300, 21)
    for (b, ncell, primary_only) in ((batch, len(bc), True), (batch, len(bc), False),
                                     ((r["seq"], r["len"], r["cell"], r["umi"], r["flags"]), 300, False)):
        ref = orc.count_run(abc["oc"], abc["og"], *b, primary_only=primary_only, want_codes=True)
        rec = sb.pack_records(*b)
        for device in (False, True):
            s = sb.CountSession(abc["gc"], abc["gg"], ncell, primary_only)
            s.record_codes(True)
            codes = []
            for part in np.array_split(np.arange(len(rec)), 3):
                chunk = rec[part[0]:part[-1] + 1]
                s.push_packed(torch.from_numpy(chunk.view(np.int64)).cuda() if device else chunk)
                codes.append(s.last_codes(len(chunk)))
            ent, met = s.finish()
            assert np.array_equal(np.concatenate(codes), ref["codes"]) and met == ref["metrics"] and ent == ref["entries"]
    # 23-bit cell field: sessions with more columns must use the five-array push
    big = sb.CountSession(abc["gc"], abc["gg"], (1 << 23))
    with pytest.raises(sb.HlacError, match="23-bit"):
        big.push_packed(rec[:10])


def test_bucket_overflow_takes_the_sort_path(sb, orc, abc):
    """The grouping kernel holds one hash bucket (2048 keys) in shared memory; a molecule deeper than that (here every read
    of the batch carries the same barcode and UMI) overflows its bucket and the step is redone through the gather + radix
    sort + run-consensus path. Results must not depend on which path ran; a normal batch must not take the fallback."""
    from tools import synth
    cds, gen = synth.fixture_alleles()
    r = synth.make_reads(60000, [s for _, s in cds], [s for _, s in gen], 4, 123, err=0.01)
    r["umi"][:] = 0x1ABCDE
    r["cell"][:] = 2
    r["cell"][::50] = H.NOT_A_CELL
    batch = (r["seq"], r["len"], r["cell"], r["umi"], r["flags"])
    ent, met, codes, s = _run_gpu(sb, abc, batch, 4, splits=3)
    ref = orc.count_run(abc["oc"], abc["og"], *batch, want_codes=True)
    assert np.array_equal(codes, ref["codes"]) and met == ref["metrics"] and ent == ref["entries"]
    st = s.group_stats()
    assert st["fallbacks"] == 1 and st["keys"] == int((codes != 255).sum()) > 2048
    # the legacy two-call form (reduce, then entries) takes the same fallback
    s.reset()
    s.push(*batch)
    s.reduce()
    assert s.entries() == (ref["entries"], ref["metrics"]) and s.group_stats()["fallbacks"] == 2
    # mixed: one deep molecule among ordinary ones; and an ordinary batch stays on the fused path
    r2 = synth.make_reads(120000, [s for _, s in cds], [s for _, s in gen], 40, 124)
    b2 = tuple(np.concatenate([x, y]) for x, y in zip(batch, (r2["seq"], r2["len"], r2["cell"], r2["umi"], r2["flags"])))
    ent2, met2, codes2, s2 = _run_gpu(sb, abc, b2, 40, splits=2)
    ref2 = orc.count_run(abc["oc"], abc["og"], *b2, want_codes=True)
    assert np.array_equal(codes2, ref2["codes"]) and met2 == ref2["metrics"] and ent2 == ref2["entries"] and s2.group_stats()["fallbacks"] == 1
    b3 = (r2["seq"], r2["len"], r2["cell"], r2["umi"], r2["flags"])
    ent3, met3, codes3, s3 = _run_gpu(sb, abc, b3, 40, splits=5)
    ref3 = orc.count_run(abc["oc"], abc["og"], *b3, want_codes=True)
    assert ent3 == ref3["entries"] and met3 == ref3["metrics"]
    st3 = s3.group_stats()
    assert st3["fallbacks"] == 0 and st3["keys"] == int((codes3 != 255).sum()) and st3["buckets"] >= 128


def test_seed_stride_3(sb, orc, abc):
    """The seed-scan stride of map_read is not pinned by the reference's goldens (strides 1..3 reproduce both, SURVEY.md
    finding #4; the pinned debruijn_mapping revision is not vendored): the library honours 1 (default) and 3
    (hlac_set_seed_stride). Reads with substitutions separate the two: with stride 3 only every third position after the
    scan start may seed, so some reads seed later (or not at all) and get a different coverage / class."""
    from tools import synth
    cds, gen = synth.fixture_alleles()
    r = synth.make_reads(80000, [s for _, s in cds], [s for _, s in gen], 60, 31, err=0.03)
    batch = (r["seq"], r["len"], r["cell"], r["umi"], r["flags"])
    out = {}
    try:
        for stride in (1, 3):
            sb.set_seed_stride(stride)
            orc.set_stride(stride)
            assert sb.get_seed_stride() == stride
            ent, met, codes, _ = _run_gpu(sb, abc, batch, 60, splits=2)
            ref = orc.count_run(abc["oc"], abc["og"], *batch, want_codes=True)
            assert np.array_equal(codes, ref["codes"]) and met == ref["metrics"] and ent == ref["entries"]
            out[stride] = (codes, met)
    finally:
        sb.set_seed_stride(1)
        orc.set_stride(1)
    assert not np.array_equal(out[1][0], out[3][0]), "the workload does not separate stride 1 from stride 3"
    with pytest.raises(sb.HlacError):
        sb.set_seed_stride(2)


def test_wire_records_equal_soa(sb, orc, abc, fixture_reads):
    """hlac_count_push_wire (bit-packed records: 28 bytes per 91-bp read, one H2D copy per push) must give exactly what the
    five-array push gives - host and device buffers, several pushes, --primary-alignments, non-whitelisted reads - and the
    oracle's result; other field widths and read lengths too."""
    import torch
    from tools import synth
    cds, gen = synth.fixture_alleles()
    r = synth.make_reads(70000, [s for _, s in cds], [s for _, s in gen], 300, 23)
    b = (r["seq"], r["len"], r["cell"], r["umi"], r["flags"])
    for primary_only in (False, True):
        ref = orc.count_run(abc["oc"], abc["og"], *b, primary_only=primary_only, want_codes=True)
        for cb, ub in ((9, 21), (17, 32)):
            rec = sb.pack_wire(*b, 91, cb, ub)
            assert rec.shape[1] == (7 if (cb, ub) == (9, 21) else 8)
            for device in (False, True):
                s = sb.CountSession(abc["gc"], abc["gg"], 300, primary_only)
                s.record_codes(True)
                codes = []
                for part in np.array_split(np.arange(len(rec)), 3):
                    chunk = rec[part[0]:part[-1] + 1]
                    s.push_wire(torch.from_numpy(chunk.view(np.int32)).cuda() if device else chunk, 91, cb, ub)
                    codes.append(s.last_codes(len(chunk)))
                ent, met = s.finish()
                assert np.array_equal(np.concatenate(codes), ref["codes"]) and met == ref["metrics"] and ent == ref["entries"]
    # a read length whose bases end in the middle of a word next to the fields, and one that fills its words exactly
    rng = np.random.default_rng(3)
    rc = lambda s: s[::-1].translate(str.maketrans("ACGT", "TGCA"))
    for L in (64, 75, 128):
        reads = []
        for i in range(4000):
            src = (cds if rng.random() < 0.7 else gen)[int(rng.integers(0, 6))][1]
            st = int(rng.integers(0, len(src) - L))
            sq = src[st:st + L]
            reads.append((rc(sq) if rng.random() < 0.5 else sq, b"c%d" % int(rng.integers(0, 9)), b"ACGTACGT%s" % (b"ACGT"[i % 4:i % 4 + 1]), i % 11 != 0))
        bc = {b"c%d" % i: i for i in range(8)}    # c8 is not in the list
        batch = H.make_batch(reads, bc, pack=sb.pack_reads)
        ref = orc.count_run(abc["oc"], abc["og"], *batch)
        s = sb.CountSession(abc["gc"], abc["gg"], 8)
        s.push_wire(sb.pack_wire(*batch, L, 4, 19), L, 4, 19)
        assert s.finish() == (ref["entries"], ref["metrics"])
    with pytest.raises(sb.HlacError, match="cell_bits"):   # 300 columns do not fit 8 bits + the not-a-cell value
        s300 = sb.CountSession(abc["gc"], abc["gg"], 300)
        s300.push_wire(np.zeros((1, 7), np.uint32), 91, 8, 21)
