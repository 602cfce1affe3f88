"""Pins the CPU oracle on every golden vector the reference's own tests hold for the hot path
(src/main.rs:432-500 -> test/test_allele_fasta.mtx, test/test_call.mtx; test/fake_db/hla_nuc.fasta.idx;
src/hla.rs:272-316 parser KATs; src/em.rs:483-528 datasets)."""
import os

import numpy as np

import helpers as H


def test_golden_allele_fasta(orc, fixture_reads):
    """`-b test.bam -g genomic_ABC.fa -f cds_ABC.fa -c barcodes1.tsv` == test_allele_fasta.mtx (main.rs:432-466)."""
    cds = orc.Index.from_fasta(os.path.join(H.FIX, "cds_ABC.fa"))
    gen = orc.Index.from_fasta(os.path.join(H.FIX, "genomic_ABC.fa"))
    bc = H.load_barcodes(os.path.join(H.FIX, "barcodes1.tsv"))
    res = orc.count_run(cds, gen, *H.make_batch(fixture_reads, bc), want_codes=True)
    nr, nc, ent = H.read_mtx(os.path.join(H.FIX, "test_allele_fasta.mtx"))
    assert (res["nrows"], len(bc)) == (nr, nc)
    assert sorted(res["entries"]) == ent
    assert [e[2] for e in sorted(res["entries"])] == [9, 10, 1, 16, 23, 3, 12, 8, 3]
    # model known-answers recorded in SURVEY.md §8(c) (regression values, not reference outputs)
    assert res["metrics"] == [2864, 203, 2184, 643, 23, 1]
    assert int((res["codes"] != 255).sum()) == 436
    assert cds.rows() == ["A*02:01:154", "A*31:01:02:01", "A", "B*27:05:02:01", "B*51:01:01:01", "B",
                          "C*02:02:02:01", "C*14:02:01:01", "C"]


def _filter_fasta(src, dst, called):
    """em.rs:445-473: keep records whose first description token is a called allele."""
    keep = False
    with open(src) as f, open(dst, "w") as o:
        for line in f:
            if line.startswith(">"):
                keep = line.split(" ")[1] in called
            if keep:
                o.write(line)


def test_golden_call(orc, fixture_reads, tmp_path):
    """`-d fake_db -i fake_db/hla_nuc.fasta.idx -c barcodes0.tsv` == test_call.mtx (main.rs:468-500)."""
    idx = orc.Index.from_idx(os.path.join(H.FIX, "fake_db", "hla_nuc.fasta.idx"))
    g = orc.Geno(idx)
    seq, lens, cell, umi, _ = H.make_batch(fixture_reads, None)
    cid, cov = g.push(seq, lens, cell, umi)
    g.finish()
    call = g.call()
    called = {idx.tx_names[i] for i in call["called"]}
    assert called == {"A*02:01:154", "B*27:05:02:01", "C*02:02:02:01", "C*14:02:01:01"}
    cds_f, gen_f = tmp_path / "cds_pseudoaln.fasta", tmp_path / "gen_pseudoaln.fasta"
    _filter_fasta(os.path.join(H.FIX, "fake_db", "hla_nuc.fasta"), cds_f, called)
    _filter_fasta(os.path.join(H.FIX, "fake_db", "hla_gen.fasta"), gen_f, called)
    cds, gen = orc.Index.from_fasta(cds_f), orc.Index.from_fasta(gen_f)
    bc = H.load_barcodes(os.path.join(H.FIX, "barcodes0.tsv"))
    assert len(bc) == 2 and bc[b""] == 1  # the empty second line is a barcode (column 1)
    res = orc.count_run(cds, gen, *H.make_batch(fixture_reads, bc))
    nr, nc, ent = H.read_mtx(os.path.join(H.FIX, "test_call.mtx"))
    assert (res["nrows"], len(bc)) == (nr, nc) == (5, 2)
    assert sorted(res["entries"]) == ent == [(1, 0, 1)]
    # model known-answers
    assert int((cov > 0).sum()) == 2728 and int((cid != 0xFFFFFFFF).sum()) == 2704
    assert g.n_read_classes == 98 and len(g.table("umi")) == 19 and sum(g.table("umi").values()) == 358
    assert g.reads_explained().tolist() == [784, 684, 1094, 906, 1011, 1060]
    assert call["iters"] == 6
    np.testing.assert_allclose(call["theta"], [0.208734716595, 0.0350130341055, 0.241868895457, 0.238787395631,
                                               0.00831323565494, 0.267282722557], rtol=1e-9)


def test_build_index_reproduces_idx_fixture(orc):
    """build_index over read_hla_cds(filter) of fake_db == the committed hla_nuc.fasta.idx (a genuine
    debruijn_mapping output): same node sequences, ext masks, colour lists, tx order."""
    idx = orc.Index.from_idx(os.path.join(H.FIX, "fake_db", "hla_nuc.fasta.idx"))
    built = orc.Index.from_fasta(os.path.join(H.FIX, "fake_db", "hla_nuc.fasta"),
                                 os.path.join(H.FIX, "fake_db", "Allele_status.txt"))
    assert idx.n_nodes == built.n_nodes == 204 and idx.n_bases == built.n_bases == 8729
    assert idx.n_colours == built.n_colours == 39
    assert idx.tx_names == built.tx_names == ["A*02:01:154", "A*31:01:02:01", "B*27:05:02:01", "B*51:01:01:01",
                                              "C*02:02:02:01", "C*14:02:01:01"]
    assert sorted(idx.nodes()) == sorted(built.nodes())
    assert sum(len(s) - 23 for s, _, _ in idx.nodes()) == 4037
    gen = orc.Index.from_fasta(os.path.join(H.FIX, "genomic_ABC.fa"))
    assert gen.n_nodes == 578 and gen.n_bases == 27531


def test_allele_parser_kats(orc):
    """src/hla.rs:272-316"""
    assert orc.parse_allele("A*01:01:01:01") == ("A", [1, 1, 1, 1])
    assert orc.parse_allele("A*01:01:38L") == ("A", [1, 1, 38, None])
    assert orc.parse_allele("MICB*012") == ("MICB", [12, None, None, None])
    assert orc.parse_allele("MICB*012,5") is None


def test_squarem_em_rs_datasets(orc):
    """src/em.rs:483-528 datasets (the reference tests only print; values recorded in SURVEY.md §8c)."""
    th, it = orc.squarem(4, {(0,): 1, (0, 1): 19, (2,): 10, (3,): 10})
    assert it == 2
    np.testing.assert_allclose(th, [0.5, 0, 0.25, 0.25], atol=1e-12)
    th, it = orc.squarem(6, {(0,): 1, (0, 1): 19, (2,): 10, (3,): 10, (4, 5): 20})
    np.testing.assert_allclose(th, [1 / 3, 0, 1 / 6, 1 / 6, 1 / 6, 1 / 6], atol=1e-12)


def test_fixture_provenance():
    """The data fixtures are byte-identical copies of the reference's (checked only where the reference is mounted)."""
    ref = "/root/reference/test"
    if not os.path.isdir(ref):
        return
    for rel in ["test.bam", "cds_ABC.fa", "genomic_ABC.fa", "barcodes0.tsv", "barcodes1.tsv", "test_allele_fasta.mtx",
                "test_call.mtx", "fake_db/hla_nuc.fasta.idx", "fake_db/hla_nuc.fasta", "fake_db/hla_gen.fasta",
                "fake_db/Allele_status.txt"]:
        assert open(os.path.join(ref, rel), "rb").read() == open(os.path.join(H.FIX, rel), "rb").read(), rel
