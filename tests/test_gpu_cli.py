"""The reference's two end-to-end regression tests (src/main.rs:432-500) through the drop-in CLI `sc_hla_count`
(C++ host driver + CUDA kernels): same flags, count_matrix.mtx compared with the reference's golden matrices."""
import os
import subprocess

import pytest

import helpers as H

pytestmark = pytest.mark.gpu
BIN = os.path.join(H.ROOT, "schlacount_b200", "sc_hla_count")


def _run(args, cwd, env=None):
    return subprocess.run([BIN] + args, cwd=cwd, capture_output=True, text=True, env=dict(os.environ, **(env or {})))


def test_allele_fasta_cli(tmp_path):
    r = _run(["-b", os.path.join(H.FIX, "test.bam"), "-g", os.path.join(H.FIX, "genomic_ABC.fa"), "-f",
              os.path.join(H.FIX, "cds_ABC.fa"), "-c", os.path.join(H.FIX, "barcodes1.tsv"), "-o", str(tmp_path / "r"),
              "--pl-tmp", str(tmp_path / "p")], tmp_path)
    assert r.returncode == 0, r.stdout + r.stderr
    assert H.read_mtx(tmp_path / "r" / "count_matrix.mtx") == H.read_mtx(os.path.join(H.FIX, "test_allele_fasta.mtx"))
    # byte-level layout of the sprs writer (main.rs:282-283)
    assert open(tmp_path / "r" / "count_matrix.mtx").read() == open(os.path.join(H.FIX, "test_allele_fasta.mtx")).read()
    assert open(tmp_path / "r" / "labels.tsv").read().split() == ["A*02:01:154", "A*31:01:02:01", "A", "B*27:05:02:01",
                                                                  "B*51:01:01:01", "B", "C*02:02:02:01", "C*14:02:01:01", "C"]
    summ = [l.split("\t") for l in open(tmp_path / "r" / "summary.tsv").read().splitlines()]
    assert [int(x[1]) for x in summ] == [9, 10, 1, 16, 23, 3, 12, 8, 3]
    # the Metrics block (main.rs:253-276); reads outside the whitelist are filtered and accounted for on the host
    log = r.stdout + r.stderr
    for text, want in (("evaluated (with 'CB' and 'UB' tags): ", 2864), ("that were not primary: ", 203),
                       ("not being associated with a cell barcode in the list provided: ", 2184),
                       ("no alignment score above threshold: ", 1), ("alignments to CDS sequence: ", 643),
                       ("alignments to genomic sequence: ", 23)):
        assert text + str(want) in log, (text, log)
    rp = _run(["-b", os.path.join(H.FIX, "test.bam"), "-g", os.path.join(H.FIX, "genomic_ABC.fa"), "-f",
               os.path.join(H.FIX, "cds_ABC.fa"), "-c", os.path.join(H.FIX, "barcodes7.tsv"), "-o", str(tmp_path / "rp"),
               "--primary-alignments"], tmp_path)
    assert rp.returncode == 0, rp.stdout + rp.stderr
    from oracle import oracle as orc
    import schlacount_b200 as sb
    import bamlite
    orc.build()
    reads = list(bamlite.fetch(os.path.join(H.FIX, "test.bam"), *H.DEFAULT_REGION))
    bc = H.load_barcodes(os.path.join(H.FIX, "barcodes7.tsv"))
    ref = orc.count_run(orc.Index.from_fasta(os.path.join(H.FIX, "cds_ABC.fa")), orc.Index.from_fasta(os.path.join(H.FIX, "genomic_ABC.fa")),
                        *H.make_batch(reads, bc, pack=sb.pack_reads), primary_only=True)
    m = ref["metrics"]
    logp = rp.stdout + rp.stderr
    assert "evaluated (with 'CB' and 'UB' tags): %d" % m[0] in logp and "skipped due to not being primary: %d" % m[1] in logp
    assert "in the list provided: %d" % m[2] in logp and "score above threshold: %d" % m[5] in logp
    assert sorted(H.read_mtx(tmp_path / "rp" / "count_matrix.mtx")[2]) == sorted(ref["entries"])
    # out-dir must not pre-exist (main.rs:341-346)
    r2 = _run(["-b", os.path.join(H.FIX, "test.bam"), "-g", os.path.join(H.FIX, "genomic_ABC.fa"), "-f",
               os.path.join(H.FIX, "cds_ABC.fa"), "-c", os.path.join(H.FIX, "barcodes1.tsv"), "-o", str(tmp_path / "r")], tmp_path)
    assert r2.returncode == 1 and "already exists" in r2.stdout


@pytest.mark.parametrize("with_index,counts_format", [(True, None), (False, None), (True, "tables")])
def test_call_cli(tmp_path, with_index, counts_format):
    args = ["-b", os.path.join(H.FIX, "test.bam"), "-d", os.path.join(H.FIX, "fake_db"), "-c",
            os.path.join(H.FIX, "barcodes0.tsv"), "-o", str(tmp_path / "r"), "--pl-tmp", str(tmp_path / "p")]
    if with_index:
        args += ["-i", os.path.join(H.FIX, "fake_db", "hla_nuc.fasta.idx")]
    r = _run(args, tmp_path, {"HLAC_COUNTS_FORMAT": counts_format} if counts_format else None)
    assert r.returncode == 0, r.stdout + r.stderr
    if counts_format is None:   # by default the hand-over file is the reference's EqClassDb (mapping.rs:403 -> em.rs:347)
        bcs, umis, classes, triples = H.read_eqclassdb_bincode(tmp_path / "p" / "pseudoaligner.counts.bin")
        assert len(triples) == 2704 and len(classes) == 98 and sorted(classes.values()) == list(range(98))
        assert all(k.endswith(b"-1") for k in bcs) and all(len(k) == 10 for k in umis)
    else:                       # HLAC_COUNTS_FORMAT=tables: this library's compact table format
        assert open(tmp_path / "p" / "pseudoaligner.counts.bin", "rb").read(8) == b"HLACCNT1"
    assert H.read_mtx(tmp_path / "r" / "count_matrix.mtx") == H.read_mtx(os.path.join(H.FIX, "test_call.mtx")) == (5, 2, [(1, 0, 1)])
    w = [l.split("\t") for l in open(tmp_path / "p" / "weights.tsv").read().splitlines()]
    assert w[0] == ["allele_name", "em_weight", "reads_explained"]
    assert [x[0] for x in w[1:]] == ["A*02:01:154", "A*31:01:02:01", "B*27:05:02:01", "B*51:01:01:01", "C*14:02:01:01", "C*02:02:02:01"]
    assert abs(float(w[1][1]) - 0.208734716595) < 1e-9 and w[1][2] == "0.28994082840236685"
    pr = [l.split("\t") for l in open(tmp_path / "p" / "pairs.tsv").read().splitlines()]
    assert pr[1:] == [["A*02:01:154", "A*31:01:02:01", "0.34726331360946744"], ["B*27:05:02:01", "B*51:01:01:01", "0.5410502958579881"],
                      ["C*14:02:01:01", "C*02:02:02:01", "0.4474852071005917"]]
    called = [l.split(" ")[1] for l in open(tmp_path / "p" / "cds_pseudoaln.fasta") if l.startswith(">")]
    assert called == ["A*02:01:154", "B*27:05:02:01", "C*02:02:02:01", "C*14:02:01:01"]
    assert os.path.exists(tmp_path / "p" / "pseudoaligner.counts.bin") and os.path.exists(tmp_path / "p" / "gen_pseudoaln.fasta")
    if not with_index:   # make_hla_index wrote the index next to the counts (main.rs:200-201): the reference's own file, bit for bit
        assert open(tmp_path / "p" / "hla_nuc.fasta.idx", "rb").read() == open(os.path.join(H.FIX, "fake_db", "hla_nuc.fasta.idx"), "rb").read()
