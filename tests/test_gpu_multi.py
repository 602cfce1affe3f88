"""Multi-GPU counting: reads partitioned by barcode, every rank's session over GLOBAL columns, the dense
matrices summed. Three levels: (1) the partition + merge arithmetic emulated with several sessions on ONE GPU (always runs),
(2) the library's own NCCL communicator (hlac_comm_init_all + hlac_count_attach_comm, merge inside hlac_count_finish) on two
real GPUs, (3) sc_hla_count --devices. Everything is compared with the CPU oracle run over the unpartitioned stream."""
import os
import subprocess
import threading

import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sb():
    import schlacount_b200
    return schlacount_b200


def _n_gpus():
    import torch
    return torch.cuda.device_count()


def _workload(n=300_000, cells=700, seed=5):
    from tools import synth
    cds, gen = synth.fixture_alleles()
    r = synth.make_reads(n, [s for _, s in cds], [s for _, s in gen], cells, seed)
    return r, cells


def _oracle(orc, r):
    cf, gf = os.path.join(H.FIX, "cds_ABC.fa"), os.path.join(H.FIX, "genomic_ABC.fa")
    return orc.count_run(orc.Index.from_fasta(cf), orc.Index.from_fasta(gf), r["seq"], r["len"], r["cell"], r["umi"], r["flags"])


@pytest.mark.parametrize("world", [2, 3])
def test_count_partition_and_merge_emulated_vs_oracle(sb, orc, world):
    """`world` sessions on one GPU, each fed the reads of its barcodes (column % world; reads outside the whitelist spread
    round-robin), dense matrices summed, rank 0 extracts: entries and metrics equal the oracle's on the whole stream."""
    import torch
    from tools import sharding
    r, cells = _workload()
    ref = _oracle(orc, r)
    cf, gf = os.path.join(H.FIX, "cds_ABC.fa"), os.path.join(H.FIX, "genomic_ABC.fa")
    gc, gg = sb.Index.from_fasta(cf), sb.Index.from_fasta(gf)
    sess, dense = [], []
    for rank in range(world):
        sh = sharding.shard(r, rank, world)
        s = sb.CountSession(gc, gg, cells)
        for part in np.array_split(np.arange(len(sh["len"])), 3):
            s.push(*[sh[k][part] for k in ("seq", "len", "cell", "umi", "flags")])
        ptr, n = s.reduce()

        class _W:
            __cuda_array_interface__ = {"shape": (n,), "typestr": "<i4", "data": (ptr, False), "version": 2}
        sess.append(s)
        dense.append(torch.as_tensor(_W(), device=torch.device("cuda", 0)))   # the sessions' device: a view, never a cross-device copy
    total = sum(dense[1:], dense[0].clone())
    # disjoint columns: at most one rank has a non-zero in any cell
    assert int(sum((d != 0).to(torch.int32) for d in dense).max()) <= 1
    dense[0].copy_(total)
    torch.cuda.synchronize()
    ent, met0 = sess[0].entries()
    mets = [met0] + [s.entries()[1] for s in sess[1:]]
    assert ent == ref["entries"]
    assert [sum(m[i] for m in mets) for i in range(6)] == ref["metrics"]


@pytest.mark.parametrize("root", [0, -1])
def test_count_comm_two_gpus_vs_oracle(sb, orc, root):
    """hlac_comm_init_all + hlac_count_attach_comm: the merge runs inside hlac_count_finish over NCCL (root 0: reduce, only
    rank 0 returns entries; -1: all-reduce, both do). One host thread per device, as the CLI's --devices does."""
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    from tools import sharding
    r, cells = _workload(400_000, 1500, 6)
    ref = _oracle(orc, r)
    cf, gf = os.path.join(H.FIX, "cds_ABC.fa"), os.path.join(H.FIX, "genomic_ABC.fa")
    comms = sb.Comm.init_all([0, 1])
    assert [c.info["rank"] for c in comms] == [0, 1] and comms[0].info["nranks"] == 2
    out = [None, None]

    def run(rank):
        gc, gg = sb.Index.from_fasta(cf, device=rank), sb.Index.from_fasta(gf, device=rank)
        s = sb.CountSession(gc, gg, cells)
        s.attach_comm(comms[rank], root)
        sh = sharding.shard(r, rank, 2)
        res = []
        for rep in range(2):   # a second step on the same session: reset, push, finish again
            s.reset()
            for part in np.array_split(np.arange(len(sh["len"])), 2):
                s.push(*[sh[k][part] for k in ("seq", "len", "cell", "umi", "flags")])
            res.append(s.finish())
        out[rank] = res
        s.close()
    th = [threading.Thread(target=run, args=(k,)) for k in range(2)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for c in comms:
        c.close()
    for rep in range(2):
        assert out[0][rep][0] == ref["entries"] and out[0][rep][1] == ref["metrics"]
        assert out[1][rep][1] == ref["metrics"]                                  # metrics are global on every rank
        assert out[1][rep][0] == (ref["entries"] if root < 0 else [])


@pytest.mark.parametrize("lean", [False, True])
def test_geno_comm_two_gpus_vs_oracle(sb, orc, tmp_path, lean):
    """hlac_geno_attach_comm: mapped part + forward-only "unmapped" tail (mapping.rs:380-393) partitioned by barcode over two
    GPUs with global ordinals; the finish merges inside the library (path union on the device over NCCL all-gathers, each rank
    resolves the union paths it owns, fingerprints exchanged, UMI histogram all-reduced). Both ranks must return the ORACLE's
    tables of the whole stream, first-seen class ids included, and identical EM / calls."""
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    from tools import synth
    db = str(tmp_path / "db")
    names = synth.make_db(db, 400, 61, extra_genes=("DRB1", "DQB1"))
    donor = [names[3], names[250], names[405], names[777], names[801], names[1100], names[1300], names[1599], names[1601], names[1999]]
    r = synth.reads_for_db(90_000, db, donor, 800, 62, frac_cds=0.5, frac_gen=0.05)
    r["cell"][r["cell"] == synth.NOT_A_CELL] = 7
    n = len(r["len"])
    tail = n - n // 4
    fa, st = os.path.join(db, "hla_nuc.fasta"), os.path.join(db, "Allele_status.txt")
    oix = orc.Index.from_fasta(fa, st)
    o = orc.Geno(oix)
    ci0, _ = o.push(r["seq"][:tail], r["len"][:tail], r["cell"][:tail], r["umi"][:tail])
    ci1, _ = o.push(r["seq"][tail:], r["len"][tail:], r["cell"][tail:], r["umi"][tail:], forward_only=True)
    o.finish()
    ocid = np.concatenate([ci0, ci1])
    ref = o.call()
    comms = sb.Comm.init_all([0, 1])
    ordinal = np.arange(n, dtype=np.uint64)
    out, err = [None, None], [None, None]

    def run(rank):
        try:
            gix = sb.Index.from_fasta(fa, st, device=rank)
            g = sb.GenoSession(gix, record_reads=True, lean=lean)
            g.attach_comm(comms[rank])
            mine = []
            for lo, hi, fo in ((0, tail, False), (tail, n, True)):
                idx = np.nonzero((r["cell"][lo:hi] % 2) == rank)[0] + lo
                for part in np.array_split(idx, 2):
                    g.push(r["seq"][part], r["len"][part], r["cell"][part], r["umi"][part], forward_only=fo, ordinal=ordinal[part])
                mine.append(idx)
            mine = np.concatenate(mine)
            g.finish(raw=True)
            t = g.tables_csr()
            theta, iters = g.squarem(device=rank)
            out[rank] = dict(t=t, cid=g.read_class_ids(len(mine)), mine=mine, theta=theta, iters=iters, call=g.call(theta, on_device=True))
            g.close()
        except Exception as e:   # surface worker failures in the main thread
            err[rank] = e
    th = [threading.Thread(target=run, args=(k,)) for k in range(2)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for c in comms:
        c.close()
    assert err == [None, None], err
    for rank in range(2):
        t = out[rank]["t"]
        assert t["nreads"] == o.nreads and np.array_equal(t["reads_explained"], o.reads_explained())
        assert H.csr_tables_equal(t["umi"], o.table_raw("umi"))
        if lean:
            assert t["reads"] is None
        else:
            assert H.csr_tables_equal(t["reads"], o.table_raw("reads"))
        assert np.array_equal(out[rank]["cid"], ocid[out[rank]["mine"]])       # GLOBAL first-seen class ids
        assert out[rank]["call"]["called"] == ref["called"]
    assert out[0]["iters"] == out[1]["iters"] and np.array_equal(out[0]["theta"], out[1]["theta"])   # identical tables, deterministic EM
    assert len(ref["called"]) >= 5


def test_cli_devices_two_gpus(tmp_path):
    """sc_hla_count --devices 0,1 on the reference's fixture: count_matrix.mtx byte-identical to the golden
    (src/main.rs:432-466) and to the single-device run."""
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    exe = os.path.join(H.ROOT, "schlacount_b200", "sc_hla_count")
    base = ["-b", os.path.join(H.FIX, "test.bam"), "-g", os.path.join(H.FIX, "genomic_ABC.fa"), "-f", os.path.join(H.FIX, "cds_ABC.fa")]
    for name, bcs, extra in (("one", "barcodes1.tsv", []), ("two", "barcodes1.tsv", ["--devices", "0,1"]), ("seven", "barcodes7.tsv", ["--devices", "0-1"])):
        r = subprocess.run([exe] + base + ["-c", os.path.join(H.FIX, bcs), "-o", str(tmp_path / name)] + extra, capture_output=True, text=True, cwd=tmp_path)
        assert r.returncode == 0, r.stdout + r.stderr
    gold = open(os.path.join(H.FIX, "test_allele_fasta.mtx")).read()
    assert open(tmp_path / "one" / "count_matrix.mtx").read() == gold == open(tmp_path / "two" / "count_matrix.mtx").read()
    assert open(tmp_path / "one" / "summary.tsv").read() == open(tmp_path / "two" / "summary.tsv").read()
    assert "Counting on 2 GPUs" in (r.stdout + r.stderr)


def test_cli_devices_genotyping_two_gpus(tmp_path):
    """sc_hla_count -d <db> --devices 0,1 (test_call, src/main.rs:468-500): genotyping sessions on both GPUs merged inside
    the finish; every output file - the bincode EqClassDb hand-over included - identical to the single-device run."""
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    exe = os.path.join(H.ROOT, "schlacount_b200", "sc_hla_count")
    for name, extra in (("one", []), ("two", ["--devices", "0,1"])):
        args = ["-b", os.path.join(H.FIX, "test.bam"), "-d", os.path.join(H.FIX, "fake_db"), "-c", os.path.join(H.FIX, "barcodes0.tsv"),
                "-o", str(tmp_path / name), "--pl-tmp", str(tmp_path / (name + "_p")), "-i", os.path.join(H.FIX, "fake_db", "hla_nuc.fasta.idx")]
        r = subprocess.run([exe] + args + extra, capture_output=True, text=True, cwd=tmp_path)
        assert r.returncode == 0, r.stdout + r.stderr
    assert "Genotyping on 2 GPUs" in (r.stdout + r.stderr)
    assert H.read_mtx(tmp_path / "two" / "count_matrix.mtx") == H.read_mtx(os.path.join(H.FIX, "test_call.mtx"))
    for f in ("weights.tsv", "pairs.tsv", "cds_pseudoaln.fasta", "gen_pseudoaln.fasta", "pseudoaligner.counts.bin"):
        assert open(tmp_path / "one_p" / f, "rb").read() == open(tmp_path / "two_p" / f, "rb").read(), f
