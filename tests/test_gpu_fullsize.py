"""Full-size (BASELINE config 2: 10 M reads, 10 k barcodes) checks of the CUDA counting path through size-independent
properties, since the oracle needs ~50 s for this many reads: chunking invariance, read-order invariance, metric
identities, molecule conservation, and exact agreement with the oracle on a random 200 k-read subset of CELLS
(a subset of whole cells is a self-contained sub-problem because count() never mixes cells)."""
import os

import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu
N, CELLS = 10_000_000, 10_000


@pytest.fixture(scope="module")
def data():
    from tools import synth
    cds, gen = synth.fixture_alleles()
    return synth.make_reads(N, [s for _, s in cds], [s for _, s in gen], CELLS, 20191103)


def _run(sb, r, order=None, chunk=None):
    cf, gf = os.path.join(H.FIX, "cds_ABC.fa"), os.path.join(H.FIX, "genomic_ABC.fa")
    s = sb.CountSession(sb.Index.from_fasta(cf), sb.Index.from_fasta(gf), CELLS)
    arrs = [r[k] if order is None else r[k][order] for k in ("seq", "len", "cell", "umi", "flags")]
    n = len(arrs[1])
    chunk = chunk or n
    for a in range(0, n, chunk):
        s.push(*[x[a:a + chunk] for x in arrs])
    (row, col, val), met = s.entries_arrays() if s.reduce() else (None, None)
    return np.stack([col, row, val], 1), met, s


def test_full_size_properties(data, orc):
    import schlacount_b200 as sb
    base, met, s = _run(sb, data)
    # metric identities (mapping.rs:742-793): every read is counted once; filtered reads never reach the aligner
    assert met[0] == N
    passed = N - met[2] - 0   # primary_only is off: non-primary reads are only counted
    assert met[3] + met[4] + met[5] <= passed
    assert int((data["cell"] == H.NOT_A_CELL).sum()) == met[2]
    assert int((data["flags"] & 1 == 0).sum()) == met[1]
    # molecules: at most one per (cell, umi) pair among whitelisted reads
    ok = data["cell"] != H.NOT_A_CELL
    n_pairs = len(np.unique((data["cell"][ok].astype(np.uint64) << np.uint64(32)) | data["umi"][ok].astype(np.uint64)))
    assert 0 < base[:, 2].sum() <= n_pairs
    assert (np.diff(base[:, 0] * 64 + base[:, 1]) > 0).all()   # cells ascending, rows ascending, no duplicates
    # chunking invariance (the push API streams batches) and read-order invariance (count() groups by cell and UMI)
    chunked, met2, _ = _run(sb, data, chunk=1_234_567)
    assert np.array_equal(base, chunked) and met == met2
    perm = np.random.default_rng(1).permutation(N)
    shuffled, met3, _ = _run(sb, data, order=perm)
    assert np.array_equal(base, shuffled) and met == met3
    # exact agreement with the oracle on all reads of a random subset of cells
    cells = np.random.default_rng(2).choice(CELLS, 200, replace=False)
    m = np.isin(data["cell"], cells)
    sub = {k: v[m] for k, v in data.items()}
    cf, gf = os.path.join(H.FIX, "cds_ABC.fa"), os.path.join(H.FIX, "genomic_ABC.fa")
    ref = orc.count_run(orc.Index.from_fasta(cf), orc.Index.from_fasta(gf), sub["seq"], sub["len"], sub["cell"], sub["umi"], sub["flags"])
    mine = base[np.isin(base[:, 0], cells)]
    assert [(int(r), int(c), int(v)) for c, r, v in mine] == ref["entries"]
    assert m.sum() > 150_000
