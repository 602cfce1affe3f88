"""Multi-GPU sharding of the counting path: reads are partitioned by barcode so that every
(cell, UMI) molecule is complete on one rank; per-rank dense count matrices then have disjoint columns and one
all-reduce(sum) merges them exactly. Used by bench.py and by the gloo world_size-2 CPU test."""
import numpy as np

NOT_A_CELL = 0xFFFFFFFF


def owner_of(cell, world, index=None):
    """Rank owning each read: cell % world; reads outside the whitelist (never scored, only counted in the metrics)
    are spread round-robin by read index so that each is counted exactly once."""
    cell = np.asarray(cell)
    idx = np.arange(len(cell)) if index is None else index
    return np.where(cell == NOT_A_CELL, idx % world, cell % world).astype(np.int64)


def shard(reads, rank, world):
    """reads: dict of arrays (seq, len, cell, umi, flags) -> this rank's sub-arrays (order preserved)."""
    m = owner_of(reads["cell"], world) == rank
    return {k: v[m] for k, v in reads.items()}


def dense_from_entries(entries, n_cells, nrows):
    d = np.zeros((n_cells, nrows), dtype=np.int64)
    for r, c, v in entries:
        d[c, r] += v
    return d


def entries_from_dense(d):
    """cells ascending, rows ascending (mapping.rs:829-832,897-906)"""
    c, r = np.nonzero(d)
    return [(int(rr), int(cc), int(d[cc, rr])) for cc, rr in zip(c, r)]
