"""Multi-GPU sharding of the counting path: reads are partitioned by barcode so that every
(cell, UMI) molecule is complete on one rank. Two ways to finish: (a) shard-local columns - a rank renumbers its
barcodes c -> c // world, holds only its own column block and needs no data-path collective (the blocks are
interleaved back, column = local * world + rank, when the matrix is written); (b) global columns - per-rank dense
matrices have disjoint columns and one all-reduce(sum) merges them exactly. Used by bench.py and by the gloo
world_size-2 CPU tests."""
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


def to_local(cell, world):
    """shard-local column of every read of a rank's shard (NOT_A_CELL stays)"""
    cell = np.asarray(cell)
    return np.where(cell == NOT_A_CELL, cell, cell // world).astype(np.uint32)


def n_local_cells(n_cells, rank, world):
    return (n_cells - rank + world - 1) // world


def to_global_entries(entries, rank, world):
    """(row, local column, value) of rank's block -> (row, global column, value)"""
    return [(r, c * world + rank, v) for r, c, v in entries]


def merge_blocks(blocks):
    """per-rank entry lists in global columns -> one list in the writer's order (cells ascending, rows ascending)"""
    return sorted((e for b in blocks for e in b), key=lambda e: (e[1], e[0]))
