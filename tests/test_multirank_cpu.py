"""world_size-2 gloo test of the N>1 path's host logic: partition reads by barcode, count each shard
independently, all-reduce(sum) the dense matrices and the metrics — must equal the single-rank result exactly.
The per-shard counting is done by the CPU oracle here (no GPU in this container); on the GPU box bench.py runs the
same partition + all-reduce with the CUDA sessions over NCCL."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers as H


def _worker(rank, world, port, out):
    sys.path.insert(0, H.ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as orc
    from tools import sharding, synth
    cds_s, gen_s = synth.fixture_alleles()
    reads = synth.make_reads(30000, [s for _, s in cds_s], [s for _, s in gen_s], 97, 4242)
    mine = sharding.shard(reads, rank, world)
    cds = orc.Index.from_fasta(os.path.join(H.FIX, "cds_ABC.fa"))
    gen = orc.Index.from_fasta(os.path.join(H.FIX, "genomic_ABC.fa"))
    res = orc.count_run(cds, gen, mine["seq"], mine["len"], mine["cell"], mine["umi"], mine["flags"])
    dense = torch.from_numpy(sharding.dense_from_entries(res["entries"], 97, res["nrows"]))
    met = torch.tensor(res["metrics"], dtype=torch.int64)
    # disjoint columns: what this rank counted lies only in its own cells
    cols = np.nonzero(dense.numpy().sum(axis=1))[0]
    assert (cols % world == rank).all()
    dist.all_reduce(dense)
    dist.all_reduce(met)
    if rank == 0:
        full = orc.count_run(cds, gen, reads["seq"], reads["len"], reads["cell"], reads["umi"], reads["flags"])
        ok = sharding.entries_from_dense(dense.numpy()) == full["entries"] and met.tolist() == full["metrics"]
        out.put(bool(ok) and len(full["entries"]) > 50)
    dist.barrier()
    dist.destroy_process_group()


def test_partition_by_barcode_and_allreduce_gloo():
    from oracle import oracle as orc
    orc.build()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert out.get(timeout=5) is True


def _worker_local(rank, world, port, out):
    """shard-local columns: no collective on the data path; blocks are interleaved back at the end"""
    sys.path.insert(0, H.ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as orc
    from tools import sharding, synth
    cds_s, gen_s = synth.fixture_alleles()
    n_cells = 97
    reads = synth.make_reads(30000, [s for _, s in cds_s], [s for _, s in gen_s], n_cells, 777)
    mine = sharding.shard(reads, rank, world)
    cds = orc.Index.from_fasta(os.path.join(H.FIX, "cds_ABC.fa"))
    gen = orc.Index.from_fasta(os.path.join(H.FIX, "genomic_ABC.fa"))
    local = sharding.to_local(mine["cell"], world)
    ok = local != sharding.NOT_A_CELL
    assert local[ok].max() < sharding.n_local_cells(n_cells, rank, world)
    res = orc.count_run(cds, gen, mine["seq"], mine["len"], local, mine["umi"], mine["flags"])
    blocks = [None] * world
    dist.all_gather_object(blocks, sharding.to_global_entries(res["entries"], rank, world))   # result assembly, off the data path
    met = torch.tensor(res["metrics"], dtype=torch.int64)
    dist.all_reduce(met)
    if rank == 0:
        full = orc.count_run(cds, gen, reads["seq"], reads["len"], reads["cell"], reads["umi"], reads["flags"])
        out.put(sharding.merge_blocks(blocks) == full["entries"] and met.tolist() == full["metrics"] and len(full["entries"]) > 50)
    dist.barrier()
    dist.destroy_process_group()


def test_partition_by_barcode_local_columns_gloo():
    from oracle import oracle as orc
    orc.build()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29950 + os.getpid() % 300
    procs = [ctx.Process(target=_worker_local, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert out.get(timeout=5) is True


def test_owner_of_covers_every_read_once():
    from tools import sharding
    rng = np.random.default_rng(0)
    cell = rng.integers(0, 50, 1000).astype(np.uint32)
    cell[rng.random(1000) < 0.1] = sharding.NOT_A_CELL
    for world in (1, 2, 4, 8):
        own = sharding.owner_of(cell, world)
        assert own.min() >= 0 and own.max() < world
        ok = cell != sharding.NOT_A_CELL
        assert (own[ok] == cell[ok] % world).all()
