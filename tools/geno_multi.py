"""Multi-GPU genotyping (BASELINE config 5 shape): run under torchrun, one rank per GPU.
   python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/geno_multi.py [alleles_per_gene] [reads]
Reads (incl. an unmapped-style forward-only tail) are partitioned by barcode with their global ordinals; the sessions
carry the library's own NCCL communicator (hlac_geno_attach_comm), so hlac_geno_finish merges by itself: path union on the
device over all-gathers, every rank resolves the union paths it owns, fingerprints exchanged, UMI histogram all-reduced;
every rank runs EM + caller on identical tables. HLAC_GENO_MULTI_LEGACY=1 runs the round-1 protocol instead (paths
exported to the host, pickled through torch.distributed, merged on the host, every rank resolves everything).
Rank 0 checks the result against a single-GPU run of the whole stream."""
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import schlacount_b200 as sb  # noqa: E402
from tools import synth  # noqa: E402

npg = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2_000_000
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
seed = 20191101 + 5
db = os.path.join(tempfile.gettempdir(), "hladb_multi_%d_%d" % (npg, rank))
names = synth.make_db(db, npg, seed)   # every rank builds the same DB (seeded)
donor = [names[3], names[npg // 2], names[npg + 5], names[npg + 5], names[2 * npg + 1], names[3 * npg - 1]]
r = synth.reads_for_db(n, db, donor, 20000, seed, frac_cds=0.45, frac_gen=0.05)   # ~50 % off-target ("unmapped-style")
n_tail = n // 4                        # last quarter plays the --unmapped pass: forward-only (mapping.rs:380-393)
fa, st = os.path.join(db, "hla_nuc.fasta"), os.path.join(db, "Allele_status.txt")
gix = sb.Index.from_fasta(fa, st, device=local)
ordinal = np.arange(n, dtype=np.uint64)
cell = r["cell"].copy()
cell[cell == synth.NOT_A_CELL] = 0     # genotyping uses every barcode; keep ids injective enough for the test


def run(mask):
    g = sb.GenoSession(gix, lean=True)   # counts_reads stays on the device; the caller runs on the device-resident vectors
    t0 = time.time()
    for lo, hi, fo in ((0, n - n_tail, False), (n - n_tail, n, True)):
        idx = np.nonzero(mask[lo:hi])[0] + lo
        if len(idx):
            g.push(r["seq"][idx], r["len"][idx], cell[idx], r["umi"][idx], forward_only=fo, ordinal=ordinal[idx])
    return g, time.time() - t0


class _W:
    def __init__(self, ptr, nelem):
        self.__cuda_array_interface__ = {"shape": (max(nelem, 1),), "typestr": "<i4", "data": (ptr, False), "version": 2}


legacy = bool(os.environ.get("HLAC_GENO_MULTI_LEGACY"))
comm = sb.Comm.from_torch_distributed(local) if (world > 1 and not legacy) else None
for rep in range(2):   # the second pass is the warm one (kernel attributes, the per-index colour bitsets, allocator)
    g, t_push = run((cell % world) == rank)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.time()
    if comm is not None:
        g.attach_comm(comm)
        tab = g.finish(raw=True)
        n_union = g.stats()["n_paths"]
    else:
        if world > 1:
            parts = [None] * world
            dist.all_gather_object(parts, g.export_paths())
            n_union = g.import_merged_paths(parts)
        else:
            n_union = g.stats()["n_paths"]
        ptr, nc = g.finish_begin()
        if world > 1:
            dist.all_reduce(torch.as_tensor(_W(ptr, nc), device=dev))
            torch.cuda.synchronize()   # the session reads the histogram back on its own stream
        tab = g.finish_end(raw=True)
    torch.cuda.synchronize()
    t_fin = time.time() - t0
    if world > 1:
        tt = torch.tensor([t_push, t_fin], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_push, t_fin = float(tt[0]), float(tt[1])
    if rep == 0:
        cold = dict(push_s=t_push, merge_finish_s=t_fin)
theta, iters = g.squarem(device=local)
call = g.call(theta)
out = dict(world=world, reads=n, alleles=len(names), cold_first_pass=cold, protocol="legacy-host" if (legacy or world == 1) else "in-library", push_s=t_push, merge_finish_s=t_fin,
           union_paths=n_union, resolve=g.resolve_stats(), timing=g.timing(), tables=tab,
           em_iters=iters, called=[gix.tx_names[i] for i in call["called"]])
if rank == 0:
    ref, _ = run(np.ones(n, bool))
    rt = ref.finish(raw=True)
    rtheta, riters = ref.squarem(device=local)
    out["single_gpu"] = dict(tables=rt, em_iters=riters, max_abs_theta_diff=float(np.abs(rtheta - theta).max()),
                             called=[gix.tx_names[i] for i in ref.call(rtheta)["called"]])
    out["matches_single_gpu"] = bool(rt["nreads"] == tab["nreads"] and rt["n_read_classes"] == tab["n_read_classes"]
                                     and rt["n_umi_classes"] == tab["n_umi_classes"] and riters == iters
                                     and np.array_equal(rtheta, theta) and ref.call(rtheta)["called"] == call["called"])
    print(json.dumps(out))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
