#!/usr/bin/env python
"""bench.py — reads pseudoaligned + counted per second (BASELINE.json metric) on N B200s.

Workload (N=1): BASELINE config 2 — synthetic 5' GEX, 10 M reads x 91 bp, 10 k barcodes, HLA-A/B/C known-genotype
counting against the six fixture alleles (CDS + genomic). One step = one full pass of the hot path over the whole
read set, ending with ONE assembled result on the host: k_count_map_q (4-way lazy map_read + class -> gene/slot, keys
appended to per-molecule hash buckets) -> k_group_consensus (shared-memory grouping + count()) -> dense count matrix
(-> summed over the ranks through the library's own NCCL communicator when N > 1: hlac_count_attach_comm) -> ordered COO
triplets + metrics written into host-mapped memory; the host synchronises once per step.
  value : reads/s with the batch already resident in HBM (timed with CUDA events on the session stream).
  e2e   : the same through hlac_count_push_wire with pinned HOST buffers (bit-packed 28-byte records): H2D of every
          input byte and the result on the host inside the timed region.
N > 1: reads are partitioned by barcode, every rank maps its own 10 M-read shard (weak scaling) against
GLOBAL column indices; hlac_count_finish merges inside the timed region (--merge reduce: dense matrices summed onto rank
0, which extracts the whole matrix; allreduce: every rank ends with the whole matrix; none: shard-local column blocks, no
collective - the round-1 number). After the timed region the assembled matrix is checked against the CPU oracle on a
sample of whole cells of every rank.
A second record ("geno") times BASELINE config 3 - genotyping against a ~30 k-allele synthetic DB - on rank 0.
--impl reference: the CPU oracle (a port of the reference; the Rust reference cannot be built in this image) on the
host cores, on a bounded sample of the same workload.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "reads pseudoaligned+counted/sec"
UNIT = "reads/s"
N_READS = 10_000_000
N_CELLS = 10_000
SEED = 20191101 + 2
# algorithmic HBM bytes per read: map stage 32 B in (24 B 2-bit sequence + 4 B cell + 4 B UMI) + 4 B out
MAP_BYTES_PER_READ = 36
SORT_BYTES_PER_KEY = 36   # 12-B record read, written sorted, read by the segmented reduce
STEP_BYTES_PER_READ = 72   # map 36 B + grouping 36 B per read
# measured DRAM traffic of the dominant kernel: written by tools/ncu_digest.py from the latest `ncu --set full` capture
NCU_TRAFFIC_JSON = os.path.join(ROOT, "profiles", "r02_count_map_traffic.json")


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index, period=0.004):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._halt = threading.Event()

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {getattr(nv, k): k.replace("nvmlClocksEventReason", "").replace("nvmlClocksThrottleReason", "")
                     for k in dir(nv) if k.startswith("nvmlClocksEventReason") or k.startswith("nvmlClocksThrottleReason")}
            while not self._halt.is_set():
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, nm in names.items():
                    if isinstance(bit, int) and bit and (r & bit) and nm not in ("GpuIdle", "None", "All"):
                        self.reasons.add(nm)
                time.sleep(self.period)
        except Exception as e:  # NVML unavailable: report that rather than inventing numbers
            self.reasons.add("nvml_error:%s" % type(e).__name__)

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


def allele_set(genes, workdir=None):
    """-> (cds fasta path, genomic fasta path, cds seqs, genomic seqs). genes=3: the six fixture alleles (config 2);
    genes=8: two synthetic alleles for each of the 8 admissible genes (config 4 shape, src/hla.rs:108-117)."""
    from tools import synth
    if genes == 3:
        cds, gen = synth.fixture_alleles()
        return (os.path.join(synth.FIX, "cds_ABC.fa"), os.path.join(synth.FIX, "genomic_ABC.fa"),
                [s for _, s in cds], [s for _, s in gen])
    import tempfile
    d = workdir or os.path.join(tempfile.gettempdir(), "hlac_bench_genes%d_%d" % (genes, os.getpid()))
    synth.make_db(d, 2, SEED, extra_genes=("DRB1", "DPA1", "DPB1", "DQA1", "DQB1")[:genes - 3])
    cds, gen = synth.read_fasta(os.path.join(d, "hla_nuc.fasta")), synth.read_fasta(os.path.join(d, "hla_gen.fasta"))
    return os.path.join(d, "hla_nuc.fasta"), os.path.join(d, "hla_gen.fasta"), [s for _, s in cds], [s for _, s in gen]


def bind_near_gpu(index):
    """Run this process on the CPUs NVML reports as local to GPU `index`, before any pinned buffer is allocated, so that
    the host side of every H2D copy sits on the GPU's NUMA node (8 ranks copying at once otherwise share one socket's
    memory and the inter-socket link). Best effort: returns the CPU count it bound to, or 0."""
    if os.environ.get("HLAC_BENCH_NO_BIND"):
        return 0
    try:
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(index)
        words = nv.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, m in enumerate(words) for b in range(64) if (m >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return 0


def gen_shard(n_reads, n_cells_total, rank, world, seed, alleles=None, local_ids=False):
    """Rank's shard: cells of this rank are {c : c % world == rank} (partition by barcode). local_ids: leave the cell
    indices shard-local (c // world): the rank's session then only holds its own columns."""
    from tools import synth
    cds_s, gen_s = (alleles[2], alleles[3]) if alleles else [[s for _, s in x] for x in synth.fixture_alleles()]
    local_cells = (n_cells_total - rank + world - 1) // world
    r = synth.make_reads(n_reads, cds_s, gen_s, local_cells, seed + 1000 * rank)
    if not local_ids:
        ok = r["cell"] != synth.NOT_A_CELL
        r["cell"][ok] = r["cell"][ok] * world + rank
    return r


def oracle_worker(args):
    cf, gf, arrs = args[:3]
    from oracle import oracle as orc
    cds, gen = orc.Index.from_fasta(cf), orc.Index.from_fasta(gf)
    t = time.perf_counter()
    res = orc.count_run(cds, gen, *arrs)
    return time.perf_counter() - t, res["entries"], res["metrics"]


def run_oracle(r, n, procs, fastas=None):
    """The reference's CPU path (oracle port) on the first n reads, partitioned by barcode over `procs` processes
    (the reference itself is single-threaded; running one instance per barcode subset is how it would be spread over
    cores). Returns (seconds of the slowest worker, entries)."""
    from tools import synth
    cf, gf = fastas or (os.path.join(synth.FIX, "cds_ABC.fa"), os.path.join(synth.FIX, "genomic_ABC.fa"))
    sl = {k: v[:n] for k, v in r.items()}
    if procs <= 1:
        dt, ent, met = oracle_worker((cf, gf, (sl["seq"], sl["len"], sl["cell"], sl["umi"], sl["flags"])))
        return dt, ent
    import multiprocessing as mp
    part = np.where(sl["cell"] == synth.NOT_A_CELL, np.arange(n) % procs, sl["cell"] % procs)
    jobs = []
    for p in range(procs):
        m = part == p
        jobs.append((cf, gf, (sl["seq"][m], sl["len"][m], sl["cell"][m], sl["umi"][m], sl["flags"][m])))
    with mp.get_context("fork").Pool(procs) as pool:
        outs = pool.map(oracle_worker, jobs)
    ent = sorted(e for o in outs for e in o[1])
    # the slowest worker's compute time: shipping the arrays to the workers is not charged to the reference
    return max(o[0] for o in outs), ent


def geno_record(args, device, hbm_peak):
    """BASELINE config 3 on one GPU: genotyping against a synthetic IMGT-like DB (8 genes x --geno-alleles-per-gene alleles
    derived from the fixture alleles, class I + class II), --geno-reads reads of a 12-allele donor: pseudoalignment + path
    interning, eq_class_counts (path resolution, UMI grouping), SQUAREM, caller. Reads/s are whole-pipeline numbers for the
    genotyping half of the reference (mapping_wrapper + em_wrapper, src/mapping.rs:310-405, src/em.rs:339-476)."""
    import tempfile

    import torch

    import schlacount_b200 as sb
    from oracle import oracle as orc
    from tools import synth
    npg, n = args.geno_alleles_per_gene, args.geno_reads
    seed = 20191101 + 3
    db = os.path.join(tempfile.gettempdir(), "hlac_bench_db_%d_%d" % (npg, os.getpid()))
    t0 = time.perf_counter()
    names = synth.make_db(db, npg, seed, extra_genes=("DRB1", "DPA1", "DPB1", "DQA1", "DQB1"))
    donor = [names[g * npg + o] for g in range(8) for o in ((3 + 17 * g) % npg, (npg // 2 + 29 * g) % npg)][:12] + \
            [names[6 * npg + 5], names[7 * npg + 11]]
    r = synth.reads_for_db(n, db, donor, 10000, seed)
    t_gen = time.perf_counter() - t0
    fa, st = os.path.join(db, "hla_nuc.fasta"), os.path.join(db, "Allele_status.txt")
    t0 = time.perf_counter()
    gix = sb.Index.from_fasta(fa, st, device=device)
    t_build = time.perf_counter() - t0
    dev = torch.device("cuda", device)
    u2i = {np.dtype("uint64"): np.int64, np.dtype("uint16"): np.int16, np.dtype("uint32"): np.int32, np.dtype("uint8"): np.uint8}
    d = {k: torch.from_numpy(v.view(u2i[v.dtype])).to(dev) for k, v in r.items()}
    hp = {k: torch.from_numpy(v.view(u2i[v.dtype])).pin_memory() for k, v in r.items() if k != "flags"}
    hn = {k: v.numpy().view(r[k].dtype) for k, v in hp.items()}
    best = {}
    for rep in range(4):   # rep 0 warms up (kernel attributes, the per-index colour bitsets, allocations); best of the other three per leg
                           # (wall-clock legs with a dozen host round trips each: the shared boxes show stray 0.2-1 s host stalls)
        rec = {}
        for leg in ("device", "host"):
            g = sb.GenoSession(gix, lean=True)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            if leg == "device":
                g.push(d["seq"], d["len"], d["cell"], d["umi"])
            else:
                for a in range(0, n, 1_000_000):
                    b = min(n, a + 1_000_000)
                    g.push(hn["seq"][a:b], hn["len"][a:b], hn["cell"][a:b], hn["umi"][a:b])
            torch.cuda.synchronize()
            t_push = time.perf_counter() - t0
            t0 = time.perf_counter()
            tab = g.finish(raw=True)
            t_fin = time.perf_counter() - t0
            t0 = time.perf_counter()
            theta, iters = g.squarem(device=device)
            t_em = time.perf_counter() - t0
            t0 = time.perf_counter()
            call = g.call(theta)
            t_call = time.perf_counter() - t0
            tm, stt = g.timing(), g.stats()
            rec[leg] = dict(push_s=t_push, finish_s=t_fin, em_s=t_em, call_s=t_call, total_s=t_push + t_fin + t_em + t_call,
                            timing_ms=tm, stats=stt, tables=tab, em_iters=iters, called=[gix.tx_names[i] for i in call["called"]],
                            resolve=g.resolve_stats())
            g.close()
        for leg in rec:
            if rep and (leg not in best or rec[leg]["total_s"] < best[leg]["total_s"]):
                best[leg] = rec[leg]
    dv, ho = best["device"], best["host"]
    rs = dv["resolve"]
    out = {"workload": "config3: genotyping, %d alleles (8 genes, class I + II) synthetic IMGT-like DB, %d reads x 91 bp, 12-allele donor, "
                       "lean tables (counts_reads stays on the device); wall clock, best of 3 after one warm-up pass, per leg" % (len(names), n),
           "metric": "reads pseudoaligned+genotyped/sec", "unit": UNIT,
           "value": n / dv["total_s"], "e2e": {"value": n / ho["total_s"], "unit": UNIT,
                                               "h2d_bytes_per_step": int(sum(v.nbytes for v in hn.values())), "ms_per_step": ho["total_s"] * 1e3},
           "stages_ms": {"map": dv["timing_ms"]["map_ms"], "intern": dv["timing_ms"]["intern_ms"], "push_wall": dv["push_s"] * 1e3,
                         "finish_wall": dv["finish_s"] * 1e3, "finish_kernels": dv["timing_ms"]["finish_ms"],
                         "squarem_wall": dv["em_s"] * 1e3, "call_wall": dv["call_s"] * 1e3},
           "map_reads_per_s": n / (dv["timing_ms"]["map_ms"] / 1e3), "em_iters": dv["em_iters"], "called": dv["called"], "donor": sorted(set(donor)),
           "index": dict(gix.info, build_s=t_build), "tables": dv["tables"], "stats": dv["stats"], "data_gen_s": t_gen,
           "gpu_launches": int(dv["timing_ms"]["launches"])}
    if rs and rs.get("ms", 0) > 0:
        gbs = rs["bytes"] / (rs["ms"] / 1e3) / 1e9
        out["roofline"] = {"kernel": "k_resolve_paths_*", "bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                           "launch_ms": rs["ms"], "traffic": None,
                           "algorithmic_bytes": rs["bytes"], "merges": rs["merges"], "element_steps": rs["element_steps"],
                           "element_steps_per_s": rs["element_steps"] / (rs["ms"] / 1e3),
                           "note": "algorithmic bytes = per merge one colour membership bitset + per path its colour ids in and its class vector out; "
                                   "the vectors live in shared memory, so the kernel is bound by shared-memory scans and barriers, not HBM"}
    # CPU baseline + parity: the oracle on the first reads of the same stream
    ns = 20000
    orc.build()
    oix = orc.Index.from_fasta(fa, st)
    o = orc.Geno(oix)
    t0 = time.perf_counter()
    ci, cv = o.push(r["seq"][:ns], r["len"][:ns], r["cell"][:ns], r["umi"][:ns])
    o.finish()
    ocall = o.call()
    dt = time.perf_counter() - t0
    g2 = sb.GenoSession(gix, record_reads=True, lean=True)
    g2.push(r["seq"][:ns], r["len"][:ns], r["cell"][:ns], r["umi"][:ns])
    t2 = g2.finish()
    th2, it2 = g2.squarem(device=device)
    ok = bool(np.array_equal(g2.last_cov(), cv) and np.array_equal(g2.read_class_ids(ns), ci) and t2["counts_umi"] == o.table("umi")
              and t2["reads_explained"] == o.reads_explained().tolist() and t2["nreads"] == o.nreads
              and g2.call(th2)["called"] == ocall["called"])
    assert ok, "genotyping sample differs from the CPU oracle"
    out["cpu_baseline"] = {"value": ns / dt, "unit": UNIT, "cores": 1, "kind": "port",
                           "sample": "first %d reads of the stream through map + eq_class_counts + SQUAREM (%d iterations) + caller, single thread; "
                                     "per-read coverage and class ids, counts_umi, reads_explained and the calls equal the GPU's" % (ns, ocall["iters"])}
    return out


def cell_sample_mask(r, k, n_cells):
    """reads of the cells c with c % stride == 0 ... — a self-contained sub-problem (count() never mixes cells), plus the same
    share of the reads outside the whitelist, so that the sample is a scaled-down copy of the workload"""
    idx = np.arange(len(r["cell"]))
    return np.where(r["cell"] == 0xFFFFFFFF, idx % n_cells < k, r["cell"] < k)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--reads", type=int, default=N_READS, help="reads per GPU per step (default: config 2's 10 M)")
    ap.add_argument("--cells", type=int, default=N_CELLS, help="barcodes per GPU")
    ap.add_argument("--cpu-sample", type=int, default=1_500_000, help="reads of the workload timed on the CPU oracle")
    ap.add_argument("--ref-procs", type=int, default=0, help="--impl reference: worker processes (0 = all cores)")
    ap.add_argument("--chunk", type=int, default=-1,
                    help="e2e push size in reads: -1 (default) = a tapering schedule (40 %, 30 %, 15 %, 7.5 %, 5 %, 2.5 % of the batch: few "
                         "copies, and a short last one so that little map work is left when the last byte has landed), 0 = whole batch, "
                         "N = equal chunks of N reads")
    ap.add_argument("--genes", type=int, default=3, choices=[3, 4, 5, 6, 7, 8],
                    help="3 = the six fixture alleles (config 2); 8 = 2 synthetic alleles x 8 genes (config 4 shape)")
    ap.add_argument("--e2e-format", default="wire", choices=["wire", "packed", "soa"],
                    help="input format of both legs: wire = bit-packed 28-byte records (hlac_count_push_wire), packed = 32-byte "
                         "records (hlac_count_push_packed), soa = the five SoA arrays (hlac_count_push, 35 B/read)")
    ap.add_argument("--merge", default="reduce", choices=["reduce", "allreduce", "none"],
                    help="N > 1: 'reduce' = global columns, the ranks' dense matrices summed onto rank 0 inside hlac_count_finish "
                         "(library-owned NCCL communicator), rank 0 extracts the whole COO; 'allreduce' = every rank ends with the "
                         "whole matrix; 'none' = shard-local column blocks, no data-path collective")
    ap.add_argument("--geno", default="auto", choices=["auto", "on", "off"],
                    help="config-3 genotyping record (rank 0): auto = on for the default single-GPU workload")
    ap.add_argument("--geno-alleles-per-gene", type=int, default=3750, help="x 8 genes = 30 000 alleles (config 3)")
    ap.add_argument("--geno-reads", type=int, default=2_000_000)
    ap.add_argument("--parity-cells", type=int, default=6, help="N > 1: cells per rank whose assembled columns are checked against the oracle")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    merge = args.merge if world > 1 else "single"
    config = {"workload": "%s: synthetic 5' GEX, %d reads x 91 bp per GPU, %d barcodes per GPU, %d CDS + %d genomic "
                          "alleles (%d genes x 2), known-genotype counting" % ("config2" if args.genes == 3 else "config4-shape",
                                                                               args.reads, args.cells, 2 * args.genes, 2 * args.genes, args.genes),
              "reads_per_gpu": args.reads, "barcodes_per_gpu": args.cells,
              "parallelism": "reads sharded by barcode x%d%s" % (world, {
                  "single": "", "none": ", per-rank column blocks, no data-path collective",
                  "reduce": ", one matrix assembled on rank 0: ncclReduce of the dense matrices inside hlac_count_finish",
                  "allreduce": ", every rank assembles the whole matrix: ncclAllReduce inside hlac_count_finish"}[merge]),
              "l2": "inputs (%.0f MB/step) exceed the 126 MB L2; no explicit flush" % (args.reads * 28 / 1e6),
              "seed": SEED}

    if args.impl == "reference":
        if rank != 0:
            return 0
        procs = args.ref_procs or (os.cpu_count() or 1)
        n = min(args.reads, args.cpu_sample * max(1, procs // 2))
        al = allele_set(args.genes)
        r = gen_shard(n, args.cells, 0, 1, SEED, al)
        from oracle import oracle as orc
        orc.build()
        times = []
        for i in range(args.warmup + args.steps):
            dt, _ = run_oracle(r, n, procs, al[:2])
            if i >= args.warmup:
                times.append(dt)
        per = sum(times) / len(times)
        v = n / per
        line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": per * 1e3, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": v, "unit": UNIT, "cores": procs, "kind": "port",
                                 "sample": "first %d reads of the step's %d, partitioned by barcode over %d processes "
                                           "(the reference itself is single-threaded)" % (n, args.reads, procs)},
                "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist
    import schlacount_b200 as sb
    from tools import synth

    bound_cpus = bind_near_gpu(local_rank)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    comm = None
    if world > 1:
        # NCCL prints its version banner on stdout when a communicator is created; keep stdout clean for the single JSON
        # line by pointing fd 1 at stderr while that happens.
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.all_reduce(torch.zeros(1, device=dev))
            torch.cuda.synchronize()
            if merge != "none":
                comm = sb.Comm.from_torch_distributed(local_rank)   # the library's own communicator; torch only carries the id
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    local = merge == "none"
    n_cells_total = args.cells if local else args.cells * world   # columns one session holds
    al = allele_set(args.genes)
    r = gen_shard(args.reads, args.cells * world, rank, world, SEED, al, local_ids=local)
    n = args.reads
    cf, gf = al[0], al[1]
    gc, gg = sb.Index.from_fasta(cf, device=local_rank), sb.Index.from_fasta(gf, device=local_rank)
    sess = sb.CountSession(gc, gg, n_cells_total)
    stream = torch.cuda.current_stream()
    sess.set_stream(stream.cuda_stream)
    sess.reserve(n)
    if comm is not None:
        sess.attach_comm(comm, 0 if merge == "reduce" else -1)
    nrows = sess.nrows
    has_result = rank == 0 or merge in ("none", "allreduce")

    def finish_step():   # ONE call: grouping + consensus (+ merge over the ranks) + COO + metrics, one host synchronisation
        return sess.finish_arrays(copy=False)

    # the same records for both legs: resident in HBM for `value`, in pinned host memory for `e2e`
    u2i = {np.dtype("uint64"): np.int64, np.dtype("uint16"): np.int16, np.dtype("uint32"): np.int32, np.dtype("uint8"): np.uint8}
    if args.chunk < 0:
        cuts = [0] + [int(n * f) for f in (0.40, 0.70, 0.85, 0.925, 0.975)] + [n]
    else:
        step = args.chunk or n
        cuts = list(range(0, n, step)) + [n]
    cuts = sorted(set(cuts))
    fmt = args.e2e_format
    if fmt == "soa":
        d = {k: torch.from_numpy(v.view(u2i[v.dtype])).to(dev) for k, v in r.items()}
        h = {k: torch.from_numpy(v.view(u2i[v.dtype])).pin_memory() for k, v in r.items()}
        hn = {k: v.numpy().view(r[k].dtype) for k, v in h.items()}
        in_bytes = sum(int(v.nbytes) for v in hn.values())
        push_dev = lambda: sess.push(d["seq"], d["len"], d["cell"], d["umi"], d["flags"])
        push_host = lambda a, b: sess.push(hn["seq"][a:b], hn["len"][a:b], hn["cell"][a:b], hn["umi"][a:b], hn["flags"][a:b])
        fmt_name = "five SoA arrays, %d B/read (hlac_count_push)"
    elif fmt == "packed":
        rec_np = sb.pack_records(r["seq"], r["len"], r["cell"], r["umi"], r["flags"])
        drec = torch.from_numpy(rec_np.view(np.int64)).to(dev)
        hrec_t = torch.from_numpy(rec_np.view(np.int64)).pin_memory()
        hrec = hrec_t.numpy().view(np.uint64)
        in_bytes = int(hrec.nbytes)
        push_dev = lambda: sess.push_packed(drec)
        push_host = lambda a, b: sess.push_packed(hrec[a:b])
        fmt_name = "packed records, %d B/read (hlac_count_push_packed)"
    else:
        read_len = int(r["len"][0])
        cell_bits = max(1, int(n_cells_total).bit_length())          # columns 0..n_cells-1 and the all-ones not-a-cell value
        umi_bits = max(1, int(r["umi"].max()).bit_length())
        rec_np = sb.pack_wire(r["seq"], r["len"], r["cell"], r["umi"], r["flags"], read_len, cell_bits, umi_bits)
        drec = torch.from_numpy(rec_np.view(np.int32)).to(dev)
        hrec_t = torch.from_numpy(rec_np.view(np.int32)).pin_memory()
        hrec = hrec_t.numpy().view(np.uint32)
        in_bytes = int(hrec.nbytes)
        push_dev = lambda: sess.push_wire(drec, read_len, cell_bits, umi_bits)
        push_host = lambda a, b: sess.push_wire(hrec[a:b], read_len, cell_bits, umi_bits)
        fmt_name = "wire records, %%d B/read (hlac_count_push_wire: %d-bp bases + primary + %d-bit cell + %d-bit UMI key)" % (read_len, cell_bits, umi_bits)

    def h2d_ceiling():
        """What the platform gives when every rank does nothing but copy its pinned input to its GPU at the same time: the
        ceiling of the end-to-end leg (GB/s per GPU; the slowest rank counts)."""
        if fmt == "soa":
            return None
        dst = torch.empty_like(drec)
        for _ in range(2):
            dst.copy_(hrec_t, non_blocking=True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(5):
            dst.copy_(hrec_t, non_blocking=True)
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t[0])
        del dst
        return in_bytes / (ms / 1e3) / 1e9

    def step_device():
        sess.reset()
        push_dev()
        return finish_step()

    def step_e2e():
        sess.reset()
        for a, b in zip(cuts[:-1], cuts[1:]):
            push_host(a, b)
        return finish_step()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            out = fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()
        launches0 = sess.timing()["launches"]   # of the last warm-up step (the session counts its own kernel launches per step)
        sampler = ClockSampler(local_rank)
        sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        for _ in range(steps):
            sess_out = fn()
        e1.record(stream)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        clocks = sampler.stop()
        ms = e0.elapsed_time(e1)
        tm = sess.timing()
        assert tm["launches"] == launches0, "kernel launches per step changed between warm-up and the timed steps"
        if world > 1:
            t = torch.tensor([ms, wall * 1e3], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, wall = float(t[0]), float(t[1]) / 1e3
        return ms, wall, tm, clocks, sess_out

    # NOTE sess.timing() accumulates since the last reset(): the last step's kernel times / launches are what it returns.
    ms_dev, wall_dev, tm_dev, clocks, out_dev = timed(step_device, args.steps, max(args.warmup, 3))
    if has_result:   # the result arrays are views of the session's pinned buffer: keep a copy across the next loop
        out_dev = (tuple(a.copy() for a in out_dev[0]), out_dev[1])
    ms_e2e, wall_e2e, tm_e2e, clocks_e2e, out_e2e = timed(step_e2e, args.steps, max(args.warmup, 3))
    ceiling = h2d_ceiling()
    ent, met = out_dev
    if has_result:
        assert all(np.array_equal(a, b) for a, b in zip(ent, out_e2e[0])), "device-resident and host-fed passes disagree"
    assert met == out_e2e[1], "metrics of the device-resident and host-fed passes disagree"
    gstats = sess.group_stats()
    n_entries, n_molecules = (int(len(ent[0])), int(ent[2].sum())) if has_result else (0, 0)
    d2h_own = (3 * 4 * n_entries + 17 * 8) if has_result else 17 * 8   # COO triplets + the result header (nnz + counters)
    if world > 1 and local:   # whole-job totals for the report (outside the timed regions): every rank holds its own column block
        t = torch.tensor([n_entries, n_molecules, d2h_own] + list(met), device=dev, dtype=torch.int64)
        dist.all_reduce(t)
        n_entries, n_molecules, d2h_own = int(t[0]), int(t[1]), int(t[2])
        met = [int(x) for x in t[3:]]

    # ---- in-run parity: the assembled matrix against the CPU oracle on whole cells ----
    from oracle import oracle as orc
    orc.build()
    parity = None
    if world > 1 and not local:
        # every rank runs the oracle on all reads of a few of ITS cells; rank 0 compares with the columns it assembled
        mine = np.unique(r["cell"][r["cell"] != synth.NOT_A_CELL])[:: max(1, (args.cells // max(args.parity_cells, 1)))][:args.parity_cells]
        m = np.isin(r["cell"], mine)
        _, ent_cpu = run_oracle({k: v[m] for k, v in r.items()}, int(m.sum()), 1, al[:2])
        gathered = [None] * world if rank == 0 else None
        dist.gather_object((mine.tolist(), ent_cpu), gathered, dst=0)
        if rank == 0:
            cols = np.array(sorted(c for g in gathered for c in g[0]))
            want = sorted(e for g in gathered for e in g[1])
            sel = np.isin(ent[1], cols)
            got = sorted(zip(ent[0][sel].tolist(), ent[1][sel].tolist(), ent[2][sel].tolist()))
            assert got == want, "assembled multi-GPU matrix differs from the oracle on the sampled cells"
            parity = {"checked_cells": int(len(cols)), "checked_entries": len(want), "ranks": world, "equal": True}

    total_reads = n * world
    per_step_ms = ms_dev / args.steps
    value = total_reads / (per_step_ms / 1e3)
    e2e_ms = ms_e2e / args.steps

    line = None
    if rank == 0:
        kern = {"map_ms": tm_dev["map_ms"], "group_consensus_ms": tm_dev["group_ms"], "coo_ms": tm_dev["coo_ms"]}
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        map_gbs = n * MAP_BYTES_PER_READ / (tm_dev["map_ms"] / 1e3) / 1e9 if tm_dev["map_ms"] > 0 else None
        traffic, traffic_src = None, "no ncu capture of this kernel version under profiles/ yet"
        try:
            tj = json.load(open(NCU_TRAFFIC_JSON))
            traffic = tj["dram_bytes_per_read"] * n
            traffic_src = "ncu --set full, %s: %.2f B/read measured x reads per launch" % (tj.get("source", NCU_TRAFFIC_JSON), tj["dram_bytes_per_read"])
        except Exception:
            pass
        step_gbs = n * STEP_BYTES_PER_READ / (per_step_ms / 1e3) / 1e9
        roof = {"bound": "hbm", "kernel": "k_count_map_q", "achieved": map_gbs, "peak": peak, "unit": "GB/s",
                "frac": (map_gbs / peak) if map_gbs else None, "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": peak_src, "algorithmic_bytes_per_read": MAP_BYTES_PER_READ, "launch_ms": tm_dev["map_ms"],
                "step": {"algorithmic_bytes_per_read": STEP_BYTES_PER_READ, "achieved": step_gbs, "frac": step_gbs / peak,
                         "note": "whole step (map + grouping + consensus + COO + the one host sync) on SURVEY.md 8d's 72 B/read, per GPU"},
                "note": "DRAM traffic ~ the algorithmic bytes; the kernel is latency/divergence-bound, not bandwidth-bound — DESIGN.md §5, profiles/"}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": per_step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u32", "data": "synthetic", "config": config,
                "e2e": {"value": total_reads / (e2e_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": in_bytes * world,
                        "d2h_bytes_per_step": d2h_own, "ms_per_step": e2e_ms, "h2d_ms": tm_e2e["h2d_ms"],
                        "h2d_gbs_per_gpu": in_bytes / (tm_e2e["h2d_ms"] / 1e3) / 1e9 if tm_e2e["h2d_ms"] > 0 else None,
                        "h2d_ceiling_gbs_per_gpu": ceiling,
                        "h2d_ceiling_note": "all %d ranks copying their pinned input and nothing else, slowest rank; the e2e leg cannot beat "
                                            "bytes / this (%.2f ms per step)" % (world, in_bytes / ceiling / 1e6) if ceiling else None,
                        "input_format": fmt_name % (in_bytes // n), "pushes_per_step": len(cuts) - 1,
                        "push_reads": [b - a for a, b in zip(cuts[:-1], cuts[1:])]},
                # counted by the session itself (hlac_count_timing): every kernel of ours launched in one step, times the steps.
                # No library kernel runs on the normal path (cub only sorts on the bucket-overflow fallback; fallbacks below).
                "gpu_launches": int(tm_dev["launches"]) * args.steps, "library_launches": 0 if gstats["fallbacks"] == 0 else None,
                "launches_per_step": int(tm_dev["launches"]), "group": gstats, "map_kernel_shape": sess.shape(),
                "kernels_ms_last_step": kern, "roofline": roof, "clocks": clocks, "clocks_e2e": clocks_e2e,
                "wall_ms_per_step": wall_dev / args.steps * 1e3,
                "result": {"matrix_entries": n_entries, "molecules": n_molecules, "metrics": met}, "cpus_bound": bound_cpus}
        if parity:
            line["parity_multi_gpu"] = parity
        if comm is not None:
            line["comm"] = comm.info
        # CPU baseline: the oracle (port) on a bounded sample of the same workload, one core like the reference. The sample is
        # all reads of the first k cells (+ the same share of non-whitelisted reads), so its entries are exactly the columns
        # < k of the full result: an in-run parity check for free.
        if world == 1:
            k = max(1, int(args.cells * min(1.0, args.cpu_sample / n)))
            m = cell_sample_mask(r, k, args.cells)
            ns = int(m.sum())
            dt, ent_cpu = run_oracle({kk: v[m] for kk, v in r.items()}, ns, 1, al[:2])
            sel = ent[1] < k
            got = list(zip(ent[0][sel].tolist(), ent[1][sel].tolist(), ent[2][sel].tolist()))
            assert got == ent_cpu, "GPU count matrix differs from the CPU oracle on the sampled cells"
            line["cpu_baseline"] = {"value": ns / dt, "unit": UNIT, "cores": 1, "kind": "port",
                                    "sample": "all %d reads of the first %d of %d cells (+ the same share of non-whitelisted reads), single "
                                              "thread (the reference is single-threaded); its %d matrix entries equal the GPU's columns < %d" % (ns, k, args.cells, len(ent_cpu), k)}
    want_geno = args.geno == "on" or (args.geno == "auto" and world == 1 and args.genes == 3 and args.reads == N_READS)
    if want_geno and rank == 0:
        sess.close()
        torch.cuda.empty_cache()
        try:
            line["geno"] = geno_record(args, local_rank, peak)
        except Exception as e:   # the counting line must survive a failure of the second record
            line["geno"] = {"error": "%s: %s" % (type(e).__name__, e)}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
