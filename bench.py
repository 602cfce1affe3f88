#!/usr/bin/env python
"""bench.py — reads pseudoaligned + counted per second (BASELINE.json metric) on N B200s.

Workload (N=1): BASELINE config 2 — synthetic 5' GEX, 10 M reads x 91 bp, 10 k barcodes, HLA-A/B/C known-genotype
counting against the six fixture alleles (CDS + genomic). One step = one full pass of the hot path over the whole
read set: k_count_map (4-way lazy map_read + class -> gene/slot) -> radix sort of (cell, UMI) keys -> k_consensus ->
dense count matrix (-> NCCL all-reduce when N > 1) -> COO entries on the host.
  value : reads/s with the packed batch already resident in HBM (timed with CUDA events on the session stream).
  e2e   : the same through hlac_count_push with pinned HOST buffers: H2D of every input byte and D2H of the result
          inside the timed region.
N > 1: reads are partitioned by barcode, every rank maps its own 10 M-read shard (weak scaling), the
per-rank dense matrices have disjoint columns and are summed with one NCCL all-reduce.
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
# dram__bytes_read.sum + dram__bytes_write.sum of k_count_map per read, from the `ncu --set full` capture
# profiles/r01_count_map_v3_ncu_digest.txt (70.75 MB + 1.78 MB over 2 M reads); scaled to the reads of one launch
MAP_DRAM_BYTES_PER_READ_NCU = (70.790656e6 + 1.859072e6) / 2_000_000   # profiles/r01_count_map_v4_ncu_digest.txt


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
    ap.add_argument("--chunk", type=int, default=1_000_000, help="e2e push chunk in reads (0 = whole batch)")
    ap.add_argument("--genes", type=int, default=3, choices=[3, 4, 5, 6, 7, 8],
                    help="3 = the six fixture alleles (config 2); 8 = 2 synthetic alleles x 8 genes (config 4 shape)")
    ap.add_argument("--e2e-format", default="packed", choices=["packed", "soa"],
                    help="host input format of the end-to-end leg: packed 32-byte records (hlac_count_push_packed) or the five "
                         "SoA arrays (hlac_count_push, 35 B/read)")
    ap.add_argument("--merge", default="none", choices=["none", "allreduce"],
                    help="N > 1: 'none' = every rank keeps the columns of its own barcodes (shard-local cell indices, no "
                         "data-path collective); 'allreduce' = global cell indices, one NCCL all-reduce of the dense matrix, "
                         "rank 0 extracts the whole COO")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    config = {"workload": "%s: synthetic 5' GEX, %d reads x 91 bp per GPU, %d barcodes per GPU, %d CDS + %d genomic "
                          "alleles (%d genes x 2), known-genotype counting" % ("config2" if args.genes == 3 else "config4-shape",
                                                                               args.reads, args.cells, 2 * args.genes, 2 * args.genes, args.genes),
              "reads_per_gpu": args.reads, "barcodes_per_gpu": args.cells, "parallelism": "reads sharded by barcode x%d%s" % (world, "" if world == 1 else (
                  ", per-rank column blocks, no data-path collective" if args.merge == "none" else ", NCCL all-reduce of the dense matrix")),
              "l2": "inputs (%.0f MB/step) exceed the 126 MB L2; no explicit flush" % (args.reads * 35 / 1e6),
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
    if world > 1:
        # NCCL prints its version banner on stdout when the communicator is created (first collective); keep stdout
        # clean for the single JSON line by pointing fd 1 at stderr while that happens.
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.all_reduce(torch.zeros(1, device=dev))
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    local = world > 1 and args.merge == "none"
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
    nrows = sess.nrows

    class _Dense:   # torch view of the session's dense count matrix (for the NCCL all-reduce)
        def __init__(self, ptr, nelem):
            self.__cuda_array_interface__ = {"shape": (nelem,), "typestr": "<i4", "data": (ptr, False), "version": 2}

    def finish_step():
        ptr, nelem = sess.reduce()
        if world > 1 and not local:
            t = torch.as_tensor(_Dense(ptr, nelem), device=dev)
            dist.all_reduce(t)
        return sess.entries_arrays(copy=False) if (rank == 0 or local) else (None, None)

    # device-resident inputs
    d = {k: torch.from_numpy(v.view(np.int64) if v.dtype == np.uint64 else v.view(np.int16) if v.dtype == np.uint16
                             else v.view(np.int32) if v.dtype == np.uint32 else v).to(dev) for k, v in r.items()}

    def step_device():
        sess.reset()
        sess.push(d["seq"], d["len"], d["cell"], d["umi"], d["flags"])
        return finish_step()

    # pinned host inputs
    h = {k: torch.from_numpy(v.view(np.int64) if v.dtype == np.uint64 else v.view(np.int16) if v.dtype == np.uint16
                             else v.view(np.int32) if v.dtype == np.uint32 else v).pin_memory() for k, v in r.items()}
    hn = {k: v.numpy() for k, v in h.items()}
    hn = {"seq": hn["seq"].view(np.uint64), "len": hn["len"].view(np.uint16), "cell": hn["cell"].view(np.uint32),
          "umi": hn["umi"].view(np.uint32), "flags": hn["flags"]}
    chunk = args.chunk or n

    def step_e2e_soa():
        sess.reset()
        for a in range(0, n, chunk):
            b = min(n, a + chunk)
            sess.push(hn["seq"][a:b], hn["len"][a:b], hn["cell"][a:b], hn["umi"][a:b], hn["flags"][a:b])
        return finish_step()

    # the same batch as packed records (hlac_packed_batch: 32 B per 91-bp read, one H2D copy per push), in pinned memory
    hrec_t = torch.from_numpy(sb.pack_records(hn["seq"], hn["len"], hn["cell"], hn["umi"], hn["flags"]).view(np.int64)).pin_memory()
    hrec = hrec_t.numpy().view(np.uint64)

    def step_e2e_packed():
        sess.reset()
        for a in range(0, n, chunk):
            sess.push_packed(hrec[a:min(n, a + chunk)])
        return finish_step()

    step_e2e = step_e2e_packed if args.e2e_format == "packed" else step_e2e_soa

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            out = fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()
        sess.timing()
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
        if world > 1:
            t = torch.tensor([ms, wall * 1e3], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, wall = float(t[0]), float(t[1]) / 1e3
        return ms, wall, tm, clocks, sess_out

    # NOTE sess.timing() accumulates since the last reset(): the last step's kernel times are what it returns.
    ms_dev, wall_dev, tm_dev, clocks, out_dev = timed(step_device, args.steps, max(args.warmup, 3))
    if rank == 0 or local:   # the result arrays are views of the session's pinned buffer: keep a copy across the next loop
        out_dev = (tuple(a.copy() for a in out_dev[0]), out_dev[1])
    ms_e2e, wall_e2e, tm_e2e, clocks_e2e, out_e2e = timed(step_e2e, args.steps, max(args.warmup, 3))
    ent, met = out_dev
    if rank == 0 or local:
        assert all(np.array_equal(a, b) for a, b in zip(ent, out_e2e[0])), "device-resident and host-fed passes disagree"
    n_entries, n_molecules = (int(len(ent[0])), int(ent[2].sum())) if ent is not None else (0, 0)
    d2h_own = (3 * 4 * n_entries + 10 * 8 + 8 + 8) if ent is not None else 0   # COO triplets + counters + key count + nnz
    if world > 1:   # whole-job totals for the report (outside the timed regions)
        if local:       # every rank holds its own column block
            t = torch.tensor([n_entries, n_molecules, d2h_own], device=dev, dtype=torch.int64)
            dist.all_reduce(t)
            n_entries, n_molecules, d2h_own = int(t[0]), int(t[1]), int(t[2])
        t = torch.tensor(list(met if met is not None else sess.entries_arrays(copy=False)[1]), device=dev, dtype=torch.int64)
        dist.all_reduce(t)
        met = [int(x) for x in t]

    total_reads = n * world
    per_step_ms = ms_dev / args.steps
    value = total_reads / (per_step_ms / 1e3)
    e2e_ms = ms_e2e / args.steps
    # host-side gaps (sync for the key count, COO extraction) are inside both event-bracketed regions
    in_bytes = int(hrec.nbytes) if args.e2e_format == "packed" else sum(int(v.nbytes) for v in hn.values())
    d2h_bytes = d2h_own

    line = None
    if rank == 0:
        scored = met[0] - met[1] * 0 - met[2] - met[5]
        nkeys = n_molecules
        kern = {"map_ms": tm_dev["map_ms"], "sort_ms": tm_dev["sort_ms"], "consensus_ms": tm_dev["consensus_ms"]}
        # dominant kernel of OURS: the map kernel (the cub sort is library code and reported beside it)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        map_gbs = n * MAP_BYTES_PER_READ / (tm_dev["map_ms"] / 1e3) / 1e9 if tm_dev["map_ms"] > 0 else None
        roof = {"bound": "hbm", "kernel": "k_count_map", "achieved": map_gbs, "peak": peak, "unit": "GB/s",
                "frac": (map_gbs / peak) if map_gbs else None, "traffic": MAP_DRAM_BYTES_PER_READ_NCU * n,
                "traffic_source": "ncu --set full, profiles/r01_count_map_v4_ncu_digest.txt (bytes/read x reads per launch)",
                "peak_source": peak_src,
                "algorithmic_bytes_per_read": MAP_BYTES_PER_READ, "launch_ms": tm_dev["map_ms"],
                "note": "DRAM traffic equals the algorithmic bytes; the kernel is latency-bound (63 % issue-active, 14.7 of 32 lanes "
                        "active per instruction) — DESIGN.md §5, profiles/"}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": per_step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u32", "data": "synthetic", "config": config,
                "e2e": {"value": total_reads / (e2e_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": in_bytes * world,
                        "d2h_bytes_per_step": d2h_bytes, "ms_per_step": e2e_ms, "h2d_ms": tm_e2e["h2d_ms"],
                        "input_format": "packed records, %d B/read (hlac_count_push_packed)" % (in_bytes // n) if args.e2e_format == "packed"
                        else "five SoA arrays, %d B/read (hlac_count_push)" % (in_bytes // n)},
                # our own kernels in the timed region per step (device-resident loop); the radix sort between them is
                # cub (library): histogram + exclusive-sum + one onesweep pass per 8 key bits
                "gpu_launches": 5 * args.steps,   # k_count_map, k_consensus, k_coo_count, k_coo_scan, k_coo_write
                "library_launches": (2 + (int(r["umi"].max()).bit_length() + max(1, (n_cells_total - 1).bit_length()) + 7) // 8) * args.steps,
                "kernels_ms_last_step": kern, "roofline": roof, "clocks": clocks, "clocks_e2e": clocks_e2e,
                "wall_ms_per_step": wall_dev / args.steps * 1e3,
                "result": {"matrix_entries": n_entries, "molecules": nkeys, "metrics": met}, "cpus_bound": bound_cpus}
        # CPU baseline: the oracle (port) on a bounded sample of the same workload, one core like the reference
        if world == 1:
            from oracle import oracle as orc
            orc.build()
            ns = min(n, args.cpu_sample)
            dt, ent_cpu = run_oracle(r, ns, 1, al[:2])
            line["cpu_baseline"] = {"value": ns / dt, "unit": UNIT, "cores": 1, "kind": "port",
                                    "sample": "first %d reads of rank 0's shard, single thread (the reference is single-threaded)" % ns}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
