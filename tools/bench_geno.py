"""Config 3 (genotyping: pseudoalignment vs a synthetic IMGT-like DB + EM + calls) timing on one GPU.
   python tools/bench_geno.py [alleles_per_gene=10000] [reads=2000000] [check_reads=20000]
Prints one JSON line. Not the driver's bench (bench.py is config 2); numbers go to DESIGN.md / profiles/."""
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import schlacount_b200 as sb  # noqa: E402
from tools import synth  # noqa: E402

npg = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2_000_000
ncheck = int(sys.argv[3]) if len(sys.argv) > 3 else 20000
lean = (int(sys.argv[4]) if len(sys.argv) > 4 else 1) != 0   # leave counts_reads on the device (hlac_geno_set_lean)
seed = 20191101 + 3
db = tempfile.mkdtemp(prefix="hladb_")
t0 = time.time()
names = synth.make_db(db, npg, seed)
donor = [names[3], names[npg // 2], names[npg + 5], names[npg + 5], names[2 * npg + 1], names[3 * npg - 1]]
r = synth.reads_for_db(n, db, donor, 10000, seed)
t_gen = time.time() - t0
fa, st = os.path.join(db, "hla_nuc.fasta"), os.path.join(db, "Allele_status.txt")
t0 = time.time()
gix = sb.Index.from_fasta(fa, st)
t_build = time.time() - t0
view = {np.dtype("uint64"): np.int64, np.dtype("uint16"): np.int16, np.dtype("uint32"): np.int32, np.dtype("uint8"): np.uint8}
d = {k: torch.from_numpy(v.view(view[v.dtype])).cuda() for k, v in r.items()}
out = {"alleles": len(names), "reads": n, "lean": lean, "index": gix.info, "gen_s": t_gen, "index_build_s": t_build}
for rep in range(2):
    g = sb.GenoSession(gix, lean=lean)
    torch.cuda.synchronize()
    t0 = time.time()
    g.push(d["seq"], d["len"], d["cell"], d["umi"])
    t_push = time.time() - t0
    t0 = time.time()
    tab = g.finish(raw=True)
    t_fin = time.time() - t0
    t0 = time.time()
    theta, iters = g.squarem()
    t_em = time.time() - t0
    t0 = time.time()
    call = g.call(theta)
    t_call = time.time() - t0
    out.update(push_s=t_push, finish_s=t_fin, em_s=t_em, em_iters=iters, call_s=t_call, timing=g.timing(), stats=g.stats(),
               tables=tab,
               called=[gix.tx_names[i] for i in call["called"]], donor=donor,
               reads_per_s_map=n / (g.timing()["map_ms"] / 1e3), reads_per_s_push=n / t_push)
if ncheck:
    from oracle import oracle as orc
    orc.build()
    oix = orc.Index.from_fasta(fa, st)
    o = orc.Geno(oix)
    t0 = time.time()
    ci, cv = o.push(r["seq"][:ncheck], r["len"][:ncheck], r["cell"][:ncheck], r["umi"][:ncheck])
    out["oracle_reads_per_s"] = ncheck / (time.time() - t0)
    o.finish()
    g2 = sb.GenoSession(gix, record_reads=True)
    g2.push(r["seq"][:ncheck], r["len"][:ncheck], r["cell"][:ncheck], r["umi"][:ncheck])
    t2 = g2.finish()
    out["check"] = bool(np.array_equal(g2.last_cov(), cv) and np.array_equal(g2.read_class_ids(ncheck), ci)
                        and t2["counts_umi"] == o.table("umi") and t2["counts_reads"] == o.table("reads")
                        and t2["reads_explained"] == o.reads_explained().tolist())
print(json.dumps(out))
