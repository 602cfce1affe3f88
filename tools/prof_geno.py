"""Tiny driver for ncu / HLAC_DEBUG_TIMING: genotyping pushes + finish + EM + caller at config-3 scale (8 genes).
Usage: python tools/prof_geno.py [alleles_per_gene=3750] [reads=2000000]"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import schlacount_b200 as sb  # noqa: E402
from tools import synth  # noqa: E402

npg = int(sys.argv[1]) if len(sys.argv) > 1 else 3750
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2_000_000
db = tempfile.mkdtemp(prefix="hladb_")
seed = 20191101 + 3   # the same DB and donor as bench.py's config-3 record
names = synth.make_db(db, npg, seed, extra_genes=("DRB1", "DPA1", "DPB1", "DQA1", "DQB1"))
donor = [names[g * npg + o] for g in range(8) for o in ((3 + 17 * g) % npg, (npg // 2 + 29 * g) % npg)][:12] + [names[6 * npg + 5], names[7 * npg + 11]]
r = synth.reads_for_db(n, db, donor, 10000, seed)
gix = sb.Index.from_fasta(os.path.join(db, "hla_nuc.fasta"), os.path.join(db, "Allele_status.txt"))
view = {np.dtype("uint64"): np.int64, np.dtype("uint16"): np.int16, np.dtype("uint32"): np.int32, np.dtype("uint8"): np.uint8}
d = {k: torch.from_numpy(v.view(view[v.dtype])).cuda() for k, v in r.items()}
import time  # noqa: E402
for rep in range(3):
    g = sb.GenoSession(gix, lean=True)
    torch.cuda.synchronize()
    t0 = time.time()
    g.push(d["seq"], d["len"], d["cell"], d["umi"])
    torch.cuda.synchronize()
    t1 = time.time()
    t = g.finish(raw=True)
    t2 = time.time()
    theta, iters = g.squarem()
    t3 = time.time()
    c = g.call(theta)
    t4 = time.time()
    print("rep %d: push %.1f ms, finish %.1f ms, squarem %.1f ms (%d iterations), call %.1f ms" % (rep, 1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), iters, 1e3 * (t4 - t3)), file=sys.stderr)
print("ok", t, iters, g.timing(), g.stats(), g.resolve_stats())
