"""Tiny driver for ncu: genotyping pushes + finish + EM at config-3 scale. Usage: python tools/prof_geno.py [alleles_per_gene] [reads]"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import schlacount_b200 as sb  # noqa: E402
from tools import synth  # noqa: E402

npg = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2_000_000
db = tempfile.mkdtemp(prefix="hladb_")
names = synth.make_db(db, npg, 20191104)
donor = [names[3], names[npg // 2], names[npg + 5], names[npg + 5], names[2 * npg + 1], names[3 * npg - 1]]
r = synth.reads_for_db(n, db, donor, 10000, 20191104)
gix = sb.Index.from_fasta(os.path.join(db, "hla_nuc.fasta"), os.path.join(db, "Allele_status.txt"))
view = {np.dtype("uint64"): np.int64, np.dtype("uint16"): np.int16, np.dtype("uint32"): np.int32, np.dtype("uint8"): np.uint8}
d = {k: torch.from_numpy(v.view(view[v.dtype])).cuda() for k, v in r.items()}
for _ in range(2):
    g = sb.GenoSession(gix, lean=True)
    g.push(d["seq"], d["len"], d["cell"], d["umi"])
    t = g.finish(raw=True)
    theta, iters = g.squarem()
    c = g.call(theta)
print("ok", t, iters, g.timing(), g.stats())
