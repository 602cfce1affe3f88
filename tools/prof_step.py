"""Tiny driver for ncu: a few device-resident counting steps (no CPU baseline, no e2e). Usage:
   python tools/prof_step.py [reads] [steps]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import schlacount_b200 as sb  # noqa: E402
from tools import synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
cds, gen = synth.fixture_alleles()
r = synth.make_reads(n, [s for _, s in cds], [s for _, s in gen], 10000, 20191103)
cf, gf = os.path.join(synth.FIX, "cds_ABC.fa"), os.path.join(synth.FIX, "genomic_ABC.fa")
gc, gg = sb.Index.from_fasta(cf), sb.Index.from_fasta(gf)
s = sb.CountSession(gc, gg, 10000)
view = {np.dtype("uint64"): np.int64, np.dtype("uint16"): np.int16, np.dtype("uint32"): np.int32, np.dtype("uint8"): np.uint8}
fmt = os.environ.get("HLAC_PROF_FORMAT", "wire")
if fmt == "wire":
    drec = torch.from_numpy(sb.pack_wire(r["seq"], r["len"], r["cell"], r["umi"], r["flags"], 91, 14, 21).view(np.int32)).cuda()
else:
    drec = torch.from_numpy(sb.pack_records(r["seq"], r["len"], r["cell"], r["umi"], r["flags"]).view(np.int64)).cuda()
s.reserve(n)
ms = []
for _ in range(steps):
    s.reset()
    if fmt == "wire":
        s.push_wire(drec, 91, 14, 21)
    else:
        s.push_packed(drec)
    ent, met = s.finish_arrays(copy=False)
    ms.append(s.timing()["map_ms"])
print("ok", len(ent[0]), met, s.timing(), s.probe_stats(), "|", s.shape(), "| map_ms min %.4f" % min(ms[1:] or ms))
