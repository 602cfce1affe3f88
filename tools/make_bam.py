"""Synthetic coordinate-sorted BAM (+ .bai) with CB/UB tags, for timing the host BAM reader (bam_reader.cpp) and the
CLI end to end at sizes the fixture does not reach. Records look like 10x 5' GEX alignments on contig "6": 91-bp reads,
the tag set Cell Ranger writes (CB, UB plus the ones the reader has to skip), a few records without CB/UB and an
unmapped tail. The index puts every chunk in bin 0 (always selected by reg2bin) and points every linear-index window at the first record,
which is all fetch() needs (a valid, if unselective, index).
   python tools/make_bam.py out.bam [n_records=1000000] [n_unmapped=0] [n_on_contig_1=0]
The optional records on contig "1" precede the selection and have to be skipped by a fetch on "6" (the linear index
points at the start of the file)."""
import os
import struct
import sys
import zlib

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools import synth  # noqa: E402

NT16 = {c: i for i, c in enumerate("=ACMGRSVTWYHKDBN")}


def bgzf_block(data, level=1):
    c = zlib.compressobj(level, zlib.DEFLATED, -15)
    comp = c.compress(data) + c.flush()
    bsize = len(comp) + 25
    return (b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", bsize) + comp
            + struct.pack("<II", zlib.crc32(data) & 0xFFFFFFFF, len(data)))


class BgzfWriter:
    def __init__(self, path):
        self.f = open(path, "wb")
        self.buf = bytearray()
        self.coffset = 0

    def tell(self):   # virtual offset of the next byte written
        return (self.coffset << 16) | len(self.buf)

    def write(self, b):
        self.buf += b
        while len(self.buf) >= 0xFF00:
            self._flush(0xFF00)

    def _flush(self, n):
        blk = bgzf_block(bytes(self.buf[:n]))
        self.f.write(blk)
        self.coffset += len(blk)
        del self.buf[:n]

    def close(self):
        if self.buf:
            self._flush(len(self.buf))
        self.f.write(bgzf_block(b""))   # EOF marker
        self.f.close()


def record(tid, pos, flag, seq, name, cb, ub, extra_aux=b""):
    l = len(seq)
    packed = bytearray((l + 1) // 2)
    for i, ch in enumerate(seq):
        packed[i >> 1] |= NT16[ch] << (0 if i & 1 else 4)
    aux = b"NHC\x01" + b"HIC\x01" + b"ASC\x59" + b"nMC\x00" + extra_aux + b"CRZ" + (cb or b"N" * 16)[:16] + b"\0" + b"CYZ" + b"F" * 16 + b"\0"
    if cb is not None:
        aux += b"CBZ" + cb + b"\0"
    aux += b"URZ" + (ub or b"ACGTACGTAC") + b"\0" + b"UYZ" + b"F" * 10 + b"\0"
    if ub is not None:
        aux += b"UBZ" + ub + b"\0"
    aux += b"RGZsample:0:1:HXXXX:1\0"
    cigar = struct.pack("<I", (l << 4) | 0) if tid >= 0 else b""
    body = (struct.pack("<iiBBHHHiiii", tid, pos, len(name) + 1, 255 if tid >= 0 else 0, 4681, len(cigar) // 4, flag, l, -1, -1, 0)
            + name + b"\0" + cigar + bytes(packed) + b"\xff" * l + aux)
    return struct.pack("<i", len(body)) + body


def main():
    out = sys.argv[1]
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
    n_un = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    n_before = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    rng = np.random.default_rng(20191101)
    cds, gen = synth.fixture_alleles()
    r = synth.make_reads(n + n_un, [s for _, s in cds], [s for _, s in gen], 5000, 20191101, frac_not_cell=0.0)
    bases = np.array(list("ACGT"))
    w = BgzfWriter(out)
    refs = [("1", 248956422), ("6", 170805979)]
    hdr = b"@HD\tVN:1.4\tSO:coordinate\n" + b"".join(b"@SQ\tSN:%s\tLN:%d\n" % (nm.encode(), ln) for nm, ln in refs)
    w.write(b"BAM\x01" + struct.pack("<i", len(hdr)) + hdr + struct.pack("<i", len(refs)))
    for nm, ln in refs:
        w.write(struct.pack("<i", len(nm) + 1) + nm.encode() + b"\0" + struct.pack("<i", ln))
    first = w.tell()
    for i in range(n_before):
        w.write(record(0, 1000 + 50 * i, 0, "ACGT" * 22 + "ACG", b"b%d" % i, b"AAACCCGGGTTTACGT-1", b"ACGTACGTAC"))
    pos = np.sort(rng.integers(29_900_000, 31_400_000, size=n))
    for i in range(n + n_un):
        codes = [(int(r["seq"][i][k >> 5]) >> (62 - 2 * (k & 31))) & 3 for k in range(91)]
        seq = "".join(bases[codes])
        cell = int(r["cell"][i])
        cb = b"".join(b"ACGT"[(cell >> (2 * k)) & 3:((cell >> (2 * k)) & 3) + 1] for k in range(16)) + b"-1"
        u = int(r["umi"][i])
        ub = b"".join(b"ACGT"[(u >> (2 * k)) & 3:((u >> (2 * k)) & 3) + 1] for k in range(9, -1, -1))
        if i % 97 == 0:
            ub = None        # records without UB are dropped by the reader (mapping.rs:254-270)
        flag = 0 if r["flags"][i] else 256
        if i == n:
            mapped_end = w.tell()
        if i < n:
            w.write(record(1, int(pos[i]), flag | (16 if i & 1 else 0), seq, b"r%d" % i, cb, ub))
        else:
            w.write(record(-1, -1, 4, seq, b"u%d" % i, cb, ub))
    if n_un == 0:
        mapped_end = w.tell()
    w.close()
    # .bai: per reference n_bin, bins (bin id, n_chunk, chunks), n_intv, ioffsets; everything in bin 0 of reference 1 ("6")
    with open(out + ".bai", "wb") as f:
        f.write(b"BAI\x01" + struct.pack("<i", len(refs)))
        if n_before:                                                        # reference "1": its records, bin 0 + linear index
            f.write(struct.pack("<i", 1) + struct.pack("<Ii", 0, 1) + struct.pack("<QQ", first, first) + struct.pack("<i", 1) + struct.pack("<Q", first))
        else:
            f.write(struct.pack("<ii", 0, 0))                               # reference "1": nothing
        n_intv = (int(pos[-1]) >> 14) + 1   # linear index: every 16 kb window starts at the first record (a valid lower bound)
        f.write(struct.pack("<i", 1) + struct.pack("<Ii", 0, 1) + struct.pack("<QQ", first, mapped_end) + struct.pack("<i", n_intv))
        f.write(struct.pack("<%dQ" % n_intv, *([first] * n_intv)))
        f.write(struct.pack("<Q", n_un))
    print("wrote %s: %d mapped + %d unmapped records, %.1f MB" % (out, n, n_un, os.path.getsize(out) / 1e6))


if __name__ == "__main__":
    main()
