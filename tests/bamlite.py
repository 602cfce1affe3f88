"""Minimal pure-Python BGZF/BAM decoder used ONLY by the tests (independent of the product's C++ reader).

Yields records in file order — what htslib's `fetch` returns for a coordinate-sorted BAM restricted to a region
(reference: `BamSeqReader::next`, src/mapping.rs:234-278). Region filtering mirrors htslib overlap semantics:
a record overlaps [beg, end) iff pos < end and end_pos > beg, where end_pos comes from the CIGAR.
"""
import struct
import zlib

_SEQ = "=ACMGRSVTWYHKDBN"


def bgzf_inflate(path):
    data = open(path, "rb").read()
    out = []
    off = 0
    while off < len(data):
        # gzip member header: 10 fixed + XLEN(2) + extra; BSIZE in the 'BC' subfield
        xlen = struct.unpack_from("<H", data, off + 10)[0]
        extra = data[off + 12: off + 12 + xlen]
        bsize = None
        p = 0
        while p < len(extra):
            si1, si2, slen = extra[p], extra[p + 1], struct.unpack_from("<H", extra, p + 2)[0]
            if si1 == 66 and si2 == 67:
                bsize = struct.unpack_from("<H", extra, p + 4)[0] + 1
            p += 4 + slen
        assert bsize is not None
        cdata = data[off + 12 + xlen: off + bsize - 8]
        out.append(zlib.decompress(cdata, -15))
        off += bsize
    return b"".join(out)


def read_bam(path):
    """Returns (references [(name, len)], records list of dict)."""
    buf = bgzf_inflate(path)
    assert buf[:4] == b"BAM\x01"
    l_text = struct.unpack_from("<i", buf, 4)[0]
    off = 8 + l_text
    n_ref = struct.unpack_from("<i", buf, off)[0]
    off += 4
    refs = []
    for _ in range(n_ref):
        l_name = struct.unpack_from("<i", buf, off)[0]
        name = buf[off + 4: off + 4 + l_name - 1].decode()
        l_ref = struct.unpack_from("<i", buf, off + 4 + l_name)[0]
        refs.append((name, l_ref))
        off += 8 + l_name
    recs = []
    while off < len(buf):
        block_size = struct.unpack_from("<i", buf, off)[0]
        b = buf[off + 4: off + 4 + block_size]
        off += 4 + block_size
        ref_id, pos, l_read_name, mapq, bin_, n_cigar, flag, l_seq, nref, npos, tlen = struct.unpack_from(
            "<iiBBHHHiiii", b, 0)
        p = 32 + l_read_name
        cigar = struct.unpack_from("<%dI" % n_cigar, b, p)
        p += 4 * n_cigar
        sq = b[p: p + (l_seq + 1) // 2]
        p += (l_seq + 1) // 2
        seq = "".join(_SEQ[(sq[i >> 1] >> (4 if (i & 1) == 0 else 0)) & 15] for i in range(l_seq))
        p += l_seq  # qual
        tags = {}
        while p < len(b):
            tag = b[p: p + 2].decode()
            typ = chr(b[p + 2])
            p += 3
            if typ == "Z" or typ == "H":
                e = b.index(b"\x00", p)
                tags[tag] = b[p:e]
                p = e + 1
            elif typ in "AcC":
                tags[tag] = b[p]
                p += 1
            elif typ in "sS":
                p += 2
            elif typ in "iIf":
                p += 4
            elif typ == "B":
                sub = chr(b[p])
                cnt = struct.unpack_from("<i", b, p + 1)[0]
                p += 5 + cnt * {"c": 1, "C": 1, "s": 2, "S": 2, "i": 4, "I": 4, "f": 4}[sub]
            else:
                raise ValueError("bad aux type " + typ)
        ref_len = 0
        for c in cigar:
            op, ln = c & 15, c >> 4
            if op in (0, 2, 3, 7, 8):
                ref_len += ln
        recs.append(dict(ref_id=ref_id, pos=pos, flag=flag, seq=seq, tags=tags,
                         end=pos + max(ref_len, 1)))
    return refs, recs


def fetch(path, chrom, beg, end):
    """Records overlapping chrom:[beg,end) in file order; only those carrying CB and UB
    (src/mapping.rs:254-270). Yields (seq:str, cb:bytes, ub:bytes, primary:bool)."""
    refs, recs = read_bam(path)
    names = [r[0] for r in refs]
    tid = names.index(chrom)
    for r in recs:
        if r["ref_id"] != tid or not (r["pos"] < end and r["end"] > beg):
            continue
        if "CB" not in r["tags"] or "UB" not in r["tags"]:
            continue
        primary = not (r["flag"] & 0x100 or r["flag"] & 0x800)
        yield r["seq"], r["tags"]["CB"], r["tags"]["UB"], primary


def fetch_unmapped(path):
    """The `fetch_str(b"*")` pass (src/mapping.rs:378-379): records with no reference id."""
    _, recs = read_bam(path)
    for r in recs:
        if r["ref_id"] != -1:
            continue
        if "CB" not in r["tags"] or "UB" not in r["tags"]:
            continue
        primary = not (r["flag"] & 0x100 or r["flag"] & 0x800)
        yield r["seq"], r["tags"]["CB"], r["tags"]["UB"], primary
