"""Synthetic 5' GEX-like inputs for the benchmark and the large parity tests (BASELINE.json configs 2-5).

IMGT/HLA is not available offline, so allele sets are derived from the six fixture alleles
(tests/golden/ref_fixtures/{cds_ABC.fa,genomic_ABC.fa}); reads follow SURVEY.md §8(d): 91 bp, ~70 % from CDS,
~5 % intronic (genomic only), ~25 % off-target random sequence, 0.3 % substitution error, random strand, 1-3 reads
per UMI, 10-bp UMIs, a few percent of reads from barcodes outside the whitelist, ~7 % non-primary records.
Everything is seeded; nothing here touches /root/reference.
"""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIX = os.path.join(ROOT, "tests", "golden", "ref_fixtures")
READ_LEN = 91
NOT_A_CELL = 0xFFFFFFFF
_CODE = np.full(256, 0, dtype=np.uint8)
for _i, _c in enumerate(b"ACGT"):
    _CODE[_c] = _i
    _CODE[ord(chr(_c).lower())] = _i


def read_fasta(path):
    recs = []
    for line in open(path):
        line = line.rstrip("\n")
        if line.startswith(">"):
            recs.append([line[1:], []])
        elif line:
            recs[-1][1].append(line)
    return [(h, "".join(s)) for h, s in recs]


def fixture_alleles():
    """-> (cds [(header, seq)], genomic [(header, seq)]) of the six fixture alleles"""
    return read_fasta(os.path.join(FIX, "cds_ABC.fa")), read_fasta(os.path.join(FIX, "genomic_ABC.fa"))


def write_fasta(path, recs):
    with open(path, "w") as f:
        for h, s in recs:
            f.write(">%s\n" % h)
            for i in range(0, len(s), 60):
                f.write(s[i:i + 60] + "\n")


def _codes(s):
    return _CODE[np.frombuffer(s.encode(), dtype=np.uint8)]


def pack_codes(codes):
    """uint8 [n, L] 2-bit codes -> uint64 [n, ceil(L/32)] MSB-first (hlac_batch.seq layout)"""
    n, L = codes.shape
    wpr = (L + 31) // 32
    pad = np.zeros((n, wpr * 32), dtype=np.uint8)
    pad[:, :L] = codes
    sh = (62 - 2 * np.arange(32, dtype=np.uint64)).astype(np.uint64)
    return (pad.reshape(n, wpr, 32).astype(np.uint64) << sh).sum(axis=2, dtype=np.uint64)


def make_reads(n, cds_seqs, gen_seqs, n_cells, seed, frac_cds=0.70, frac_gen=0.05, err=0.003, frac_not_cell=0.03,
               frac_nonprimary=0.07, umi_len=10, chunk=1 << 20):
    """-> dict(seq u64[n,3], len u16[n], cell u32[n], umi u32[n], flags u8[n]). cds_seqs/gen_seqs: ASCII strings, allele i
    of both lists being the same allele. Reads of one molecule share (cell, umi, allele, strand)."""
    rng = np.random.default_rng(seed)
    L = READ_LEN
    cds = [_codes(s) for s in cds_seqs]
    gen = [_codes(s) for s in gen_seqs]
    cat = np.concatenate(cds + gen)
    offs = np.cumsum([0] + [len(x) for x in cds + gen])[:-1]
    lens = np.array([len(x) for x in cds + gen])
    na = len(cds)
    out = dict(seq=np.zeros((n, (L + 31) // 32), np.uint64), len=np.full(n, L, np.uint16), cell=np.zeros(n, np.uint32),
               umi=np.zeros(n, np.uint32), flags=np.zeros(n, np.uint8))
    done = 0
    while done < n:
        m = min(chunk, n - done)
        # molecules: 1-3 reads each
        nmol = m  # upper bound
        per = rng.integers(1, 4, size=nmol)
        mol = np.repeat(np.arange(nmol), per)[:m]
        nmol = int(mol[-1]) + 1
        kind = rng.random(nmol)
        allele = rng.integers(0, na, size=nmol)
        src = np.where(kind < frac_cds, allele, allele + na)          # cds or genomic copy of the allele
        offt = kind >= frac_cds + frac_gen                              # off-target molecule
        # 5'-biased anchor for cds molecules, uniform for genomic
        u = rng.random(nmol)
        anchor = np.where(kind < frac_cds, (u ** 1.5) * (lens[src] - L), u * (lens[src] - L)).astype(np.int64)
        strand = rng.integers(0, 2, size=nmol)
        cell = rng.integers(0, n_cells, size=nmol).astype(np.uint32)
        cell[rng.random(nmol) < frac_not_cell] = NOT_A_CELL
        umi = (rng.integers(0, 1 << (2 * umi_len), size=nmol, dtype=np.uint64) | (1 << (2 * umi_len))).astype(np.uint32)
        # reads
        jit = rng.integers(-40, 41, size=m)
        start = np.clip(anchor[mol] + jit, 0, lens[src[mol]] - L)
        idx = (offs[src[mol]] + start)[:, None] + np.arange(L)[None, :]
        codes = cat[idx]
        ro = offt[mol]
        nro = int(ro.sum())
        if nro:
            codes[ro] = rng.integers(0, 4, size=(nro, L), dtype=np.uint8)
        nerr = rng.binomial(m * L, err)
        er, ec = rng.integers(0, m, size=nerr), rng.integers(0, L, size=nerr)
        codes[er, ec] = (codes[er, ec] + rng.integers(1, 4, size=nerr, dtype=np.uint8)) & 3
        rs = strand[mol] == 1
        codes[rs] = 3 - codes[rs][:, ::-1]
        out["seq"][done:done + m] = pack_codes(codes)
        out["cell"][done:done + m] = cell[mol]
        out["umi"][done:done + m] = umi[mol]
        out["flags"][done:done + m] = (rng.random(m) >= frac_nonprimary).astype(np.uint8)
        done += m
    return out


def config2(n_reads=10_000_000, n_cells=10_000, seed=20191101 + 2):
    """BASELINE config 2: 10 M reads, 10 k barcodes, the six fixture alleles (HLA-A/B/C x 2). -> (cds_fa, gen_fa, reads)"""
    cds, gen = fixture_alleles()
    reads = make_reads(n_reads, [s for _, s in cds], [s for _, s in gen], n_cells, seed)
    return os.path.join(FIX, "cds_ABC.fa"), os.path.join(FIX, "genomic_ABC.fa"), reads


# ---- synthetic IMGT-like databases (configs 3-5) --------------------------------------------------------------------
def _mutate(rng, seq, n_sub, lo, hi):
    s = bytearray(seq.encode())
    for p in rng.integers(lo, hi, size=n_sub):
        s[p] = b"ACGT"[(b"ACGT".index(s[p]) + int(rng.integers(1, 4))) % 4]
    return s.decode()


def make_db(out_dir, n_per_gene, seed, genes=("A", "B", "C"), extra_genes=(), extra_div=8):
    """Writes hla_nuc.fasta, hla_gen.fasta, Allele_status.txt for a synthetic DB derived from the six fixture alleles.
    Every allele is a fixture allele of its gene with 1-6 substitutions concentrated in the exon 2-3 stretch of the CDS
    (positions 74-620) and, independently, in the genomic copy. `extra_genes` (e.g. DRB1, DQA1 ...) are fabricated from
    the A/B/C templates by ~12 % divergence (1 / extra_div of the bases substituted) so that they behave as separate genes
    that still share some k-mers with their template. Names follow src/hla.rs:49.
    -> list of allele names in FASTA order."""
    os.makedirs(out_dir, exist_ok=True)
    rng = np.random.default_rng(seed)
    cds, gen = fixture_alleles()
    by_gene = {}
    for (h, s), (_, g) in zip(cds, gen):
        by_gene.setdefault(h.split(" ")[1].split("*")[0], []).append((s, g))
    templates = dict(by_gene)
    for k, eg in enumerate(extra_genes):
        base = by_gene[("A", "B", "C")[k % 3]]
        templates[eg] = [(_mutate(rng, s, len(s) // extra_div, 0, len(s)), _mutate(rng, g, len(g) // extra_div, 0, len(g))) for s, g in base]
    nuc, gn, status, names = [], [], [], []
    hid = 10000
    for gene in list(genes) + list(extra_genes):
        for i in range(n_per_gene):
            s, g = templates[gene][i % len(templates[gene])]
            if i >= len(templates[gene]):
                k = int(rng.integers(1, 7))
                s = _mutate(rng, s, k, 74, min(620, len(s)))
                g = _mutate(rng, g, k, 200, len(g) - 200)
            f1, f2, f3 = 1 + i // 400, 1 + (i // 20) % 20, 1 + i % 20
            name = "%s*%02d:%02d:%02d" % (gene, f1, f2, f3)
            hid += 1
            nuc.append(("HLA:HLA%05d %s %d bp" % (hid, name, len(s)), s))
            gn.append(("HLA:HLA%05d %s %d bp" % (hid, name, len(g)), g))
            status.append("%s,1,1,Confirmed,1,%d,Full,gDNA" % (name, len(s)))
            names.append(name)
    write_fasta(os.path.join(out_dir, "hla_nuc.fasta"), nuc)
    write_fasta(os.path.join(out_dir, "hla_gen.fasta"), gn)
    with open(os.path.join(out_dir, "Allele_status.txt"), "w") as f:
        f.write("# file: Allele_status.txt (synthetic)\nAllele,Cells,Groups,Confirmed,Start,End,Partial,Type\n")
        f.write("\n".join(status) + "\n")
    return names


def reads_for_db(n, db_dir, donor, n_cells, seed, **kw):
    """Reads simulated from the `donor` alleles (names) of a synthetic DB (genotyping input)."""
    nuc = {h.split(" ")[1]: s for h, s in read_fasta(os.path.join(db_dir, "hla_nuc.fasta"))}
    gn = {h.split(" ")[1]: s for h, s in read_fasta(os.path.join(db_dir, "hla_gen.fasta"))}
    return make_reads(n, [nuc[a] for a in donor], [gn[a] for a in donor], n_cells, seed, **kw)
