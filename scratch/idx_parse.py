import struct, sys
raw = open('/root/repo/tests/golden/ref_fixtures/fake_db/hla_nuc.fasta.idx','rb').read()
pos = 0
def u64():
    global pos; v = struct.unpack_from('<Q', raw, pos)[0]; pos += 8; return v
def u32():
    global pos; v = struct.unpack_from('<I', raw, pos)[0]; pos += 4; return v
def u8():
    global pos; v = raw[pos]; pos += 1; return v
def vec(f):
    n = u64(); return [f() for _ in range(n)]
storage = vec(u64); nb = u64(); start = vec(u64); length = vec(u32); exts = vec(u8); data = vec(u32); stranded = u8()
def mphf():
    bitvecs = []
    n = u64()
    for _ in range(n):
        bits = u64(); v = vec(u64); bitvecs.append((bits, v))
    ranks = vec(lambda: vec(u64))
    return bitvecs, ranks
def boom(valf, keys=True):
    m = mphf()
    k = vec(u64) if keys else None
    v = vec(valf)
    return m, k, v
left = boom(u32); right = boom(u32)
eq = vec(lambda: vec(u32))
dbg_index = boom(lambda: (u32(), u32()), keys=False)
def string():
    global pos; n = u64(); s = raw[pos:pos+n]; pos += n; return s
tx = vec(string)
nmap = u64()
print('EOF ok', pos == len(raw), 'nodes', len(start), 'bases', nb, 'tx', tx, 'map', nmap)
for name, b in (('left', left), ('right', right), ('index', dbg_index)):
    (bvs, ranks), k, v = b
    print(name, 'levels', [(bits, len(w)) for bits, w in bvs], 'ranks', [len(r) for r in ranks], 'nkeys', len(k) if k else None, 'nvals', len(v))
    print('  ranks0 head', ranks[0][:4])
import pickle
pickle.dump(dict(storage=storage, nb=nb, start=start, length=length, exts=exts, data=data, left=left, right=right, eq=eq, dbg_index=dbg_index, tx=tx), open('/root/repo/scratch/idx.pkl','wb'))
