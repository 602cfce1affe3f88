import pickle
d = pickle.load(open('/root/repo/scratch/idx.pkl','rb'))
st, ln, storage = d['start'], d['length'], d['storage']
def base(i): return (storage[i>>5] >> (62-2*(i&31))) & 3
seqs = [''.join('ACGT'[base(s+j)] for j in range(l)) for s,l in zip(st,ln)]
print(st[:10], ln[:10])
print('starts contiguous', all(st[i+1]==st[i]+ln[i] for i in range(len(st)-1)))
first = [s[:24] for s in seqs]
print('sorted by first kmer', first == sorted(first))
# order of appearance in input sequences
fa = {}
name=None
for line in open('/root/repo/tests/golden/ref_fixtures/fake_db/hla_nuc.fasta'):
    if line.startswith('>'): name=line.split()[1]; fa[name]=''
    else: fa[name]+=line.strip()
txs=[t.decode() for t in d['tx']]
# position of first occurrence (tx order, then pos)
def firstocc(s):
    for ti,t in enumerate(txs):
        p = fa[t].find(s)
        if p>=0: return (ti,p)
    return None
occ=[firstocc(s) for s in seqs]
print(occ[:20])
print('sorted by first occurrence', occ==sorted(occ))
print(d['data'][:20])
print('eq classes', d['eq'][:10])
