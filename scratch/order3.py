import pickle
d = pickle.load(open('/root/repo/scratch/idx.pkl','rb'))
st, ln, storage, data = d['start'], d['length'], d['storage'], d['data']
def base(i): return (storage[i>>5] >> (62-2*(i&31))) & 3
kms=[]
for n,(s,l) in enumerate(zip(st,ln)):
    v=0
    for j in range(l):
        v=((v<<2)|base(s+j)) & ((1<<48)-1)
        if j>=23: kms.append((v,n,j-23))
kms.sort()
seen=[];ss=set()
for v,n,o in kms:
    c=data[n]
    if c not in ss: ss.add(c); seen.append(c)
print('ascending kmer order -> colour first appearance identity:', seen==list(range(len(seen))), seen[:20])
# alternative: canonical? stranded so no. try descending, or bucket by low bits
def test(keyf,name):
    seen=[];ss=set()
    for v,n,o in sorted(kms,key=keyf):
        c=data[n]
        if c not in ss: ss.add(c); seen.append(c)
    print(name, seen==list(range(len(seen))), seen[:12])
test(lambda t:-t[0],'desc')
