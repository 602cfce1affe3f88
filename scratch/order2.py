import pickle
d = pickle.load(open('/root/repo/scratch/idx.pkl','rb'))
(bvs, ranks), keys, vals = d['dbg_index']
seen=[]; s=set()
for n,o in vals:
    if n not in s: s.add(n); seen.append(n)
print('first-appearance order is identity:', seen==list(range(len(seen))), seen[:20])
print(vals[:10])
