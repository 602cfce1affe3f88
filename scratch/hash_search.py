import pickle, itertools
d = pickle.load(open('/root/repo/scratch/idx.pkl','rb'))
M = (1<<64)-1
P0,P1,P2,P3,P4,P5 = 0xa0761d6478bd642f,0xe7037ed1a0b428db,0x8ebc6af09c88c6e3,0x589965cc75374cc3,0x1d8e4e27c47d124f,0xeb44accab455d165
def wymum(a,b):
    r = (a&M)*(b&M); return ((r>>64) ^ r) & M
def swap(x): return ((x & 0xFFFFFFFF) << 32) | (x >> 32)
(bvs, ranks), keys, vals = d['left']
def contains(bv, i): return (bv[1][i>>6] >> (i&63)) & 1
def rank_of(level, idx):
    # rank = number of set bits before idx in this level + all previous levels
    pop = 0
    for l in range(level): pop += sum(bin(w).count('1') for w in bvs[l][1])
    w = bvs[level][1]
    pop += sum(bin(w[j]).count('1') for j in range(idx>>6)) + bin(w[idx>>6] & ((1<<(idx&63))-1)).count('1')
    return pop
cands = {}
for seedf_name, seedf in (('1<<2i', lambda i: 1<<(2*i)), ('i', lambda i: i), ('1<<i', lambda i: 1<<i)):
  for rd_name, rd in (('swap', swap), ('plain', lambda x: x)):
    for core_name, core in (('mum(s^P0,x^P1)', lambda s,x: wymum(s^P0, x^P1)), ('mum(x^s^P0,P1)', lambda s,x: wymum(x^s^P0, P1)), ('mum(s^x, P1)', lambda s,x: wymum(s^x, P1)),
                            ('mum(mum(x^P0? ', lambda s,x: wymum(s^P0, wymum(x^P1, P2)))):
      for fin_name, fin in (('mum(h,len^P5)', lambda h: wymum(h, 8^P5)), ('mum(h^len,P5)', lambda h: wymum(h^8, P5)), ('none', lambda h: h)):
        for red_name, red in (('fold-fastmod', lambda h,n: ((((h & 0xFFFFFFFF) ^ (h>>32)) * n) >> 32)), ('mod', lambda h,n: h % n), ('hi-fastmod', lambda h,n: ((h>>32)*n)>>32), ('lo-fastmod', lambda h,n: ((h&0xFFFFFFFF)*n)>>32)):
          ok = 0
          for i, k in enumerate(keys):
            got = None
            for lvl, (bits, w) in enumerate(bvs):
              h = red(fin(core(seedf(lvl), rd(k))), bits)
              if contains(bvs[lvl], h): got = rank_of(lvl, h); break
            if got == i: ok += 1
          if ok > 20: print(ok, len(keys), seedf_name, rd_name, core_name, fin_name, red_name)
print('done')
