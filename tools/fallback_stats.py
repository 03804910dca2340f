import sys, numpy as np
sys.path.insert(0,'/root/repo')
from flemingo_b200 import synthetic
from flemingo_b200.engine import Engine
eng=Engine(0)
orgs=synthetic.mixed_organisms(400,300,3)
seqs=synthetic.random_sequences(500,300,1)
# per composition class
groups={}
for o in orgs:
    n=len(o.recognizer_types); np_=o.recognizer_types.count('p')
    key=(n, np_)
    groups.setdefault(key,[]).append(o)
for key in sorted(groups):
    g=groups[key]
    if key[0]<2: continue
    eng.upload_population(g); ss=eng.upload_sequences(seqs,0); eng.place(ss,fetch=False); eng.sync()
    a,b=eng.path_stats()
    print(key, len(g), 'fast',a,'fallback',b, f'{b/max(a,1):.3f}')
