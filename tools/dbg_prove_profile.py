import sys, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
import oracle_lib as O
from falcon_r1cs_b200 import api, synth
logn=int(sys.argv[1]) if len(sys.argv)>1 else 10
ctx=api.Context(logn); c=O.Circuit(logn,0)
t=time.time(); P=c.setup(7); print('oracle setup s',time.time()-t, 'threads', O.lib().orc_num_threads())
g1,g2=P.export("g1_elems"),P.export("g2_elems")
pk=api.ProvingKey(alpha_g1=g1[0],beta_g1=g1[1],delta_g1=g1[2],beta_g2=g2[0],delta_g2=g2[1],a_query=P.export("a_query"),b_g1_query=P.export("b_g1_query"),b_g2_query=P.export("b_g2_query"),h_query=P.export("h_query"),l_query=P.export("l_query"))
t=time.time(); ctx.load_pk(pk); print('load_pk s',time.time()-t)
n=8
sig,pkk,hm=synth.make_signatures(logn,n,seed=1)
rng=np.random.default_rng(1)
r=np.stack([api.fr_rand(rng) for _ in range(n)]); s=np.stack([api.fr_rand(rng) for _ in range(n)])
ctx.prove_batch(sig,pkk,hm,r,s)
ctx.profile_enable(True)
for rep in range(3):
    t=time.time(); proofs,st=ctx.prove_batch(sig,pkk,hm,r,s); dt=time.time()-t
    print('prove_batch n=%d: %.1f ms/proof'%(n,dt*1e3/n))
for k in ctx.PROF:
    ms,cnt,work=ctx.profile_get(k)
    if cnt: print('%-12s %8.3f ms/launch  (%d launches) work=%d'%(k,ms/cnt,cnt,work))
print('imad peak LP/s %.3e'%ctx.imad_peak())
t=time.time(); z,_=ctx.witness_batch(sig,pkk,hm); print('witness_batch host path', (time.time()-t)*1e3/n,'ms/sig')
t=time.time(); zo,_,_=c.witness(sig[0],pkk[0],hm[0],construct_matrices=True); print('oracle witness s',time.time()-t)
t=time.time(); c.prove(P,zo,r[0],s[0]); print('oracle prove s',time.time()-t)
