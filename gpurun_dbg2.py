import sys, ctypes as C
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
import oracle_lib as O
from falcon_r1cs_b200 import api, lib as L
ctx=api.Context(9)
g1=np.zeros(12,dtype=np.uint64); g2=np.zeros(24,dtype=np.uint64)
O.lib().orc_generators(O.ptr(g1),O.ptr(g2))
def mul(p,k,g2f):
    w=24 if g2f else 12
    out=np.zeros(w,dtype=np.uint64); kk=O.ints_to_limbs([k])[0]
    (O.lib().orc_g2_mul if g2f else O.lib().orc_g1_mul)(O.ptr(p),O.ptr(kk),O.ptr(out)); return out
lib=L.load()
for g2f,gen,w,fn in ((0,g1,12,lib.frcs_debug_windows_g1),(1,g2,24,lib.frcs_debug_windows_g2)):
    for n in (1,3):
        bases=np.stack([mul(gen,1000+j,g2f) for j in range(n)])
        out=np.zeros((16,n,w),dtype=np.uint64)
        fn.argtypes=[C.c_void_p,C.c_uint64,L.u64p,L.u64p]
        rc=fn(ctx.h,n,bases.ctypes.data_as(L.u64p),out.ctypes.data_as(L.u64p)); assert rc==0
        ok=[[bool((out[k,j]==mul(bases[j],1<<(16*k),g2f)).all()) for k in range(16)] for j in range(n)]
        print('g2' if g2f else 'g1', n, ok)
        if g2f and n==1:
            for k in (1,2):
                e=mul(bases[0],1<<(16*k),1)
                print(k,'x.c0',(out[k,0][:6]==e[:6]).all(),'x.c1',(out[k,0][6:12]==e[6:12]).all(),'y.c0',(out[k,0][12:18]==e[12:18]).all(),'y.c1',(out[k,0][18:]==e[18:]).all())
