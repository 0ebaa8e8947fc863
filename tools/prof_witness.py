"""Witness generation + satisfaction check only (for ncu launch lists): python tools/prof_witness.py [nsig] [logn]"""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from falcon_r1cs_b200 import api, synth, lib as L
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
logn = int(sys.argv[2]) if len(sys.argv) > 2 else 10
ctx = api.Context(logn); lib = L.load()
sig, pk, hm = synth.make_signatures(logn, n, seed=3)
dev = torch.device("cuda", 0)
d = [torch.from_numpy(x.view(np.int16)).to(dev) for x in (sig, pk, hm)]
z = torch.empty((n, ctx.n_z, 4), dtype=torch.int64, device=dev)
st = torch.zeros(n, dtype=torch.int32, device=dev); fu = torch.zeros(n, dtype=torch.int64, device=dev)
s = torch.cuda.current_stream().cuda_stream
for rep in range(3):
    L.check(lib.frcs_witness_batch_dev(ctx.h, n, *[C.c_void_p(t.data_ptr()) for t in d], C.c_void_p(z.data_ptr()), C.c_void_p(st.data_ptr()), C.c_void_p(s)), "w")
    L.check(lib.frcs_r1cs_eval_batch_dev(ctx.h, n, C.c_void_p(z.data_ptr()), None, None, None, C.c_void_p(fu.data_ptr()), C.c_void_p(s)), "e")
torch.cuda.synchronize()
assert os.environ.get('NOCHECK') or (int(st.abs().sum()) == 0 and int((fu != -1).sum()) == 0)
print("ok")
