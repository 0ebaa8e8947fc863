"""Satisfaction check alone (CUDA events, z resident): python tools/time_sat.py [nsig] [logn]; prints checks/s"""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from falcon_r1cs_b200 import api, synth, lib as L
n = int(sys.argv[1]) if len(sys.argv) > 1 else 592
logn = int(sys.argv[2]) if len(sys.argv) > 2 else 10
ctx = api.Context(logn); lib = L.load()
sig, pk, hm = synth.make_signatures(logn, min(n, 128), seed=3)
reps = (n + sig.shape[0] - 1) // sig.shape[0]
sig, pk, hm = [np.tile(x, (reps, 1))[:n] for x in (sig, pk, hm)]
dev = torch.device("cuda", 0)
d = [torch.from_numpy(np.ascontiguousarray(x).view(np.int16)).to(dev) for x in (sig, pk, hm)]
z = torch.empty((n, ctx.n_z, 4), dtype=torch.int64, device=dev)
st = torch.zeros(n, dtype=torch.int32, device=dev); fu = torch.zeros(n, dtype=torch.int64, device=dev)
s = torch.cuda.current_stream().cuda_stream
L.check(lib.frcs_witness_batch_dev(ctx.h, n, *[C.c_void_p(t.data_ptr()) for t in d], C.c_void_p(z.data_ptr()), C.c_void_p(st.data_ptr()), C.c_void_p(s)), "w")
def sat():
    L.check(lib.frcs_r1cs_eval_batch_dev(ctx.h, n, C.c_void_p(z.data_ptr()), None, None, None, C.c_void_p(fu.data_ptr()), C.c_void_p(s)), "e")
for _ in range(3): sat()
torch.cuda.synchronize()
assert int(st.abs().sum()) == 0 and int((fu != -1).sum()) == 0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 10
e0.record()
for _ in range(K): sat()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
print("%s n=%d logn=%d  check %.3f ms  %.0f checks/s  %.0f GB/s of z" % (os.environ.get("FRCS_LIB", "main").split("/")[-1], n, logn, ms, n / ms * 1e3, n * ctx.n_z * 32 / ms / 1e6))
