"""Satisfaction-check timing (CUDA events, device-resident z): python tools/time_r1cs.py [nsig] [logn] [reps]"""
import ctypes as C, sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from falcon_r1cs_b200 import api, synth, lib as L
n = int(sys.argv[1]) if len(sys.argv) > 1 else 592
logn = int(sys.argv[2]) if len(sys.argv) > 2 else 10
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
ctx = api.Context(logn); lib = L.load()
sig, pk, hm = synth.make_signatures(logn, n, seed=3)
dev = torch.device("cuda", 0)
d = [torch.from_numpy(x.view(np.int16)).to(dev) for x in (sig, pk, hm)]
z = torch.empty((n, ctx.n_z, 4), dtype=torch.int64, device=dev)
st = torch.zeros(n, dtype=torch.int32, device=dev); fu = torch.zeros(n, dtype=torch.int64, device=dev)
s = torch.cuda.current_stream().cuda_stream
L.check(lib.frcs_witness_batch_dev(ctx.h, n, *[C.c_void_p(t.data_ptr()) for t in d], C.c_void_p(z.data_ptr()), C.c_void_p(st.data_ptr()), C.c_void_p(s)), "w")
def run():
    L.check(lib.frcs_r1cs_eval_batch_dev(ctx.h, n, C.c_void_p(z.data_ptr()), None, None, None, C.c_void_p(fu.data_ptr()), C.c_void_p(s)), "e")
for _ in range(3): run()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(reps): run()
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / reps
assert int(st.abs().sum()) == 0 and int((fu != -1).sum()) == 0
print(json.dumps({"sub": os.environ.get("FRCS_R1CS_SUB", "default"), "nsig": n, "logn": logn, "ms": round(ms, 4), "checks_per_s": round(n / ms * 1e3)}))
