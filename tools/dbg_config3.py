import ctypes as C, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from falcon_r1cs_b200 import api, synth, lib as L
ctx = api.Context(10); lib = L.load()
n3 = 65536
s3, p3, h3 = synth.make_signatures(10, 2048, seed=4321)
s3, p3, h3 = [np.ascontiguousarray(np.tile(x, (32, 1))[:n3]) for x in (s3, p3, h3)]
d = torch.empty(402653184 // 2, dtype=torch.int16, device="cuda")
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    d[: s3.size].copy_(torch.from_numpy(s3.view(np.int16)).reshape(-1)); torch.cuda.synchronize()
    print("pageable H2D 134 MB: %.3f s" % (time.perf_counter() - t0), flush=True)
for rep in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    fu, st = ctx.witness_check_batch(s3, p3, h3)
    torch.cuda.synchronize(); print("pageable api: %.3f s" % (time.perf_counter() - t0), flush=True)
ps, pp, ph = [torch.from_numpy(x.view(np.int16)).pin_memory() for x in (s3, p3, h3)]
fu = torch.zeros(n3, dtype=torch.int64).pin_memory(); st = torch.zeros(n3, dtype=torch.int32).pin_memory()
for rep in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    L.check(lib.frcs_witness_check_batch(ctx.h, n3, C.cast(ps.data_ptr(), L.u16p), C.cast(pp.data_ptr(), L.u16p), C.cast(ph.data_ptr(), L.u16p), C.cast(fu.data_ptr(), L.i64p), C.cast(st.data_ptr(), L.i32p)), "x")
    torch.cuda.synchronize(); print("pinned api: %.3f s" % (time.perf_counter() - t0), flush=True)
print(os.cpu_count(), open("/proc/loadavg").read())
