import ctypes as C, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from falcon_r1cs_b200 import api, synth, lib as L
ctx = api.Context(10); lib = L.load()
n3 = 16384
s3, p3, h3 = synth.make_signatures(10, 2048, seed=4321)
s3, p3, h3 = [np.ascontiguousarray(np.tile(x, (8, 1))[:n3]) for x in (s3, p3, h3)]
def run(tag):
    ctx.witness_check_batch(s3, p3, h3)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    fu, st = ctx.witness_check_batch(s3, p3, h3)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(tag, "%.0f wit/s" % (n3 / dt), flush=True)
run("fresh")
vk = ctx.setup(api.random_trapdoor(np.random.default_rng(7)))
run("after setup")
sig, pk, hm = synth.make_signatures(10, 64, seed=1)
rng = np.random.default_rng(1)
r = np.stack([api.fr_rand(rng) for _ in range(64)]); s = np.stack([api.fr_rand(rng) for _ in range(64)])
ctx.prove_batch(sig, pk, hm, r, s)
run("after prove")
ctx.profile_enable(True); ctx.prove_batch(sig, pk, hm, r, s); [ctx.profile_get(k) for k in ctx.PROF]; ctx.profile_enable(False)
run("after profiled prove")
x = torch.empty((592, ctx.n_z, 4), dtype=torch.int64, device="cuda"); del x
run("after torch alloc/free")
ctx.imad_peak()
run("after imad_peak")
