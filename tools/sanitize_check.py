"""compute-sanitizer target: one frcs_witness_check_batch of 128 Falcon-512 signatures + a gadget call + a dual-NTT batch"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from falcon_r1cs_b200 import api, synth, gadgets as G, lib as L
logn = int(sys.argv[1]) if len(sys.argv) > 1 else 9
ctx = api.Context(logn)
sig, pk, hm = synth.make_signatures(logn, 128, seed=9)
fu, st = ctx.witness_check_batch(sig, pk, hm)
assert (fu == -1).all() and (st == 0).all()
z, st = ctx.witness_batch(sig[:2], pk[:2], hm[:2])
az, bz, cz, fu = ctx.r1cs_eval_batch(z)
assert (fu == -1).all()
r = G.mod_q(ctx, [12290, 5], expected=[1, 5])
assert r.satisfied.all()
assert G.NTTPolyVar.ntt_circuit(ctx, sig[:2]).satisfied.all()
ctx.close()
d = api.Context(logn, kind=L.KIND_DUAL_NTT)
fu, st = d.witness_check_batch(sig[:70], pk[:70], hm[:70])
assert (fu == -1).all() and (st == 0).all()
d.close()
print("sanitize_check ok")
