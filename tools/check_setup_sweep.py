"""Regression sweep: Falcon-512 proofs with the key generated on the device (frcs_setup) at several batch sizes, every
proof compared with the oracle; on a mismatch the proving-key queries are compared too.  (This caught a race between
legacy-stream uploads from pageable memory and kernels on the context's non-blocking stream.)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, oracle_lib as O
from falcon_r1cs_b200 import api, synth
logn = 9
c = O.Circuit(logn, 0)
td = api.random_trapdoor(np.random.default_rng(7))
P = c.setup(trapdoor=td)
g1, g2 = P.export("g1_elems"), P.export("g2_elems")
vk = {"alpha_g1": g1[0], "beta_g2": g2[0], "gamma_g2": g2[2], "delta_g2": g2[1], "gamma_abc_g1": P.export("gamma_abc_g1")}
PK = api.ProvingKey(alpha_g1=g1[0], beta_g1=g1[1], delta_g1=g1[2], beta_g2=g2[0], delta_g2=g2[1], a_query=P.export("a_query"), b_g1_query=P.export("b_g1_query"), b_g2_query=P.export("b_g2_query"), h_query=P.export("h_query"), l_query=P.export("l_query"))
N = 40
sig, pk, hm = synth.make_signatures(logn, N, seed=1234)
rng = np.random.default_rng(99)
r = np.stack([api.fr_rand(rng) for _ in range(N)]); s = np.stack([api.fr_rand(rng) for _ in range(N)])
want = []
for i in range(N):
    z, _, _ = c.witness(sig[i], pk[i], hm[i])
    want.append(c.prove(P, z, r[i], s[i])[0])
want = np.stack(want)
for mode in ("setup",):
    for n in (15, 17, 15, 17, 5, 15, 31, 33, 2, 15):
        ctx = api.Context(logn)
        if mode == "load":
            ctx.load_pk(PK)
        else:
            ctx.setup(td)
        got, st = ctx.prove_batch(sig[:n], pk[:n], hm[:n], r[:n], s[:n])
        eq = (got == want[:n]).all(axis=1)
        print(mode, n, "all equal" if eq.all() else "MISMATCH at %s" % np.nonzero(~eq)[0][:10].tolist(), flush=True)
        if not eq.all():
            i = int(np.nonzero(~eq)[0][0])
            print("   parts A,B,C equal:", (got[i][:12] == want[i][:12]).all(), (got[i][12:36] == want[i][12:36]).all(), (got[i][36:] == want[i][36:]).all())
            for name in ("a_query", "b_g1_query", "b_g2_query", "h_query", "l_query"):
                print("   ", name, "mismatching points:", int((ctx.export_pk(name) != P.export(name)).any(axis=1).sum()))
            got2, _ = ctx.prove_batch(sig[:n], pk[:n], hm[:n], r[:n], s[:n])
            print("   second call equal to oracle:", (got2 == want[:n]).all(), " equal to first:", (got2 == got).all())
        ctx.close()
