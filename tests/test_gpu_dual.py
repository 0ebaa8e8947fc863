"""GPU parity for FalconDualNTTVerificationCircuit (circuits/falcon_dual_ntt.rs:26-132, gadgets/dual_poly.rs:8-52;
SURVEY.md section 8f.3): matrices, witness generation, R1CS evaluation and a Groth16 proof through the C ABI,
bit-exact against the oracle."""
import numpy as np
import pytest

from falcon_r1cs_b200 import api, synth
from falcon_r1cs_b200 import lib as L

pytestmark = pytest.mark.gpu
Q = 12289
_CTX = {}


def dual_ctx(logn):
    if logn not in _CTX:
        _CTX[logn] = api.Context(logn, kind=L.KIND_DUAL_NTT)
    return _CTX[logn]


@pytest.mark.parametrize("logn", [9, 10])
def test_dual_matrices_equal_oracle(circuits, logn):
    ctx, c = dual_ctx(logn), circuits(logn, 2)
    assert (ctx.n_inst, ctx.n_wit, ctx.n_cons, ctx.domain_log2) == (c.n_inst, c.n_wit, c.n_cons, c.domain_log2)
    assert ctx.nnz == (c.nnz_a, c.nnz_b, c.nnz_c)
    for which in range(3):
        rp, col, val = ctx.get_matrix(which)
        orp, ocol, oval = c.csr(which)
        assert (rp == orp).all() and (col == ocol).all() and (val == oval).all()


@pytest.mark.parametrize("logn", [9, 10])
def test_dual_witness_bit_exact(circuits, logn):
    ctx, c = dual_ctx(logn), circuits(logn, 2)
    n = 1 << logn
    sig, pk, hm = synth.make_signatures(logn, 20, seed=81)
    rng = np.random.default_rng(6)
    extra = [(np.zeros(n), np.zeros(n), np.zeros(n)), (np.full(n, Q - 1), np.full(n, Q - 1), np.full(n, Q - 1)),
             (rng.integers(0, Q, n), rng.integers(0, Q, n), rng.integers(0, Q, n)),
             (np.full(n, 6143), rng.integers(0, Q, n), np.full(n, 6144))]
    sig = np.concatenate([sig, np.array([e[0] for e in extra], np.uint16)])
    pk = np.concatenate([pk, np.array([e[1] for e in extra], np.uint16)])
    hm = np.concatenate([hm, np.array([e[2] for e in extra], np.uint16)])
    z, st = ctx.witness_batch(sig, pk, hm)
    assert (st[:20] == 0).all()
    for i in range(sig.shape[0]):
        zo, sto, _ = c.witness(sig[i], pk[i], hm[i])
        assert st[i] == {0: 0, -1: -16, -2: -17}[sto], i
        bad = np.nonzero((z[i] != zo).any(axis=1))[0]
        assert bad.size == 0, (i, bad[:10])


@pytest.mark.parametrize("logn,n", [(9, 3), (10, 70)])
def test_dual_r1cs_eval_bit_exact(circuits, logn, n):
    """small batch (warp-per-row long rows) and a batch >= 64 (bundled long rows)"""
    ctx, c = dual_ctx(logn), circuits(logn, 2)
    sig, pk, hm = synth.make_signatures(logn, n, seed=82)
    z, st = ctx.witness_batch(sig, pk, hm)
    assert (st == 0).all()
    z[1, c.n_inst + 7] = z[1, 0]               # sig.pos[7] := 1
    z[n - 1, c.n_z - 5] = z[n - 1, 0] if not z[n - 1, c.n_z - 5].any() else np.zeros(4, np.uint64)  # a norm-chain bit
    az, bz, cz, fu = ctx.r1cs_eval_batch(z)
    for i in sorted({0, 1, 2, n // 2, n - 2, n - 1}):
        oa, ob, oc, ofu = c.r1cs_eval(z[i])
        assert (az[i] == oa).all() and (bz[i] == ob).all() and (cz[i] == oc).all(), i
        assert fu[i] == ofu, i
    assert fu[0] == -1 and fu[1] >= 0 and fu[n - 1] >= 0
    good = np.ones(n, bool)
    good[[1, n - 1]] = False
    assert (fu[good] == -1).all()
    fu2, st2 = ctx.witness_check_batch(sig, pk, hm)
    assert (fu2 == -1).all() and (st2 == 0).all()


def test_dual_proof_byte_identical(circuits, oracle):
    ctx, c = dual_ctx(9), circuits(9, 2)
    P = c.setup(seed=4001)
    g1, g2 = P.export("g1_elems"), P.export("g2_elems")
    ctx.load_pk(api.ProvingKey(alpha_g1=g1[0], beta_g1=g1[1], delta_g1=g1[2], beta_g2=g2[0], delta_g2=g2[1],
                               a_query=P.export("a_query"), b_g1_query=P.export("b_g1_query"),
                               b_g2_query=P.export("b_g2_query"), h_query=P.export("h_query"),
                               l_query=P.export("l_query")))
    sig, pk, hm = synth.make_signatures(9, 2, seed=83)
    rng = np.random.default_rng(9)
    r = np.stack([api.fr_rand(rng) for _ in range(2)])
    s = np.stack([api.fr_rand(rng) for _ in range(2)])
    proofs, st = ctx.prove_batch(sig, pk, hm, r, s)
    assert (st == 0).all()
    for i in range(2):
        z, _, _ = c.witness(sig[i], pk[i], hm[i])
        want, want_bytes = c.prove(P, z, r[i], s[i])
        assert (proofs[i] == want).all(), i
        assert api.proof_compress(proofs[i]) == bytes(want_bytes)
        assert c.verify_trapdoor(P, z, r[i], s[i], proofs[i])
