"""GPU parity for FalconSchoolBookVerificationCircuit (circuits/falcon_schoolbook.rs:26-132; BASELINE
configs[3]): matrices, witness generation, R1CS evaluation and one Groth16 proof through the C ABI,
bit-exact against the oracle."""
import numpy as np
import pytest

from falcon_r1cs_b200 import api, synth
from falcon_r1cs_b200 import lib as L

pytestmark = pytest.mark.gpu
Q = 12289
_CTX = {}


def sb_ctx(logn):
    if logn not in _CTX:
        _CTX[logn] = api.Context(logn, kind=L.KIND_SCHOOLBOOK)
    return _CTX[logn]


@pytest.mark.parametrize("logn", [9, 10])
def test_schoolbook_matrices_equal_oracle(circuits, logn):
    ctx, c = sb_ctx(logn), circuits(logn, 1)
    assert (ctx.n_inst, ctx.n_wit, ctx.n_cons, ctx.domain_log2) == (c.n_inst, c.n_wit, c.n_cons, c.domain_log2)
    assert ctx.nnz == (c.nnz_a, c.nnz_b, c.nnz_c)
    for which in range(3):
        rp, col, val = ctx.get_matrix(which)
        orp, ocol, oval = c.csr(which)
        assert (rp == orp).all() and (col == ocol).all() and (val == oval).all()


@pytest.mark.parametrize("logn", [9, 10])
def test_schoolbook_witness_and_eval_bit_exact(circuits, logn):
    ctx, c = sb_ctx(logn), circuits(logn, 1)
    n = 3
    sig, pk, hm = synth.make_signatures(logn, n, seed=31)
    # edge: a public key with zero coefficients (neg_pk = q - 0 = q, not 0); hm is kept, so v changes and the
    # norm bound may fail: the status and the full z must still agree with the oracle
    pk[1, ::3] = 0
    z, st = ctx.witness_batch(sig, pk, hm)
    for i in range(n):
        zo, sto, _ = c.witness(sig[i], pk[i], hm[i])
        assert st[i] == {0: 0, -1: -16, -2: -17}[sto]
        bad = np.nonzero((z[i] != zo).any(axis=1))[0]
        assert bad.size == 0, (i, bad[:10])
    z[2, c.n_inst + 11] = z[2, c.n_inst + 12] + np.uint64(3)  # corrupt: first violated row must match
    az, bz, cz, fu = ctx.r1cs_eval_batch(z)
    for i in range(n):
        oa, ob, oc, ofu = c.r1cs_eval(z[i])
        assert (az[i] == oa).all() and (bz[i] == ob).all() and (cz[i] == oc).all()
        assert fu[i] == ofu
    assert fu[0] == -1 and fu[2] >= 0


def test_schoolbook_witness_statuses(circuits):
    ctx, c = sb_ctx(9), circuits(9, 1)
    n = 512
    rng = np.random.default_rng(6)
    sig = np.stack([np.zeros(n), rng.integers(0, Q, n), np.full(n, Q - 1)]).astype(np.uint16)
    pk = np.stack([np.zeros(n), rng.integers(0, Q, n), np.full(n, Q - 1)]).astype(np.uint16)
    hm = np.stack([np.zeros(n), rng.integers(0, Q, n), np.full(n, Q - 1)]).astype(np.uint16)
    z, st = ctx.witness_batch(sig, pk, hm)
    for i in range(3):
        zo, sto, _ = c.witness(sig[i], pk[i], hm[i], panic_on_range=True)
        assert (z[i] == zo).all(), i
        assert st[i] == {0: 0, -1: -16, -2: -17}[sto]


def test_schoolbook_proof_byte_identical(circuits, oracle):
    """Falcon-512 schoolbook circuit (315,956 constraints, domain 2^19): create_proof under a fixed (r, s)"""
    ctx, c = sb_ctx(9), circuits(9, 1)
    P = c.setup(seed=3001)
    g1, g2 = P.export("g1_elems"), P.export("g2_elems")
    ctx.load_pk(api.ProvingKey(alpha_g1=g1[0], beta_g1=g1[1], delta_g1=g1[2], beta_g2=g2[0], delta_g2=g2[1],
                               a_query=P.export("a_query"), b_g1_query=P.export("b_g1_query"),
                               b_g2_query=P.export("b_g2_query"), h_query=P.export("h_query"),
                               l_query=P.export("l_query")))
    sig, pk, hm = synth.make_signatures(9, 2, seed=32)
    rng = np.random.default_rng(8)
    r = np.stack([api.fr_rand(rng) for _ in range(2)])
    s = np.stack([api.fr_rand(rng) for _ in range(2)])
    proofs, st = ctx.prove_batch(sig, pk, hm, r, s)
    assert (st == 0).all()
    z, _, _ = c.witness(sig[0], pk[0], hm[0])
    h = ctx.witness_map(z)
    assert (h == c.witness_map(z)).all()
    want, want_bytes = c.prove(P, z, r[0], s[0])
    assert (proofs[0] == want).all()
    assert api.proof_compress(proofs[0]) == bytes(want_bytes)
    z1, _, _ = c.witness(sig[1], pk[1], hm[1])
    assert c.verify_trapdoor(P, z1, r[1], s[1], proofs[1])
