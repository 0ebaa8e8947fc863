"""GPU parity for the whole path: create_proof(circuit, pk, r, s) through the C ABI
(frcs_load_pk + frcs_prove_batch / frcs_prove_from_z) vs the oracle's restatement of
ark-groth16's prover: proofs byte-identical under the same (r, s), and valid."""
import numpy as np
import pytest

from falcon_r1cs_b200 import api, synth

pytestmark = pytest.mark.gpu
_STATE = {}


def setup_for(logn, contexts, circuits):
    if logn not in _STATE:
        ctx, c = contexts(logn), circuits(logn, 0)
        P = c.setup(seed=1000 + logn)
        g1, g2 = P.export("g1_elems"), P.export("g2_elems")
        pk = api.ProvingKey(alpha_g1=g1[0], beta_g1=g1[1], delta_g1=g1[2], beta_g2=g2[0], delta_g2=g2[1],
                            a_query=P.export("a_query"), b_g1_query=P.export("b_g1_query"),
                            b_g2_query=P.export("b_g2_query"), h_query=P.export("h_query"),
                            l_query=P.export("l_query"))
        ctx.load_pk(pk)
        _STATE[logn] = (ctx, c, P)
    return _STATE[logn]


def fr(oracle, v):
    return oracle.fr_from_canonical(oracle.ints_to_limbs([v]))[0]


@pytest.mark.parametrize("logn", [9, 10])
def test_proof_byte_identical_and_valid(contexts, circuits, oracle, logn):
    ctx, c, P = setup_for(logn, contexts, circuits)
    n = 3
    sig, pk, hm = synth.make_signatures(logn, n, seed=51)
    rng = np.random.default_rng(3)
    r = np.stack([api.fr_rand(rng) for _ in range(n)])
    s = np.stack([api.fr_rand(rng) for _ in range(n)])
    proofs, st = ctx.prove_batch(sig, pk, hm, r, s)
    assert (st == 0).all()
    for i in range(n):
        z, sto, _ = c.witness(sig[i], pk[i], hm[i])
        want, want_bytes = c.prove(P, z, r[i], s[i])
        assert (proofs[i] == want).all(), i
        assert api.proof_compress(proofs[i]) == bytes(want_bytes)
        assert c.verify_trapdoor(P, z, r[i], s[i], proofs[i])
    # the same through frcs_prove_from_z
    z, _ = ctx.witness_batch(sig[:1], pk[:1], hm[:1])
    assert (ctx.prove_from_z(z, r[:1], s[:1])[0] == proofs[0]).all()


def test_proof_edge_randomizers(contexts, circuits, oracle):
    """r = 0 (ark-groth16 skips g1_b), s = 0, r = s = 1"""
    ctx, c, P = setup_for(9, contexts, circuits)
    sig, pk, hm = synth.make_signatures(9, 1, seed=52)
    z, _, _ = c.witness(sig[0], pk[0], hm[0])
    for rv, sv in ((0, 5), (7, 0), (0, 0), (1, 1)):
        r, s = fr(oracle, rv), fr(oracle, sv)
        got = ctx.prove_from_z(z, r, s)[0]
        want, _ = c.prove(P, z, r, s)
        assert (got == want).all(), (rv, sv)


def test_create_random_proof_call_shape(contexts, circuits, oracle):
    """examples/pok_sig.rs:24-47: build_circuit -> create_random_proof -> verify"""
    ctx, c, P = setup_for(9, contexts, circuits)
    sig, pk, hm = synth.make_signatures(9, 1, seed=53)
    circuit = api.FalconNTTVerificationCircuit.build_circuit(pk[0], b"testing message", sig[0], hm=hm[0])
    rng = np.random.default_rng(0)
    proof = api.create_random_proof(ctx, circuit, rng)
    rng2 = np.random.default_rng(0)
    r, s = api.fr_rand(rng2), api.fr_rand(rng2)
    z, _, _ = c.witness(sig[0], pk[0], hm[0])
    assert c.verify_trapdoor(P, z, r, s, proof)
    # an invalid signature: the reference panics in witness generation; we report the status
    bad = sig[0].copy()
    bad[:] = 6000
    with pytest.raises(ValueError):
        api.create_random_proof(ctx, api.FalconNTTVerificationCircuit.build_circuit(pk[0], b"m", bad, hm=hm[0]), rng)


def test_gpu_proofs_pass_the_pairing_verifier(contexts, circuits, oracle):
    """examples/pok_sig.rs:45-47: verify_proof(&pvk, &proof, &public_inputs) on GPU-made proofs"""
    ctx, c, P = setup_for(9, contexts, circuits)
    g1e, g2e = P.export("g1_elems"), P.export("g2_elems")
    vk = {"alpha_g1": g1e[0], "beta_g2": g2e[0], "gamma_g2": g2e[2], "delta_g2": g2e[1],
          "gamma_abc_g1": P.export("gamma_abc_g1")}
    sig, pk, hm = synth.make_signatures(9, 2, seed=54)
    rng = np.random.default_rng(11)
    r = np.stack([api.fr_rand(rng) for _ in range(2)])
    s = np.stack([api.fr_rand(rng) for _ in range(2)])
    proofs, st = ctx.prove_batch(sig, pk, hm, r, s)
    z, _ = ctx.witness_batch(sig, pk, hm)
    for i in range(2):
        assert api.verify_proof(vk, proofs[i], z[i, 1:ctx.n_inst])
    assert not api.verify_proof(vk, proofs[0], z[1, 1:ctx.n_inst])  # another signature's statement


def test_prove_without_pk_fails(circuits):
    ctx = api.Context(9)
    try:
        z = np.zeros((ctx.n_z, 4), dtype=np.uint64)
        with pytest.raises(Exception) as e:
            ctx.prove_from_z(z, np.zeros(4, dtype=np.uint64), np.zeros(4, dtype=np.uint64))
        assert "proving key" in str(e.value)
    finally:
        ctx.close()


def test_proof_batch_many_groups(contexts, circuits, oracle, monkeypatch):
    """A batch that spans several proof groups (groups of 16: 16 + 16 + 8): every group after the second reuses the
    double-buffered sort buffers of the group two before it (z sort, B sort without the infinity bases, l+h sort), so
    the stream / event ordering of prove.cu is on the path.  Every proof byte-identical to the oracle's."""
    from concurrent.futures import ThreadPoolExecutor
    ctx, c, P = setup_for(9, contexts, circuits)
    n = 40
    sig, pk, hm = synth.make_signatures(9, n, seed=55)
    rng = np.random.default_rng(5)
    r = np.stack([api.fr_rand(rng) for _ in range(n)])
    s = np.stack([api.fr_rand(rng) for _ in range(n)])
    proofs, st = ctx.prove_batch(sig, pk, hm, r, s)
    assert (st == 0).all()

    def same(i):
        z, _, _ = c.witness(sig[i], pk[i], hm[i])
        want, _ = c.prove(P, z, r[i], s[i])
        return bool((proofs[i] == want).all())
    with ThreadPoolExecutor(max_workers=4) as ex:
        ok = list(ex.map(same, range(n)))
    assert all(ok), [i for i, o in enumerate(ok) if not o]


def test_proof_parity_without_b_skip_and_small_groups():
    """The fall-back the B MSMs take when a key has few infinity bases (one shared sort of z for a_query, b_g1 and b_g2:
    FRCS_NO_BSKIP=1) and groups of 2 (FRCS_GROUP=2: three groups for three proofs), in a process of their own because
    both switches are read once: the byte-identity test must pass there too."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, FRCS_NO_BSKIP="1", FRCS_GROUP="2")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu",
                          os.path.join(root, "tests", "test_gpu_prove.py") + "::test_proof_byte_identical_and_valid"],
                         env=env, cwd=root, capture_output=True, text=True, timeout=1200)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "2 passed" in out.stdout, out.stdout[-500:]
