"""GPU parity for Groth16 parameter generation (Groth16::circuit_specific_setup, examples/pok_sig.rs:30-31):
frcs_setup from the oracle's toxic waste must reproduce the oracle's proving and verifying key point for
point, and proofs made with the device-resident key must equal the oracle's."""
import numpy as np
import pytest

from falcon_r1cs_b200 import api, synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("logn", [9, 10])
def test_setup_equals_oracle(circuits, oracle, logn):
    c = circuits(logn, 0)
    P = c.setup(seed=4000 + logn)
    ctx = api.Context(logn)
    try:
        vk = ctx.setup(P.trapdoor())
        for name in ("a_query", "b_g1_query", "b_g2_query", "h_query", "l_query"):
            got, want = ctx.export_pk(name), P.export(name)
            assert got.shape == want.shape, name
            bad = np.nonzero((got != want).any(axis=1))[0]
            assert bad.size == 0, (name, bad[:5])
        g1, g2 = P.export("g1_elems"), P.export("g2_elems")  # alpha, beta, delta | beta, delta, gamma
        assert (vk["alpha_g1"] == g1[0]).all()
        assert (vk["beta_g2"] == g2[0]).all() and (vk["delta_g2"] == g2[1]).all() and (vk["gamma_g2"] == g2[2]).all()
        assert (vk["gamma_abc_g1"] == P.export("gamma_abc_g1")).all()
        # prove with the key that never left the device
        sig, pk, hm = synth.make_signatures(logn, 2, seed=71)
        rng = np.random.default_rng(5)
        r = np.stack([api.fr_rand(rng) for _ in range(2)])
        s = np.stack([api.fr_rand(rng) for _ in range(2)])
        proofs, st = ctx.prove_batch(sig, pk, hm, r, s)
        assert (st == 0).all()
        for i in range(2):
            z, _, _ = c.witness(sig[i], pk[i], hm[i])
            want, _ = c.prove(P, z, r[i], s[i])
            assert (proofs[i] == want).all()
    finally:
        ctx.close()


@pytest.mark.parametrize("seed", [7, 1, 12345])
def test_setup_random_trapdoor_proofs_verify(circuits, oracle, seed):
    """a fresh trapdoor drawn on the host: key generated and kept on the device, a batch of proofs, each must pass
    the pairing verifier with the verifying key frcs_setup returned (regression: the table construction once
    raced with the device-to-device copies of the queries)"""
    ctx = api.Context(9)
    try:
        vk = ctx.setup(api.random_trapdoor(np.random.default_rng(seed)))
        n = 5
        sig, pk, hm = synth.make_signatures(9, n, seed=1234)
        rng = np.random.default_rng(99)
        r = np.stack([api.fr_rand(rng) for _ in range(n)])
        s = np.stack([api.fr_rand(rng) for _ in range(n)])
        proofs, st = ctx.prove_batch(sig, pk, hm, r, s)
        z, _ = ctx.witness_batch(sig, pk, hm)
        assert (st == 0).all()
        for i in range(n):
            assert api.verify_proof(vk, proofs[i], z[i, 1:ctx.n_inst]), i
    finally:
        ctx.close()
