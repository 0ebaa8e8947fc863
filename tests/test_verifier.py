"""The host-side Groth16 verifier of the product (frcs_verify_proof = ark_groth16::verify_proof,
examples/pok_sig.rs:45-47): pairing identities, then verification of oracle-made proofs.  No GPU needed."""
import numpy as np
import pytest

from falcon_r1cs_b200 import api, synth
from falcon_r1cs_b200 import lib as L

R = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001


def gens(oracle):
    g1 = np.zeros(12, dtype=np.uint64)
    g2 = np.zeros(24, dtype=np.uint64)
    oracle.lib().orc_generators(oracle.ptr(g1), oracle.ptr(g2))
    return g1, g2


def g1_mul(oracle, p, k):
    out = np.zeros(12, dtype=np.uint64)
    oracle.lib().orc_g1_mul(oracle.ptr(p), oracle.ptr(oracle.ints_to_limbs([k % R])[0]), oracle.ptr(out))
    return out


def g2_mul(oracle, p, k):
    out = np.zeros(24, dtype=np.uint64)
    oracle.lib().orc_g2_mul(oracle.ptr(p), oracle.ptr(oracle.ints_to_limbs([k % R])[0]), oracle.ptr(out))
    return out


def pairing_eq(p1, q1, p2, q2):
    return L.load().frcs_pairing_eq(*[x.ctypes.data_as(L.u64p) for x in (p1, q1, p2, q2)])


def test_pairing_is_bilinear_and_non_degenerate(oracle):
    g1, g2 = gens(oracle)
    assert L.load().frcs_pairing_is_one(g1.ctypes.data_as(L.u64p), g2.ctypes.data_as(L.u64p)) == 0
    a, b = 0x1234567890abcdef1234567, 0xfedcba9876543210fedcba98765
    assert pairing_eq(g1_mul(oracle, g1, a), g2_mul(oracle, g2, b), g1_mul(oracle, g1, a * b), g2) == 1
    assert pairing_eq(g1_mul(oracle, g1, a), g2, g1, g2_mul(oracle, g2, a)) == 1
    assert pairing_eq(g1_mul(oracle, g1, a), g2, g1, g2_mul(oracle, g2, a + 1)) == 0
    # e(P, Q)^r = 1: multiplying by the group order gives the point at infinity -> pairing one
    inf = np.zeros(12, dtype=np.uint64)
    assert L.load().frcs_pairing_is_one(inf.ctypes.data_as(L.u64p), g2.ctypes.data_as(L.u64p)) == 1


def test_verify_oracle_proof_and_reject_tampering(circuits, oracle):
    c = circuits(9, 0)
    P = c.setup(seed=5151)
    g1e, g2e = P.export("g1_elems"), P.export("g2_elems")  # alpha, beta, delta | beta, delta, gamma
    vk = {"alpha_g1": g1e[0], "beta_g2": g2e[0], "gamma_g2": g2e[2], "delta_g2": g2e[1],
          "gamma_abc_g1": P.export("gamma_abc_g1")}
    sig, pk, hm = synth.make_signatures(9, 1, seed=81)
    z, st, _ = c.witness(sig[0], pk[0], hm[0])
    rng = np.random.default_rng(2)
    r, s = api.fr_rand(rng), api.fr_rand(rng)
    proof, _ = c.prove(P, z, r, s)
    public = z[1:c.n_inst]
    assert api.verify_proof(vk, proof, public)
    bad = public.copy()
    bad[3] = public[4]
    assert not api.verify_proof(vk, proof, bad)                       # wrong statement
    other, _ = c.prove(P, z, s, r)
    mixed = proof.copy()
    mixed[36:] = other[36:]
    assert not api.verify_proof(vk, mixed, public)                    # C from another proof
    assert api.verify_proof(vk, other, public)                        # re-randomised proof still verifies
    with pytest.raises(ValueError):
        api.verify_proof(vk, proof, public[:-1])
