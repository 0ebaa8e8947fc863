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


P_MOD = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab


def fq_mont_limbs(x):
    m = x * (1 << 384) % P_MOD
    return [(m >> (64 * j)) & 0xFFFFFFFFFFFFFFFF for j in range(6)]


def test_point_validation_and_untrusted_proofs(circuits, oracle):
    """frcs_verify_proof must not feed unvalidated points into the pairing (arkworks validates at deserialisation):
    off-curve points, on-curve points outside the prime-order subgroup and out-of-range limbs are rejected"""
    lib = L.load()
    g1, g2 = gens(oracle)
    v1 = lambda p: lib.frcs_g1_validate(np.ascontiguousarray(p, np.uint64).ctypes.data_as(L.u64p))
    v2 = lambda p: lib.frcs_g2_validate(np.ascontiguousarray(p, np.uint64).ctypes.data_as(L.u64p))
    assert v1(g1) == 1 and v2(g2) == 1 and v1(np.zeros(12, np.uint64)) == 1 and v2(np.zeros(24, np.uint64)) == 1
    assert v1(g1_mul(oracle, g1, 12345)) == 1 and v2(g2_mul(oracle, g2, 99)) == 1
    off = g1.copy()
    off[0] ^= np.uint64(1)
    assert v1(off) == 0                                   # not on y^2 = x^3 + 4
    big = g1.copy()
    big[5] = np.uint64(0xFFFFFFFFFFFFFFFF)
    assert v1(big) == 0                                   # x >= p
    # a point on E(Fq) outside the order-r subgroup (the cofactor is ~2^126): first x with x^3 + 4 a square
    x = 1
    while pow((x ** 3 + 4) % P_MOD, (P_MOD - 1) // 2, P_MOD) != 1:
        x += 1
    y = pow((x ** 3 + 4) % P_MOD, (P_MOD + 1) // 4, P_MOD)  # p = 3 mod 4
    assert (y * y - x ** 3 - 4) % P_MOD == 0
    rogue = np.array(fq_mont_limbs(x) + fq_mont_limbs(y), dtype=np.uint64)
    assert v1(rogue) == 0
    off2 = g2.copy()
    off2[13] ^= np.uint64(4)
    assert v2(off2) == 0
    # through verify_proof: a valid proof with A replaced by such points is refused with an error, not evaluated
    c = circuits(9, 0)
    P = c.setup(seed=5152)
    g1e, g2e = P.export("g1_elems"), P.export("g2_elems")
    vk = {"alpha_g1": g1e[0], "beta_g2": g2e[0], "gamma_g2": g2e[2], "delta_g2": g2e[1],
          "gamma_abc_g1": P.export("gamma_abc_g1")}
    ic = np.ascontiguousarray(vk["gamma_abc_g1"], np.uint64)
    g2s = np.ascontiguousarray(np.stack([vk["beta_g2"], vk["gamma_g2"], vk["delta_g2"]]), np.uint64)
    assert lib.frcs_vk_validate(np.ascontiguousarray(vk["alpha_g1"]).ctypes.data_as(L.u64p), g2s.ctypes.data_as(L.u64p),
                                ic.ctypes.data_as(L.u64p), ic.shape[0] - 1) == 1
    sig, pk, hm = synth.make_signatures(9, 1, seed=82)
    z, st, _ = c.witness(sig[0], pk[0], hm[0])
    rng = np.random.default_rng(3)
    proof, _ = c.prove(P, z, api.fr_rand(rng), api.fr_rand(rng))
    assert api.verify_proof(vk, proof, z[1:c.n_inst])
    for bad_a in (rogue, off):
        forged = proof.copy()
        forged[:12] = bad_a
        with pytest.raises(L.FrcsError) as e:
            api.verify_proof(vk, forged, z[1:c.n_inst])
        assert e.value.code == L.E_INVALID_POINT
    forged = proof.copy()
    forged[12:36] = off2
    with pytest.raises(L.FrcsError):
        api.verify_proof(vk, forged, z[1:c.n_inst])
