"""CPU tests: the oracle against every pin the reference holds for the hot path
(SURVEY.md §8c): README constraint counts, gadget known answers, NTT_TABLE, NTT
gadget values == clear-text NTT, satisfiability, proof validity."""
import json
import os
import random

import numpy as np
import pytest

from falcon_r1cs_b200 import synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
Q = 12289


def gold(name):
    return json.load(open(os.path.join(GOLD, name)))


def test_ntt_table_matches_sage_script():
    tab = gold("ntt_table.json")
    assert len(tab) == 1024 and tab[:4] == [1, 10810, 7143, 4043]
    assert list(synth.ntt_table(1024)) == tab
    # Falcon-512 uses the first 512 entries (gadgets/misc.rs:72)
    assert list(synth.ntt_table(512)) == tab[:512]


@pytest.mark.parametrize("logn", [9, 10])
def test_readme_counts_ntt_circuit(circuits, logn):
    c = circuits(logn, 0)
    want = gold("readme_counts.json")["%d/verify with ntt" % (1 << logn)]
    assert [c.n_inst, c.n_wit, c.n_cons] == want
    # SURVEY.md App. C structure numbers
    nnz = {9: (720500, 106057, 55841, 17), 10: (2489465, 212043, 111652, 18)}[logn]
    assert (c.nnz_a, c.nnz_b, c.nnz_c, c.domain_log2) == nnz


def test_readme_counts_schoolbook_512(circuits):
    c = circuits(9, 1)
    assert [c.n_inst, c.n_wit, c.n_cons] == gold("readme_counts.json")["512/verify with schoolbook"]
    assert (c.nnz_a, c.nnz_b, c.nnz_c, c.domain_log2) == (655478, 461129, 294433, 19)


def test_readme_counts_schoolbook_1024(circuits):
    """README.md:45: Falcon-1024 verify with schoolbook = 2,049 / 1,150,004 / 1,156,150 (BASELINE configs[3])"""
    c = circuits(10, 1)
    assert [c.n_inst, c.n_wit, c.n_cons] == gold("readme_counts.json")["1024/verify with schoolbook"]
    assert [c.n_inst, c.n_wit, c.n_cons] == [2049, 1150004, 1156150]
    assert (c.nnz_a, c.nnz_b, c.nnz_c, c.domain_log2) == (2359419, 1708619, 1113124, 21)


@pytest.mark.parametrize("logn", [9, 10])
def test_ntt_conversion_gadget(oracle, logn):
    """test_ntt_mul_circuit (gadgets/poly.rs:252-301) + README 'ntt conversion' row."""
    import ctypes as C
    n = 1 << logn
    rng = np.random.default_rng(5)
    want_counts = gold("readme_counts.json")["%d/ntt conversion" % n]
    for _ in range(3):
        poly = rng.integers(0, Q, n).astype(np.uint16)
        out = np.zeros(n, dtype=np.uint16)
        counts = np.zeros(3, dtype=np.uint64)
        sat = oracle.lib().orc_kat_ntt(logn, oracle.ptr(poly, oracle.u16p), oracle.ptr(out, oracle.u16p),
                                       oracle.ptr(counts))
        assert sat == 1
        assert [int(x) for x in counts] == want_counts
        assert (out == synth.ntt(poly, logn)).all()
        clear = np.zeros(n, dtype=np.uint16)
        oracle.lib().orc_ntt_clear(logn, oracle.ptr(poly, oracle.u16p), oracle.ptr(clear, oracle.u16p))
        assert (clear == out).all()


def test_gadget_known_answers(oracle):
    k = gold("gadget_kats.json")
    for a, b, sat in k["mod_q"]:
        s, ok, counts = oracle.kat(0, 10, [a], b)
        assert s == sat and bool(ok) == sat
        assert counts == [1, 29 + 2, 30 + 1]  # input + 29 gadget witnesses + expected; 30 rows + enforce_equal
    for a, b, c, sat in k["add_mod"]:
        s, ok, _ = oracle.kat(1, 10, [a, b], c)
        assert s == sat and bool(ok) == sat
    for a, b, c, sat in k["mul_mod"]:
        s, ok, _ = oracle.kat(2, 10, [a, b], c)
        assert s == sat and bool(ok) == sat
    for a, sat in k["less_than_q"]:
        s, _, counts = oracle.kat(3, 10, [a])
        assert s == sat
        assert counts == [1, 1 + 27, 29]
    for a, lt in k["less_than_6144"]:
        s, ok, counts = oracle.kat(5, 10, [a], int(lt))
        assert s == lt
        assert counts == [1, 1 + 16, 17 + 1]
    for a, sat in k["norm_bound_512"]:
        s, _, counts = oracle.kat(4, 9, [a])
        assert s == sat, a
        assert counts == [1, 1 + 50, 52]
    for a, sat in k["norm_bound_1024"]:
        s, _, counts = oracle.kat(4, 10, [a])
        assert s == sat, a
        assert counts == [1, 1 + 52, 54]


def test_gadget_random_paths(oracle):
    """the reference's 1000-iteration random loops, at 200 iterations"""
    rnd = random.Random(7)
    for _ in range(200):
        t = rnd.randrange(1 << 30)
        assert oracle.kat(0, 10, [t], t % Q)[0] is True
        assert oracle.kat(0, 10, [t], (t + 1) % Q)[0] is False
        t2 = rnd.randrange(1 << 30)
        assert oracle.kat(1, 10, [t, t2], (t + t2) % Q)[0] is True
        assert oracle.kat(1, 10, [t, t2], (t + t2 + 1) % Q)[0] is False
        u = rnd.randrange(1 << 15)
        assert oracle.kat(3, 10, [u])[0] == (u < Q)
        assert oracle.kat(5, 10, [u], int(u < 6144))[0] == (u < 6144)
        w = rnd.randrange(1 << 27)
        assert oracle.kat(4, 9, [w])[0] == (w < 34034726)
        assert oracle.kat(4, 10, [w])[0] == (w < 70265242)


def test_inner_product_mod(oracle):
    rnd = random.Random(3)
    for dim in (2, 7, 64, 512):
        a = [rnd.randrange(Q) for _ in range(dim)]
        b = [rnd.randrange(Q) for _ in range(dim)]
        c = sum(x * y for x, y in zip(a, b)) % Q
        s, ok, counts = oracle.kat(6, 10, a + b, c)
        assert s and ok
        assert counts[2] == dim + 30 + 1  # dim products + mod row + 29 range rows + enforce_equal
        assert not oracle.kat(6, 10, a + b, (c + 1) % Q)[0]


@pytest.mark.parametrize("logn", [9, 10])
def test_circuit_satisfied_on_valid_signature(circuits, logn):
    """test_ntt_verification_r1cs (circuits/falcon_ntt.rs:133-160), synthetic signature"""
    c = circuits(logn, 0)
    sig, pk, hm = synth.make_signatures(logn, 2, seed=11)
    for i in range(2):
        z, st, fu = c.witness(sig[i], pk[i], hm[i], construct_matrices=True)
        assert st == 0 and fu == -1
        z2, st2, _ = c.witness(sig[i], pk[i], hm[i], construct_matrices=False)
        assert st2 == 0 and (z == z2).all()
        assert c.r1cs_eval(z)[3] == -1
        # public inputs are pk_ntt ++ hm_ntt (examples/pok_sig.rs:33-44)
        pub = [int(x) for x in np.concatenate([synth.ntt(pk[i], logn), synth.ntt(hm[i], logn)])]
        zc = c_oracle_canonical(z[1:c.n_inst])
        assert zc == pub


def c_oracle_canonical(z):
    import oracle_lib
    return oracle_lib.limbs_to_ints(oracle_lib.fr_to_canonical(z))


def test_invalid_signature_is_rejected(circuits):
    c = circuits(9, 0)
    sig, pk, hm = synth.make_signatures(9, 1, seed=12)
    hm2 = hm[0].copy()
    hm2[3] = (int(hm2[3]) + 1) % Q  # v changes by one coefficient -> tiny norm change: still satisfiable
    sig2 = sig[0].copy()
    sig2[:] = 6000  # huge norm
    z, st, fu = c.witness(sig2, pk[0], hm[0], construct_matrices=True, panic_on_range=True)
    assert st == -2  # the reference would panic in enforce_less_than_norm_bound
    assert fu >= 0   # and the #[cfg(test)] build yields an unsatisfied system
    # flipping one witness of a valid assignment breaks exactly that gadget's rows
    z, st, fu = c.witness(sig[0], pk[0], hm[0])
    z[c.n_inst + 5] = z[c.n_inst + 6]
    assert c.r1cs_eval(z)[3] >= 0


@pytest.mark.slow
def test_groth16_proof_verifies(circuits, oracle):
    """examples/pok_sig.rs:30-47 on the oracle: setup, prove, check (in the exponent)."""
    c = circuits(9, 0)
    sig, pk, hm = synth.make_signatures(9, 1, seed=13)
    z, st, _ = c.witness(sig[0], pk[0], hm[0])
    assert st == 0
    P = c.setup(seed=99)
    r = oracle.fr_from_canonical(oracle.ints_to_limbs([0x1234567890abcdef1234567890abcdef]))[0]
    s = oracle.fr_from_canonical(oracle.ints_to_limbs([0xfedcba0987654321fedcba0987654321]))[0]
    proof, comp = c.prove(P, z, r, s)
    assert c.verify_trapdoor(P, z, r, s, proof)
    bad = proof.copy()
    bad[40] ^= 1
    assert not c.verify_trapdoor(P, z, r, s, bad)
    assert len(bytes(comp)) == 192
    for off, g2 in ((0, 0), (12, 1), (36, 0)):
        assert oracle.lib().orc_point_check(oracle.ptr(np.ascontiguousarray(proof[off:])), g2) == 1
