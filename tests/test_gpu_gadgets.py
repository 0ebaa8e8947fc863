"""The reference's gadget known-answer tests executed against the CUDA gadget entry points (frcs_gadget_*,
csrc/gadgets.cu) -- good and bad paths, as the reference's own #[cfg(test)] macros check them (is_satisfied and the
output value): gadgets/arithmetics.rs:346-361, 480-494; gadgets/range_proofs.rs:365-389, 442-474, 529-547;
gadgets/poly.rs:252-301 -- and bit for bit against the oracle's witness assignment."""
import json
import os

import numpy as np
import pytest

from falcon_r1cs_b200 import gadgets as G, synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
KATS = json.load(open(os.path.join(GOLD, "gadget_kats.json")))
Q = G.Q


def test_mod_q_known_answers(contexts, oracle):
    ctx = contexts(10)
    assert G.shape(ctx, G.MOD_Q) == (1, 29, 30)  # arithmetics.rs:103 "30 constraints"
    rows = KATS["mod_q"]
    res = G.mod_q(ctx, [r[0] for r in rows], expected=[r[1] for r in rows])
    for i, (a, b, sat) in enumerate(rows):
        assert bool(res.satisfied[i]) == sat, (a, b)
        assert (res.out[i] == b) == sat and res.out[i] == a % Q
        osat, oz, ofu = oracle.kat_z(0, 10, [a], b)
        assert osat == sat and res.first_unsat[i] == ofu
        assert (res.wit[i] == oz[1:]).all()  # the oracle's z starts with the operand
    # without the expected-output row the gadget alone is always satisfied; mod_q(12290) = 1
    res = G.mod_q(ctx, [12290, 0, Q - 1, (1 << 200) + 12345])
    assert res.satisfied.all() and res.out[0] == 1 and res.out[3] == ((1 << 200) + 12345) % Q
    t, b = res.wit_ints(3)[:2]
    assert t * Q + b == (1 << 200) + 12345


def test_mod_q_random_1000(contexts):
    """arithmetics.rs:363-369: 1000 random t < 2^30 against t % q, and the off-by-one expectation"""
    ctx = contexts(9)
    rng = np.random.default_rng(17)
    t = [int(x) for x in rng.integers(0, 1 << 30, 1000)]
    good = G.mod_q(ctx, t, expected=[x % Q for x in t])
    bad = G.mod_q(ctx, t, expected=[(x + 1) % Q for x in t])
    assert good.satisfied.all() and good.out == [x % Q for x in t]
    assert not bad.satisfied.any() and (bad.first_unsat == 30).all()  # only the enforce_equal row fails


def test_add_mod_known_answers_and_random(contexts, oracle):
    ctx = contexts(10)
    rows = KATS["add_mod"]
    res = G.add_mod(ctx, [r[0] for r in rows], [r[1] for r in rows], expected=[r[2] for r in rows])
    for i, (a, b, c, sat) in enumerate(rows):
        assert bool(res.satisfied[i]) == sat and (res.out[i] == c) == sat, (a, b, c)
        osat, oz, ofu = oracle.kat_z(1, 10, [a, b], c)
        assert osat == sat and res.first_unsat[i] == ofu and (res.wit[i] == oz[2:]).all()
    rng = np.random.default_rng(18)  # arithmetics.rs:496-503
    t1 = [int(x) for x in rng.integers(0, 1 << 30, 1000)]
    t2 = [int(x) for x in rng.integers(0, 1 << 30, 1000)]
    good = G.add_mod(ctx, t1, t2, expected=[(x + y) % Q for x, y in zip(t1, t2)])
    bad = G.add_mod(ctx, t1, t2, expected=[(x + y + 1) % Q for x, y in zip(t1, t2)])
    assert good.satisfied.all() and not bad.satisfied.any()


def test_less_than_q_known_answers_and_random(contexts, oracle):
    ctx = contexts(10)
    assert G.shape(ctx, G.LESS_THAN_Q) == (1, 27, 29)
    rows = KATS["less_than_q"]
    res = G.enforce_less_than_q(ctx, [r[0] for r in rows])
    for i, (a, sat) in enumerate(rows):
        assert bool(res.satisfied[i]) == sat, a
        assert res.status[i] == (0 if a < Q else -16)  # the non-test build panics (range_proofs.rs:58-60)
        osat, oz, ofu = oracle.kat_z(3, 10, [a])
        assert osat == sat and res.first_unsat[i] == ofu and (res.wit[i] == oz[1:]).all()
    rng = np.random.default_rng(19)  # range_proofs.rs:391-395
    t = [int(x) for x in rng.integers(0, 1 << 15, 1000)]
    res = G.enforce_less_than_q(ctx, t)
    assert [bool(s) for s in res.satisfied] == [x < Q for x in t]


def test_less_than_6144_known_answers_and_random(contexts, oracle):
    ctx = contexts(9)
    assert G.shape(ctx, G.LESS_THAN_6144) == (1, 16, 17)
    rows = KATS["less_than_6144"]
    res = G.is_less_than_6144(ctx, [r[0] for r in rows], enforce_true=True)
    for i, (a, sat) in enumerate(rows):
        assert bool(res.satisfied[i]) == sat, a
        osat, oz, ofu = oracle.kat_z(5, 9, [a], int(sat))
        assert osat == sat and res.first_unsat[i] == ofu and (res.wit[i] == oz[1:]).all()
    rng = np.random.default_rng(20)  # range_proofs.rs:549-553
    t = [int(x) for x in rng.integers(0, 1 << 15, 1000)]
    res = G.is_less_than_6144(ctx, t, enforce_true=True)
    assert [bool(s) for s in res.satisfied] == [x < 6144 for x in t]
    # as a plain Boolean (no enforce_equal): the value is x < 6144 for 14-bit inputs, rows satisfied
    small = [x for x in t if x < (1 << 14)]
    res = G.is_less_than_6144(ctx, small)
    assert res.satisfied.all() and res.out == [int(x < 6144) for x in small]


@pytest.mark.parametrize("logn", [9, 10])
def test_norm_bound_known_answers_and_random(contexts, oracle, logn):
    ctx = contexts(logn)
    assert G.shape(ctx, G.NORM_BOUND) == ((1, 50, 52) if logn == 9 else (1, 52, 54))
    rows = KATS["norm_bound_%d" % (1 << logn)]
    res = G.enforce_less_than_norm_bound(ctx, [r[0] for r in rows])
    for i, (a, sat) in enumerate(rows):
        assert bool(res.satisfied[i]) == sat, a
        assert res.status[i] == (0 if a < G.L2_BOUND[logn] else -17)
        osat, oz, ofu = oracle.kat_z(4, logn, [a])
        assert osat == sat and res.first_unsat[i] == ofu and (res.wit[i] == oz[1:]).all()
    rng = np.random.default_rng(21)  # range_proofs.rs:476-480
    t = [int(x) for x in rng.integers(0, 1 << 27, 1000)] + [G.L2_BOUND[logn] - 1, G.L2_BOUND[logn]]
    res = G.enforce_less_than_norm_bound(ctx, t)
    assert [bool(s) for s in res.satisfied] == [x < G.L2_BOUND[logn] for x in t]


@pytest.mark.parametrize("logn", [9, 10])
def test_ntt_circuit_equals_clear_text_ntt(contexts, oracle, logn):
    """test_ntt_mul_circuit (gadgets/poly.rs:252-301): 10 random polynomials, every output equals
    NTTPolynomial::from(&poly); plus the README 'ntt conversion' counts and the oracle's witness block"""
    ctx = contexts(logn)
    n = 1 << logn
    assert G.shape(ctx, G.NTT_CIRCUIT) == (n, 29 * n, 30 * n)  # README.md:43,54
    rng = np.random.default_rng(100 + logn)
    polys = np.concatenate([rng.integers(0, Q, (10, n)), np.full((1, n), Q - 1), np.zeros((1, n), np.int64)]).astype(np.uint16)
    res = G.NTTPolyVar.ntt_circuit(ctx, polys)
    assert res.satisfied.all() and (res.status == 0).all()
    assert (res.out == synth.ntt(polys, logn)).all()
    for i in (0, 10, 11):
        sat, out, oz, counts = oracle.kat_ntt_z(logn, polys[i])
        assert sat and counts == [0, 29 * n, 30 * n]
        assert (res.out[i] == out).all() and (res.wit[i] == oz).all()
    # mod_q inside the gadget: a_k = q t_k + b_k with b_k < q and the 14 bits of b_k behind it
    w = np.array(G.to_int(res.wit[0]), dtype=object).reshape(n, 29)
    k = int(rng.integers(0, n))
    assert int(w[k, 1]) == int(res.out[0, k]) and [int(x) for x in w[k, 2:16]] == [(int(w[k, 1]) >> j) & 1 for j in range(14)]


def test_ntt_param_is_the_reference_table():
    tab = G.ntt_param_var(10)
    assert [int(tab[i]) for i in range(4)] == [1, pow(7, 512, Q), pow(7, 256, Q), pow(7, 768, Q)]  # 7^bitrev10(i)


@pytest.mark.parametrize("logn", [9, 10])
def test_l2_norm_and_bound(contexts, logn):
    ctx = contexts(logn)
    n = 1 << logn
    rng = np.random.default_rng(200 + logn)
    small = lambda: (np.rint(rng.normal(0, 60, n)).astype(np.int64) % Q).astype(np.uint16)
    v, sig = small(), small()
    cent = lambda e: min(int(e), Q - int(e))
    want = sum(cent(e) ** 2 for e in list(v) + list(sig))
    norm = G.l2_norm_var(ctx, v, sig)
    assert norm == want and want < G.L2_BOUND[logn]
    assert G.enforce_less_than_norm_bound(ctx, [norm]).satisfied[0]
    bad = sig.copy()
    bad[:] = 6000
    big = G.l2_norm_var(ctx, v, bad)
    assert big > G.L2_BOUND[logn]
    r = G.enforce_less_than_norm_bound(ctx, [big])
    assert not r.satisfied[0] and r.status[0] == -17
