"""GPU parity for subsystem (3): Radix2EvaluationDomain primitives and
R1CStoQAP::witness_map against the oracle, bit-exact."""
import numpy as np
import pytest

from falcon_r1cs_b200 import synth

pytestmark = pytest.mark.gpu


def rand_fr(oracle, n, seed):
    rng = np.random.default_rng(seed)
    x = rng.integers(0, 1 << 62, size=(n, 4), dtype=np.uint64)  # < 2^254 < r: valid canonical values
    return oracle.fr_from_canonical(x)


@pytest.mark.parametrize("log_size", [1, 4, 9, 10, 11, 13, 17, 18, 20])
def test_domain_ops(contexts, oracle, log_size):
    ctx = contexts(9)
    n = 1 << log_size
    x = rand_fr(oracle, n, log_size)
    for op in (0, 1, 2, 3):
        if log_size >= 17 and op in (1, 2):
            continue
        want = x.copy()
        oracle.lib().orc_domain_op(log_size, op, oracle.ptr(want))
        got = ctx.domain_op(log_size, op, x)
        assert (got == want).all(), (log_size, op)
    # DIF forward followed by DIT inverse (the pairing used inside witness_map) is the identity
    assert (ctx.domain_op(log_size, 4, x) == x).all()


def test_domain_edge_vectors(contexts, oracle):
    ctx = contexts(9)
    L = 12
    n = 1 << L
    zero = np.zeros((n, 4), dtype=np.uint64)
    assert not ctx.domain_op(L, 0, zero).any()
    delta = zero.copy()
    delta[0] = oracle.fr_from_canonical(oracle.ints_to_limbs([1]))[0]
    got = ctx.domain_op(L, 0, delta)  # fft of a delta = all ones
    assert (got == delta[0]).all()
    # linearity: fft(a + b) = fft(a) + fft(b), checked via the oracle's field add on canonical ints
    a, b = rand_fr(oracle, n, 1), rand_fr(oracle, n, 2)
    R = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
    ai, bi = oracle.limbs_to_ints(oracle.fr_to_canonical(a)), oracle.limbs_to_ints(oracle.fr_to_canonical(b))
    s = oracle.fr_from_canonical(oracle.ints_to_limbs([(x + y) % R for x, y in zip(ai, bi)]))
    fa, fb, fs = [oracle.limbs_to_ints(oracle.fr_to_canonical(ctx.domain_op(L, 0, v))) for v in (a, b, s)]
    assert all((x + y) % R == z for x, y, z in zip(fa, fb, fs))


@pytest.mark.parametrize("logn", [9, 10])
def test_witness_map_bit_exact(contexts, circuits, logn):
    ctx, c = contexts(logn), circuits(logn, 0)
    sig, pk, hm = synth.make_signatures(logn, 2, seed=31)
    z, st = ctx.witness_batch(sig, pk, hm)
    for i in range(2):
        h = ctx.witness_map(z[i])
        want = c.witness_map(z[i])
        assert (h == want).all()
        assert not h[-1].any()  # deg h <= n-2: h_query has n-1 entries (ark-groth16 generator)
    # an unsatisfied assignment still maps identically (h is then not a polynomial quotient)
    zb = z[0].copy()
    zb[c.n_inst + 3] = zb[c.n_inst + 4]
    assert (ctx.witness_map(zb) == c.witness_map(zb)).all()
