"""GPU parity for subsystem (4): G1/G2 MSM through the C ABI vs the oracle's Pippenger
(restating ark-ec's VariableBaseMSM), affine results bit-exact."""
import random

import numpy as np
import pytest

from falcon_r1cs_b200 import synth

pytestmark = pytest.mark.gpu
R = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001

_PK = {}


@pytest.fixture(scope="module")
def pk512(circuits):
    if "pk" not in _PK:
        _PK["pk"] = circuits(9, 0).setup(seed=5)
    return _PK["pk"]


def scal(oracle, vals):
    return oracle.ints_to_limbs([v % R for v in vals])


EDGE = [0, 1, 2, R - 1, R - 2, 1 << 15, (1 << 15) + 1, (1 << 16) - 1, 1 << 16, (1 << 16) + 1, 0x8000_8000_8000_8000,
        0xffff_ffff_ffff_ffff, (1 << 31), (1 << 32) - 1, (1 << 128) - 1, 0x7fff_8000, 0x8000_7fff_ffff,
        (1 << 254), (1 << 240) | (1 << 15), 12289, 70265242]


@pytest.mark.parametrize("n", [1, 2, 33, 1000])
def test_msm_g1_small(contexts, oracle, pk512, n):
    ctx = contexts(9)
    rnd = random.Random(n)
    bases = pk512.export("h_query")[:n]
    s = scal(oracle, [rnd.randrange(R) for _ in range(n)])
    assert (ctx.msm_g1(bases, s) == oracle.msm_g1(bases, s)).all()


def test_msm_g1_scalar_edges(contexts, oracle, pk512):
    ctx = contexts(9)
    bases = pk512.export("h_query")[:len(EDGE)]
    s = scal(oracle, EDGE)
    assert (ctx.msm_g1(bases, s) == oracle.msm_g1(bases, s)).all()
    # one scalar at a time (each digit pattern alone), and all-zero scalars -> infinity
    for i in range(len(EDGE)):
        assert (ctx.msm_g1(bases[i:i + 1], s[i:i + 1]) == oracle.msm_g1(bases[i:i + 1], s[i:i + 1])).all(), hex(EDGE[i])
    z = np.zeros((4, 4), dtype=np.uint64)
    assert not ctx.msm_g1(bases[:4], z).any()
    # repeated base: P + P + ... exercises the doubling branch of the mixed addition
    rep = np.repeat(bases[:1], 40, axis=0)
    ones = scal(oracle, [1] * 40)
    assert (ctx.msm_g1(rep, ones) == oracle.msm_g1(rep, ones)).all()
    # P and -P cancel to infinity inside one bucket
    neg = scal(oracle, [5, R - 5])
    assert not ctx.msm_g1(np.repeat(bases[:1], 2, axis=0), neg).any()


def test_msm_g1_witness_like_scalars(contexts, circuits, oracle, pk512):
    """the a_query MSM of a real assignment: ~37% zeros, ~54% ones, points at infinity in b_g1"""
    ctx, c = contexts(9), circuits(9, 0)
    sig, pk, hm = synth.make_signatures(9, 1, seed=41)
    z, st, _ = c.witness(sig[0], pk[0], hm[0])
    zc = oracle.fr_to_canonical(z)
    for name in ("a_query", "b_g1_query"):
        bases = pk512.export(name)
        assert (ctx.msm_g1(bases, zc) == oracle.msm_g1(bases, zc)).all(), name


def test_msm_g1_dense_h_sized(contexts, oracle, pk512):
    ctx = contexts(9)
    bases = pk512.export("h_query")
    rng = np.random.default_rng(8)
    s = rng.integers(0, 1 << 62, size=(bases.shape[0], 4), dtype=np.uint64)
    assert (ctx.msm_g1(bases, s) == oracle.msm_g1(bases, s)).all()


def test_msm_g2(contexts, circuits, oracle, pk512):
    ctx, c = contexts(9), circuits(9, 0)
    bases = pk512.export("b_g2_query")
    rnd = random.Random(2)
    for n in (1, 3, 500):
        nz = bases[np.nonzero(bases.any(axis=1))[0][:n]]
        s = scal(oracle, [rnd.randrange(R) for _ in range(n)])
        assert (ctx.msm_g2(nz, s) == oracle.msm_g2(nz, s)).all()
    e = scal(oracle, EDGE)
    assert (ctx.msm_g2(bases[:len(EDGE)], e) == oracle.msm_g2(bases[:len(EDGE)], e)).all()
    sig, pk, hm = synth.make_signatures(9, 1, seed=42)
    z, st, _ = c.witness(sig[0], pk[0], hm[0])
    zc = oracle.fr_to_canonical(z)
    assert (ctx.msm_g2(bases, zc) == oracle.msm_g2(bases, zc)).all()
