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


EDGE8 = EDGE + [127, 128, 129, 255, 256, 257, 0x7f7f, 0x8080, 0x80_80_80_80_80_80_80_80, 0x7f_80_7f_80, (1 << 248) - 1,
                (1 << 247) + (1 << 7), R - 128, R - 129, 0xff << 120, 0x0101_0101_0101_0101]


@pytest.mark.parametrize("wb", [16, 8])
def test_msm_window_geometries_scalar_edges(contexts, oracle, pk512, wb):
    """both window geometries of the MSM subsystem (16-bit digits / 32768 buckets, 8-bit digits / 128 buckets) on the
    digit-boundary scalars: carries into the next window, the top window, r - 1, all-ones bytes"""
    ctx = contexts(9)
    bases = pk512.export("h_query")[:len(EDGE8)]
    s = scal(oracle, EDGE8)
    assert (ctx.msm_g1(bases, s, wb) == oracle.msm_g1(bases, s)).all()
    for i in range(len(EDGE8)):
        assert (ctx.msm_g1(bases[i:i + 1], s[i:i + 1], wb) == oracle.msm_g1(bases[i:i + 1], s[i:i + 1])).all(), hex(EDGE8[i])
    assert not ctx.msm_g1(bases[:4], np.zeros((4, 4), dtype=np.uint64), wb).any()
    rep = np.repeat(bases[:1], 300, axis=0)  # one bucket, P + P + ...: doubling branch, several slice levels
    ones = scal(oracle, [1] * 300)
    assert (ctx.msm_g1(rep, ones, wb) == oracle.msm_g1(rep, ones)).all()
    assert not ctx.msm_g1(np.repeat(bases[:1], 2, axis=0), scal(oracle, [5, R - 5]), wb).any()
    g2 = pk512.export("b_g2_query")
    g2 = g2[np.nonzero(g2.any(axis=1))[0][:len(EDGE8)]]
    assert (ctx.msm_g2(g2, s, wb) == oracle.msm_g2(g2, s)).all()


@pytest.mark.parametrize("wb", [16, 8])
def test_msm_window_geometries_witness_and_dense(contexts, circuits, oracle, pk512, wb):
    """a real assignment (zeros / ones / 14-bit values / 132-bit quotients) against a_query, b_g1_query (with points
    at infinity) and b_g2_query, and dense random scalars, through both geometries"""
    ctx, c = contexts(9), circuits(9, 0)
    sig, pk, hm = synth.make_signatures(9, 1, seed=43)
    z, st, _ = c.witness(sig[0], pk[0], hm[0])
    zc = oracle.fr_to_canonical(z)
    for name in ("a_query", "b_g1_query"):
        bases = pk512.export(name)
        assert (ctx.msm_g1(bases, zc, wb) == oracle.msm_g1(bases, zc)).all(), name
    b2 = pk512.export("b_g2_query")
    assert (ctx.msm_g2(b2, zc, wb) == oracle.msm_g2(b2, zc)).all()
    rng = np.random.default_rng(12)
    hq = pk512.export("h_query")[:20000]
    s = rng.integers(0, 1 << 62, size=(hq.shape[0], 4), dtype=np.uint64)
    assert (ctx.msm_g1(hq, s, wb) == oracle.msm_g1(hq, s)).all()


def test_msm_pair_levels_forced():
    """The batched-affine pair levels (msm_impl.cuh pair_kernel: affine + affine with one inversion per 256 pairs) are
    chosen by batch size; FRCS_MSM_PAIR=2 forces them at every size, so the edge cases above (doublings inside a bucket,
    P and -P, points at infinity, single-entry buckets, one scalar at a time) run through pair_classify / pair_finish.
    The switch is read once per process, hence the subprocess."""
    import os
    import subprocess
    import sys
    if os.environ.get("FRCS_MSM_PAIR") == "2":
        pytest.skip("already inside the forced run")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, FRCS_MSM_PAIR="2")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_gpu_msm.py"), "-x", "-q", "-m", "gpu",
                        "-k", "edges or small or witness_like or window_geometries"],
                       cwd=root, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
