"""CPU tests of the product's host side: the C-ABI library loads and exports every
symbol the header declares; the host build of the field/curve source agrees with
Python big integers and with the oracle; no compute call works without a GPU."""
import os
import random
import re

import numpy as np
import pytest

from falcon_r1cs_b200 import lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R_MOD = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
Q_MOD = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab


def limbs(x, n):
    return [(x >> (64 * i)) & (2 ** 64 - 1) for i in range(n)]


def fromlimbs(a):
    return sum(int(v) << (64 * i) for i, v in enumerate(a))


def selftest(op, on_device, inp, out_words):
    inp = np.ascontiguousarray(inp, dtype=np.uint64)
    out = np.zeros((inp.shape[0], out_words), dtype=np.uint64)
    rc = L.load().frcs_selftest(op, on_device, inp.ctypes.data_as(L.u64p), inp.shape[0], out.ctypes.data_as(L.u64p))
    assert rc == 0, L.load().frcs_last_error()
    return out


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "falcon_r1cs_b200.h")).read()
    declared = set(re.findall(r"\b(frcs_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"frcs_ctx", "frcs_shape", "frcs_pk_view"}
    assert declared == set(L.PROTOTYPES), declared ^ set(L.PROTOTYPES)
    lib = L.load()
    for name in declared:
        assert hasattr(lib, name), name


def field_cases(mod, nl, seed):
    rnd = random.Random(seed)
    xs = [rnd.randrange(mod) for _ in range(100)]
    ys = [rnd.randrange(mod) for _ in range(100)]
    xs[:4] = [0, mod - 1, 1, mod - 1]
    ys[:4] = [5, mod - 1, 0, 1]
    return xs, ys


def check_field(on_device, mod, nl, base):
    R = 1 << (64 * nl)
    xs, ys = field_cases(mod, nl, base)
    inp = np.array([limbs(x, nl) + limbs(y, nl) for x, y in zip(xs, ys)], dtype=np.uint64)
    for op, f in ((0, lambda a, b: a * b * pow(R, -1, mod) % mod), (1, lambda a, b: (a + b) % mod),
                  (2, lambda a, b: (a - b) % mod)):
        out = selftest(base + op, on_device, inp, nl)
        for i in range(len(xs)):
            assert fromlimbs(out[i]) == f(xs[i], ys[i]), (op, i)
    inp1 = np.array([limbs(x, nl) for x in xs[:12]], dtype=np.uint64)
    for op in (3, 6):      # the bit-by-bit power and the windowed one (inverse_w4, used by the batched-affine MSM levels)
        out = selftest(base + op, on_device, inp1, nl)
        for i in range(12):
            x = xs[i] * pow(R, -1, mod) % mod
            assert fromlimbs(out[i]) == ((pow(x, -1, mod) * R % mod) if x else 0), (op, i)


    if base == 10:  # Fq: a*b - c*d with one Montgomery reduction (mul_sub2_inline, the Y3 of the XYZZ additions)
        rnd = random.Random(base + 18)
        edge = [0, 1, mod - 1, mod - 2, (1 << 380) - 1, (mod - 1) ^ ((1 << 64) - 1)]
        quads = [(a, b, c, d) for a in edge for b in edge[:3] for c in edge for d in edge[:3]]
        quads += [tuple(rnd.randrange(mod) for _ in range(4)) for _ in range(1500)]
        out = selftest(18, on_device, np.array([sum((limbs(v, nl) for v in q), []) for q in quads], dtype=np.uint64), nl)
        Rinv = pow(R, -1, mod)
        for i, (a, b, c, d) in enumerate(quads):
            assert fromlimbs(out[i]) == (a * b - c * d) * Rinv % mod, (18, i)

    # squaring (on the device: the dedicated wide-square + reduction of ff32.cuh), incl. all-ones limb patterns
    rnd = random.Random(base + 7)
    top = mod.bit_length()
    sq = xs + [rnd.randrange(mod) for _ in range(900)] + [(1 << (top - 1)) - 1, (1 << (top - 1)), mod - 2, (1 << 32) - 1,
                                                           (1 << (32 * (2 * nl - 1))) - 1, ((1 << (top - 1)) - 1) ^ ((1 << 64) - 1)]
    out = selftest(base + 7, on_device, np.array([limbs(x, nl) for x in sq], dtype=np.uint64), nl)
    Rinv = pow(R, -1, mod)
    for i, x in enumerate(sq):
        assert fromlimbs(out[i]) == x * x * Rinv % mod, (7, i, hex(x))


def test_host_field_arithmetic_vs_python_ints():
    check_field(0, R_MOD, 4, 0)
    check_field(0, Q_MOD, 6, 10)


def curve_cases(oracle, g2, count, seed):
    """random multiples of the generator, from the oracle"""
    rnd = random.Random(seed)
    g1 = np.zeros(12, dtype=np.uint64)
    gg2 = np.zeros(24, dtype=np.uint64)
    oracle.lib().orc_generators(oracle.ptr(g1), oracle.ptr(gg2))
    gen, w = (gg2, 24) if g2 else (g1, 12)
    mul = oracle.lib().orc_g2_mul if g2 else oracle.lib().orc_g1_mul
    pts, ks = [], []
    for _ in range(count):
        k = rnd.randrange(1, R_MOD)
        kk = np.array(limbs(k, 4), dtype=np.uint64)
        out = np.zeros(w, dtype=np.uint64)
        mul(oracle.ptr(gen), oracle.ptr(kk), oracle.ptr(out))
        pts.append(out)
        ks.append(k)
    return gen, pts, ks, mul, w


def check_curve(oracle, on_device, g2):
    gen, pts, ks, mul, w = curve_cases(oracle, g2, 6, 17 + g2)
    base = 30 if g2 else 20
    # add: k0*G + k1*G == (k0+k1)*G
    inp = np.array([np.concatenate([pts[i], pts[i + 1]]) for i in range(5)] + [np.concatenate([pts[0], pts[0]])])
    out = selftest(base, on_device, inp, w)
    for i in range(6):
        k = (ks[i] + ks[i + 1]) % R_MOD if i < 5 else 2 * ks[0] % R_MOD
        want = np.zeros(w, dtype=np.uint64)
        kk = np.array(limbs(k, 4), dtype=np.uint64)
        mul(oracle.ptr(gen), oracle.ptr(kk), oracle.ptr(want))
        assert (out[i] == want).all(), i
    # P + (-P) = infinity, P + inf = P
    neg = pts[0].copy()
    half = w // 2
    yneg = [(Q_MOD - fromlimbs(neg[half + 6 * j: half + 6 * j + 6])) % Q_MOD for j in range(half // 6)]
    for j, v in enumerate(yneg):
        neg[half + 6 * j: half + 6 * j + 6] = limbs(v, 6)
    inp = np.array([np.concatenate([pts[0], neg]), np.concatenate([pts[0], np.zeros(w, dtype=np.uint64)])])
    out = selftest(base, on_device, inp, w)
    assert not out[0].any() and (out[1] == pts[0]).all()
    # the same sums through the affine pair formulas (pair_classify / pair_finish): generic, doubling, cancellation,
    # infinity on either side, infinity twice
    zero = np.zeros(w, dtype=np.uint64)
    inp = np.array([np.concatenate([pts[i], pts[i + 1]]) for i in range(5)] + [np.concatenate([pts[0], pts[0]])] +
                   [np.concatenate([pts[0], neg]), np.concatenate([pts[0], zero]), np.concatenate([zero, pts[1]]),
                    np.concatenate([zero, zero])])
    out = selftest(base + 4, on_device, inp, w)
    ref = selftest(base, on_device, inp[:6], w)
    assert (out[:6] == ref).all()
    assert not out[6].any() and (out[7] == pts[0]).all() and (out[8] == pts[1]).all() and not out[9].any()
    # double and scalar mul
    out = selftest(base + 1, on_device, np.array(pts[:2]), w)
    for i in range(2):
        want = np.zeros(w, dtype=np.uint64)
        kk = np.array(limbs(2 * ks[i] % R_MOD, 4), dtype=np.uint64)
        mul(oracle.ptr(gen), oracle.ptr(kk), oracle.ptr(want))
        assert (out[i] == want).all()
    k2 = 0x1234567890abcdef1234567890abcdef1234567890abcdef
    inp = np.array([np.concatenate([pts[0], np.array(limbs(k2, 4), dtype=np.uint64)])])
    out = selftest(base + 2, on_device, inp, w)
    want = np.zeros(w, dtype=np.uint64)
    kk = np.array(limbs(k2, 4), dtype=np.uint64)
    mul(oracle.ptr(pts[0]), oracle.ptr(kk), oracle.ptr(want))
    assert (out[0] == want).all()


def test_host_curve_arithmetic_vs_oracle(oracle):
    check_curve(oracle, 0, 0)
    check_curve(oracle, 0, 1)


def test_no_cpu_fallback():
    """Without a CUDA device every compute entry point must fail loudly."""
    import ctypes as C
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("GPU present")
    h = C.c_void_p()
    rc = L.load().frcs_ctx_create(9, 0, 0, C.byref(h))
    assert rc == L.E_CUDA and not h.value
    assert b"no CPU fallback" in L.load().frcs_last_error() or b"cuda" in L.load().frcs_last_error().lower()


@pytest.mark.gpu
def test_device_field_and_curve_arithmetic(oracle):
    check_field(1, R_MOD, 4, 0)
    check_field(1, Q_MOD, 6, 10)
    check_curve(oracle, 1, 0)
    check_curve(oracle, 1, 1)


def test_host_proof_compress_matches_oracle(oracle):
    """ark-serialize compressed Proof: product host code vs oracle, on points from the oracle"""
    from falcon_r1cs_b200 import api
    gen1, pts1, _, _, _ = curve_cases(oracle, 0, 4, 91)
    gen2, pts2, _, _, _ = curve_cases(oracle, 1, 2, 92)
    for a, b, c in ((pts1[0], pts2[0], pts1[1]), (pts1[2], pts2[1], pts1[3]),
                    (np.zeros(12, dtype=np.uint64), np.zeros(24, dtype=np.uint64), pts1[0])):
        proof = np.concatenate([a, b, c])
        want = np.zeros(192, dtype=np.uint8)
        oracle.lib().orc_compress_proof(oracle.ptr(proof), oracle.ptr(want, oracle.u8p))
        assert api.proof_compress(proof) == bytes(want)


def test_host_only_entry_points_reject_null_arguments():
    """the entry points that need no GPU (combine, verify, compress) validate their arguments"""
    import ctypes as C
    lib = L.load()
    z = np.zeros(48, dtype=np.uint64)
    assert lib.frcs_combine_partials(0, 1, z.ctypes.data_as(L.u64p), z.ctypes.data_as(L.u64p), z.ctypes.data_as(L.u64p),
                                     z.ctypes.data_as(L.u64p)) == L.E_INVALID_ARG
    assert lib.frcs_combine_partials(1, 1, None, z.ctypes.data_as(L.u64p), z.ctypes.data_as(L.u64p),
                                     z.ctypes.data_as(L.u64p)) == L.E_INVALID_ARG
    assert lib.frcs_verify_proof(None, None, None, 0, None, None) == L.E_INVALID_ARG
    assert lib.frcs_proof_compress(None, None) == L.E_INVALID_ARG
    assert lib.frcs_pairing_is_one(None, None) == L.E_INVALID_ARG


def test_combine_partials_of_infinity_is_infinity():
    """all-zero MSM sums (every shard empty) combine to the point at infinity in A, B and C"""
    parts = np.zeros((2, 1, 144), dtype=np.uint64)
    from falcon_r1cs_b200 import api
    r = np.zeros(4, dtype=np.uint64)
    proof = api.combine_partials(parts, r, r)
    assert not proof.any()


@pytest.mark.parametrize("bits", [0, 28, 32])
def test_long_row_digits_reconstruct_the_coefficient(bits):
    """host logic behind r1cs_signed_long_kernel / r1cs_bundle_kernel: five balanced digits d_i with
    sum d_i 2^(w i) == c or c - r, each |d_i| <= 2^(w-1); coefficients too large for five digits are refused"""
    import ctypes as C
    w = bits or 32
    rnd = random.Random(1000 + bits)
    lim = 1 << (5 * w - 1)  # balanced range is a little short of +-2^(5w-1)
    vals = [0, 1, -1, 12289, -12289, (1 << 26), -(1 << 26), lim - (1 << (4 * w)), -(lim - (1 << (4 * w)))]
    vals += [rnd.randrange(-(1 << k), 1 << k) for k in (13, 40, 100, 136, 5 * w - 2) for _ in range(40)]
    too_big = [1 << 200, -(1 << 180), R_MOD // 2, (1 << (5 * w)) + 5, -(1 << (5 * w)) - 7]
    allv = vals + too_big
    inp = np.array([limbs(v % R_MOD, 4) for v in allv], dtype=np.uint64)
    dig = np.zeros((len(allv), 5), dtype=np.int64)
    ok = np.zeros(len(allv), dtype=np.int32)
    rc = L.load().frcs_debug_digits(inp.ctypes.data_as(L.u64p), len(allv), bits, dig.ctypes.data_as(C.POINTER(C.c_int64)),
                                    ok.ctypes.data_as(C.POINTER(C.c_int32)))
    assert rc == 0, L.load().frcs_last_error()
    for i, v in enumerate(vals):
        assert ok[i] == 1, (i, v)
        assert sum(int(dig[i, k]) << (w * k) for k in range(5)) == v, (i, v)
        assert all(-(1 << (w - 1)) <= int(dig[i, k]) < (1 << (w - 1)) for k in range(5))
    assert (ok[len(vals):] == 0).all()
    assert L.load().frcs_debug_digits(None, 0, 28, None, None) != 0 and L.load().frcs_debug_digits(
        inp.ctypes.data_as(L.u64p), 1, 7, dig.ctypes.data_as(C.POINTER(C.c_int64)), ok.ctypes.data_as(C.POINTER(C.c_int32))) != 0
