"""GPU parity: circuit matrices, witness generation (subsystem 1) and R1CS evaluation
(subsystem 2) through the C ABI against the oracle, bit-exact."""
import numpy as np
import pytest

from falcon_r1cs_b200 import synth

pytestmark = pytest.mark.gpu
Q = 12289


@pytest.mark.parametrize("logn", [9, 10])
def test_matrices_equal_oracle(contexts, circuits, logn):
    ctx, c = contexts(logn), circuits(logn, 0)
    assert (ctx.n_inst, ctx.n_wit, ctx.n_cons, ctx.domain_log2) == (c.n_inst, c.n_wit, c.n_cons, c.domain_log2)
    assert ctx.nnz == (c.nnz_a, c.nnz_b, c.nnz_c)
    for which in range(3):
        rp, col, val = ctx.get_matrix(which)
        orp, ocol, oval = c.csr(which)
        assert (rp == orp).all()
        assert (col == ocol).all()
        assert (val == oval).all()


@pytest.mark.parametrize("logn", [9, 10])
def test_witness_bit_exact(contexts, circuits, logn):
    ctx, c = contexts(logn), circuits(logn, 0)
    n = 24
    sig, pk, hm = synth.make_signatures(logn, n, seed=21)
    z, st = ctx.witness_batch(sig, pk, hm)
    assert (st == 0).all()
    for i in range(n):
        zo, sto, _ = c.witness(sig[i], pk[i], hm[i])
        assert sto == 0
        bad = np.nonzero((z[i] != zo).any(axis=1))[0]
        assert bad.size == 0, (i, bad[:10])


@pytest.mark.parametrize("logn", [9, 10])
def test_witness_edge_cases(contexts, circuits, logn):
    ctx, c = contexts(logn), circuits(logn, 0)
    n = 1 << logn
    rng = np.random.default_rng(4)
    cases = []
    cases.append((np.zeros(n), np.zeros(n), np.zeros(n)))                       # all zero
    cases.append((np.full(n, Q - 1), np.full(n, Q - 1), np.full(n, Q - 1)))     # maximal coefficients
    cases.append((rng.integers(0, Q, n), rng.integers(0, Q, n), rng.integers(0, Q, n)))  # random: norm too big
    e = np.zeros(n); e[0] = 1
    cases.append((e, rng.integers(0, Q, n), rng.integers(0, Q, n)))
    sig = np.array([x[0] for x in cases], dtype=np.uint16)
    pk = np.array([x[1] for x in cases], dtype=np.uint16)
    hm = np.array([x[2] for x in cases], dtype=np.uint16)
    z, st = ctx.witness_batch(sig, pk, hm)
    for i in range(len(cases)):
        zo, sto, _ = c.witness(sig[i], pk[i], hm[i], panic_on_range=True)
        assert (z[i] == zo).all(), i
        want = {0: 0, -1: -16, -2: -17}[sto]
        assert st[i] == want, (i, st[i], sto)


@pytest.mark.parametrize("logn", [9, 10])
def test_r1cs_eval_bit_exact(contexts, circuits, logn):
    ctx, c = contexts(logn), circuits(logn, 0)
    sig, pk, hm = synth.make_signatures(logn, 3, seed=22)
    z, st = ctx.witness_batch(sig, pk, hm)
    # corrupt one assignment: first violated row must match the oracle's
    z[1, c.n_inst + 7] = z[1, c.n_inst + 8] + np.uint64(1)
    z[2, 5] = z[2, 6]
    az, bz, cz, fu = ctx.r1cs_eval_batch(z)
    for i in range(3):
        oa, ob, oc, ofu = c.r1cs_eval(z[i])
        assert (az[i] == oa).all() and (bz[i] == ob).all() and (cz[i] == oc).all()
        assert fu[i] == ofu
    assert fu[0] == -1 and fu[1] >= 0 and fu[2] >= 0
    assert ctx.is_satisfied(z[0]) and not ctx.is_satisfied(z[1])


def test_r1cs_eval_random_z(contexts, circuits):
    """arbitrary (non-witness) z: full-width values everywhere"""
    ctx, c = contexts(9), circuits(9, 0)
    rng = np.random.default_rng(9)
    z = rng.integers(0, 1 << 62, size=(c.n_z, 4), dtype=np.uint64)
    az, bz, cz, fu = ctx.r1cs_eval_batch(z)
    oa, ob, oc, ofu = c.r1cs_eval(z)
    assert (az[0] == oa).all() and (bz[0] == ob).all() and (cz[0] == oc).all() and fu[0] == ofu


def test_r1cs_eval_large_batch_bit_exact(contexts, circuits):
    """batches of 64 and more take the lane-per-signature kernel for the long rows: outputs and first violated row
    identical to the oracle, incl. an assignment whose NTT inputs are not small (exact fall-back) and a flipped bit"""
    ctx, c = contexts(9), circuits(9, 0)
    n = 72
    sig, pk, hm = synth.make_signatures(9, n, seed=25)
    z, st = ctx.witness_batch(sig, pk, hm)
    assert (st == 0).all()
    z[3, c.n_inst + 5] = z[3, c.n_inst + 900]   # a signature coefficient replaced by some other witness entry
    z[40, c.n_inst + 2] = np.uint64(0x123456789ABCDEF)  # ... by a value that is not small
    z[71, c.n_z - 3] = z[71, 0]
    az, bz, cz, fu = ctx.r1cs_eval_batch(z)
    for i in (0, 3, 40, 63, 64, 71):
        oa, ob, oc, ofu = c.r1cs_eval(z[i])
        assert (az[i] == oa).all() and (bz[i] == ob).all() and (cz[i] == oc).all(), i
        assert fu[i] == ofu, i
    assert fu[0] == -1 and fu[3] >= 0 and fu[40] >= 0
    good = np.ones(n, bool)
    good[[3, 40, 71]] = False
    assert (fu[good] == -1).all()


def test_witness_batch_large_roundtrip(contexts):
    """size-independent property at a larger batch: every z satisfies the R1CS"""
    ctx = contexts(10)
    n = 96
    sig, pk, hm = synth.make_signatures(10, n, seed=23)
    z, st = ctx.witness_batch(sig, pk, hm)
    assert (st == 0).all()
    _, _, _, fu = ctx.r1cs_eval_batch(z, want=False)
    assert (fu == -1).all()
    # public inputs are the clear-text NTTs (examples/pok_sig.rs:33-44)
    one = z[0, 0]
    assert (z[:, 0] == one).all()


def test_witness_check_batch_fused(contexts, circuits):
    """BASELINE configs[2] entry point: generate + which_is_unsatisfied without exporting z; an invalid signature
    in the batch is reported through its status, the valid ones are satisfied"""
    ctx, c = contexts(9), circuits(9, 0)
    n = 40
    sig, pk, hm = synth.make_signatures(9, n, seed=24)
    sig[7] = 6000  # norm far too large: the reference panics in enforce_less_than_norm_bound
    fu, st = ctx.witness_check_batch(sig, pk, hm)
    for i in range(n):
        _, sto, ofu = c.witness(sig[i], pk[i], hm[i], construct_matrices=True, panic_on_range=True)
        assert st[i] == {0: 0, -1: -16, -2: -17}[sto], i
        if sto == 0:
            assert fu[i] == -1 == ofu
    assert st[7] == -17 and (np.delete(st, 7) == 0).all()


def _fr_small(oracle, v):
    """Montgomery image of a small integer as one z entry (4 x u64)"""
    return oracle.fr_from_canonical(oracle.ints_to_limbs([v]))[0]


def test_r1cs_eval_large_batch_f1024_bit_exact(contexts, circuits, oracle):
    """BASELINE configs[2] at its own size: Falcon-1024, batch >= 128, i.e. the lane-per-signature bundle kernels
    (r1cs_bundle_kernel + r1cs_bundle_finish_kernel) on 1,027-term rows.  az / bz / cz and the first violated row
    against the oracle for spread indices: corrupted NTT inputs (in range, above q, just below and just above the
    bundle's multiplicand limit ~2^16, far above it, not small at all), a flipped norm bit, a flipped range bit, and
    the valid neighbour of each."""
    logn = 10
    ctx, c = contexts(logn), circuits(logn, 0)
    n = 136
    sig, pk, hm = synth.make_signatures(logn, n, seed=26)
    z, st = ctx.witness_batch(sig, pk, hm)
    assert (st == 0).all()
    N, ni = 1 << logn, c.n_inst
    w_norm = 2 * N + 27 * N + 2 * 29 * N + 30 * N + 36 * N
    bad = {
        5: (ni + 17, _fr_small(oracle, (int(sig[5, 17]) + 1) % Q)),     # sig coefficient off by one (in range)
        33: (ni + N + 3, _fr_small(oracle, 12290)),                     # v coefficient above q, far below the limit
        34: (ni + 900, _fr_small(oracle, 60000)),                       # below the bundle's multiplicand limit
        66: (ni + 901, _fr_small(oracle, 70000)),                       # just above it: exact fall-back
        67: (ni + N + 1000, _fr_small(oracle, (1 << 27) + 5)),          # small view holds it, far above the limit
        99: (ni + 2, np.array([0x123456789ABCDEF, 7, 9, 11], np.uint64)),  # not small at all
        100: (ni + w_norm + 26, None),                                  # top norm bit flipped
        127: (ni + 2 * N + 27 * 5 + 3, None),                           # a range bit of v[5] flipped
        135: (3, _fr_small(oracle, 4242)),                              # a public input (pk_ntt[2]) replaced
    }
    one = z[0, 0].copy()
    for i, (col, val) in bad.items():
        if val is None:
            val = one if not z[i, col].any() else np.zeros(4, np.uint64)
        z[i, col] = val
    az, bz, cz, fu = ctx.r1cs_eval_batch(z)
    check = sorted(set(bad) | {i + 1 for i in bad if i + 1 < n} | {0, 63, 64, 128})
    for i in check:
        oa, ob, oc, ofu = c.r1cs_eval(z[i])
        for name, got, want in (("az", az[i], oa), ("bz", bz[i], ob), ("cz", cz[i], oc)):
            rows = np.nonzero((got != want).any(axis=1))[0]
            assert rows.size == 0, (i, name, rows[:8])
        assert fu[i] == ofu, (i, fu[i], ofu)
        assert (fu[i] >= 0) == (i in bad), i
    good = np.ones(n, bool)
    good[list(bad)] = False
    assert (fu[good] == -1).all()
    # the verdict-only entry (no az/bz/cz buffers: the path frcs_witness_check_batch takes) agrees
    _, _, _, fu2 = ctx.r1cs_eval_batch(z, want=False)
    assert (fu2 == fu).all()


@pytest.mark.slow
@pytest.mark.parametrize("logn", [9, 10])
def test_witness_bit_exact_1000(contexts, circuits, logn):
    """z bit-identical for >= 10^3 random signatures, both parameter sets (generated on
    the GPU in chunks, compared with the oracle on the host threads)"""
    from concurrent.futures import ThreadPoolExecutor
    ctx, c = contexts(logn), circuits(logn, 0)
    total, chunk = 1024, 128
    with ThreadPoolExecutor(max_workers=16) as ex:
        for c0 in range(0, total, chunk):
            sig, pk, hm = synth.make_signatures(logn, chunk, seed=77, first=c0)
            z, st = ctx.witness_batch(sig, pk, hm)
            assert (st == 0).all()

            def cmp(i):
                zo, sto, _ = c.witness(sig[i], pk[i], hm[i])
                return sto == 0 and np.array_equal(z[i], zo)
            ok = list(ex.map(cmp, range(chunk)))
            assert all(ok), (c0, [i for i, o in enumerate(ok) if not o][:8])


@pytest.mark.parametrize("logn,n", [(9, 3), (10, 66)])
def test_r1cs_eval_generic_long_row_path(contexts, circuits, oracle, monkeypatch, logn, n):
    """The rows of the ntt_circuit blocks normally go through the butterfly network (r1cs_ntt_rows_kernel).  With
    FRCS_NO_NTT_ROWS=1 they take the generic long-row kernels over the inlined matrix rows (signed-digit warp-per-row
    below 64 signatures, DFMA bundles from 64): both must give the oracle's A z, B z, C z and first violated row, on
    valid assignments, small corruptions and values that are not small."""
    ctx, c = contexts(logn), circuits(logn, 0)
    sig, pk, hm = synth.make_signatures(logn, n, seed=27)
    z, st = ctx.witness_batch(sig, pk, hm)
    z[1, c.n_inst + 4] = _fr_small(oracle, 12290)
    z[n - 1, c.n_inst + 9] = np.array([0xFEDCBA987654321, 3, 5, 7], np.uint64)
    z[2, 0] = _fr_small(oracle, 2)        # the constant column is not One
    results = []
    for env in ("", "1"):
        if env:
            monkeypatch.setenv("FRCS_NO_NTT_ROWS", env)
        else:
            monkeypatch.delenv("FRCS_NO_NTT_ROWS", raising=False)
        results.append(ctx.r1cs_eval_batch(z))
    monkeypatch.delenv("FRCS_NO_NTT_ROWS", raising=False)
    for az, bz, cz, fu in results:
        for i in sorted({0, 1, 2, n - 1}):
            oa, ob, oc, ofu = c.r1cs_eval(z[i])
            assert (az[i] == oa).all() and (bz[i] == ob).all() and (cz[i] == oc).all(), i
            assert fu[i] == ofu, i
    assert (results[0][3] == results[1][3]).all()


@pytest.mark.gpu
@pytest.mark.parametrize("logn", [9, 10])
def test_r1cs_verdict_only_boolean_rows(contexts, circuits, oracle, logn):
    """The verdict-only path decides a Boolean constraint (One - x) * x = 0 from the class byte of x alone when the
    constant column is One (r1cs_stream_kernel).  First violated row against the oracle for: a Boolean witness that is
    not a bit, the same with z[0] = 2 (so that x = z[0] = 2 satisfies its own row and every other Boolean row breaks), z[0]
    alone not One, a bit flipped to the other bit, and untouched neighbours; with and without the output buffers."""
    ctx, c = contexts(logn), circuits(logn, 0)
    n = 72
    sig, pk, hm = synth.make_signatures(logn, n, seed=29)
    z, st = ctx.witness_batch(sig, pk, hm)
    assert (st == 0).all()
    N, ni = 1 << logn, c.n_inst
    one = z[0, 0].copy()
    two = _fr_small(oracle, 2)
    # Boolean witnesses of the range proofs: the entries of a valid assignment that are 0 or One, far from the start
    bits = [j for j in range(ni + 2 * N + 40, ni + 2 * N + 400) if (not z[0, j].any()) or (z[0, j] == one).all()]
    assert len(bits) > 100
    z[3, bits[5]] = two
    z[17, bits[50]] = two
    z[17, 0] = two
    z[18, 0] = two
    z[40, bits[77]] = one if not z[40, bits[77]].any() else np.zeros(4, np.uint64)
    z[71, bits[99]] = np.array([0x123456789ABCDEF, 7, 9, 11], np.uint64)
    az, bz, cz, fu = ctx.r1cs_eval_batch(z)
    _, _, _, fu2 = ctx.r1cs_eval_batch(z, want=False)
    assert (fu2 == fu).all()
    for i in (0, 2, 3, 4, 16, 17, 18, 19, 40, 41, 70, 71):
        _, _, _, ofu = c.r1cs_eval(z[i])
        assert fu2[i] == ofu, (i, fu2[i], ofu)
        assert (ofu >= 0) == (i in (3, 17, 18, 40, 71)), i
