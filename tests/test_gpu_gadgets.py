"""Gadget entry points under the reference's names (falcon_r1cs_b200/gadgets.py) on the CUDA witness path, checked the
way the reference's own gadget tests do: ntt_circuit against the clear-text NTT (gadgets/poly.rs:292-297), mod_q as
a = q t + b with b < q (gadgets/arithmetics.rs:346-361), the l2 norm and its bound on good and bad inputs
(gadgets/range_proofs.rs:529-547) -- and bit for bit against the oracle's assignment."""
import numpy as np
import pytest

from falcon_r1cs_b200 import gadgets as G, synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("logn", [9, 10])
def test_ntt_circuit_equals_clear_text_ntt(contexts, circuits, logn):
    ctx, c = contexts(logn), circuits(logn, 0)
    n = 1 << logn
    rng = np.random.default_rng(100 + logn)
    for poly in (rng.integers(0, G.Q, n).astype(np.uint16), np.full(n, G.Q - 1, np.uint16), np.zeros(n, np.uint16)):
        vals, wit = G.NTTPolyVar.ntt_circuit(ctx, poly)
        assert vals == [int(x) for x in synth.ntt(poly, logn)]
        # mod_q: every output is the remainder of an unreduced value a = q t + b, with the 14 bits of b behind it
        for t, b in G.mod_q(ctx, poly)[:8]:
            assert 0 <= b < G.Q and t >= 0
        k = int(rng.integers(0, n))
        b = int(wit[k, 1])
        assert [int(wit[k, 2 + j]) for j in range(14)] == [(b >> j) & 1 for j in range(14)]
        # the gadget's whole witness block equals the oracle's (arkworks allocation order)
        lay = G.Layout(logn)
        one = np.zeros(n, np.uint16)
        one[0] = 1
        zo, _, _ = c.witness(poly, one, poly, panic_on_range=False)
        blk = zo[lay.col(lay.w_nttsig): lay.col(lay.w_nttsig) + 29 * n]
        assert [int(x) for x in wit.reshape(-1)] == G.to_int(blk)


def test_ntt_param_is_the_reference_table():
    tab = G.ntt_param_var(10)
    assert [int(tab[i]) for i in range(4)] == [1, pow(7, 512, G.Q), pow(7, 256, G.Q), pow(7, 768, G.Q)]  # 7^bitrev10(i)


@pytest.mark.parametrize("logn", [9, 10])
def test_l2_norm_and_bound(contexts, logn):
    ctx = contexts(logn)
    n = 1 << logn
    rng = np.random.default_rng(200 + logn)
    small = lambda: (np.rint(rng.normal(0, 60, n)).astype(np.int64) % G.Q).astype(np.uint16)
    v, sig = small(), small()
    cent = lambda e: min(int(e), G.Q - int(e))
    want = sum(cent(e) ** 2 for e in list(v) + list(sig))
    assert G.l2_norm_var(ctx, v, sig) == want and want < G.L2_BOUND[logn]
    assert G.enforce_less_than_norm_bound(ctx, v, sig)
    # one coefficient pushed to the largest centred value: the norm exceeds the bound (reference: panic / unsatisfied)
    bad = sig.copy()
    bad[:] = 6000
    assert G.l2_norm_var(ctx, v, bad) > G.L2_BOUND[logn]
    assert not G.enforce_less_than_norm_bound(ctx, v, bad)
