"""CPU check of the gadget views (falcon_r1cs_b200/gadgets.py): the slices it takes from an assignment hold what the
reference's gadgets allocate there, verified on the oracle's assignment -- no GPU needed."""
import numpy as np
import pytest

from falcon_r1cs_b200 import gadgets as G, synth


@pytest.mark.parametrize("logn", [9, 10])
def test_gadget_slices_of_the_oracle_assignment(circuits, logn):
    c = circuits(logn, 0)
    n = 1 << logn
    lay = G.Layout(logn)
    assert lay.n_inst == c.n_inst and lay.w_norm < c.n_wit
    rng = np.random.default_rng(5 + logn)
    poly = rng.integers(0, G.Q, n).astype(np.uint16)
    v = rng.integers(0, 200, n).astype(np.uint16)
    one = np.zeros(n, np.uint16)
    one[0] = 1
    hm = ((v.astype(np.uint32) + poly) % G.Q).astype(np.uint16)
    z, _, _ = c.witness(poly, one, hm, panic_on_range=False)
    # sig and v where the circuit allocates them (falcon_ntt.rs:53-71)
    assert G.to_int(z[lay.col(lay.w_sig): lay.col(lay.w_sig) + n]) == [int(x) for x in poly]
    assert G.to_int(z[lay.col(lay.w_v): lay.col(lay.w_v) + n]) == [int(x) for x in v]
    # ntt_circuit(sig): per output (t, b, 27 range witnesses); b = clear-text NTT (gadgets/poly.rs:292-297)
    wit = np.array(G.to_int(z[lay.col(lay.w_nttsig): lay.col(lay.w_nttsig) + 29 * n]), dtype=object).reshape(n, 29)
    assert [int(x) for x in wit[:, 1]] == [int(x) for x in synth.ntt(poly, logn)]
    for k in (0, 7, n - 1):
        b = int(wit[k, 1])
        assert [int(wit[k, 2 + j]) for j in range(14)] == [(b >> j) & 1 for j in range(14)]
    # l2_norm_var: the 18th witness of each of the 2N elements is its square (gadgets/misc.rs:30-51)
    l2 = z[lay.col(lay.w_l2): lay.col(lay.w_l2) + 36 * n].reshape(2 * n, 18, 4)
    cent = lambda e: min(int(e), G.Q - int(e))
    assert sum(G.to_int(l2[:, 17])) == sum(cent(e) ** 2 for e in list(v) + list(poly))
