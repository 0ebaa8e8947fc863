"""CPU test: the product's closed-form circuit compiler (csrc/circuit.hpp, through the host-only hook
frcs_debug_host_matrix) against the oracle's generic arkworks-style synthesis, entry by entry, for the three
circuits (circuits/falcon_ntt.rs, falcon_schoolbook.rs, falcon_dual_ntt.rs).  No GPU needed."""
import numpy as np
import pytest

from falcon_r1cs_b200 import lib as L


def host_matrix(logn, kind, which):
    lib = L.load()
    cnt = np.zeros(4, np.uint64)
    assert lib.frcs_debug_host_matrix(logn, kind, which, None, None, None, cnt.ctypes.data_as(L.u64p)) == 0
    rp = np.zeros(int(cnt[2]) + 1, np.uint32)
    col = np.zeros(int(cnt[3]), np.uint32)
    val = np.zeros((int(cnt[3]), 4), np.uint64)
    assert lib.frcs_debug_host_matrix(logn, kind, which, rp.ctypes.data_as(L.u32p), col.ctypes.data_as(L.u32p),
                                      val.ctypes.data_as(L.u64p), cnt.ctypes.data_as(L.u64p)) == 0
    return [int(x) for x in cnt], rp, col, val


@pytest.mark.parametrize("logn,kind", [(9, 0), (10, 0), (9, 1), (9, 2), (10, 2)])
def test_compiler_matrices_equal_oracle(circuits, oracle, logn, kind):
    c = circuits(logn, kind)
    for which in range(3):
        cnt, rp, col, val = host_matrix(logn, kind, which)
        orp, ocol, oval = c.csr(which)
        assert cnt[:3] == [c.n_inst, c.n_wit, c.n_cons]
        assert (rp == orp).all() and (col == ocol).all()
        assert (val == oracle.fr_to_canonical(oval)).all()


def test_dual_ntt_counts(circuits):
    """no README row exists for the dual circuit; SURVEY.md section 8f.3 predicts ~193.6 k constraints at N = 1024:
    186 N + 4 + norm witnesses, 189 N + 8 + norm rows (2 pairs of N + 4 rows, 4 x 30 N, 63 N, 4 N)"""
    for logn, nw, nr in ((9, 50, 52), (10, 52, 54)):
        n = 1 << logn
        c = circuits(logn, 2)
        assert (c.n_inst, c.n_wit, c.n_cons) == (1 + 2 * n, 186 * n + 4 + nw, 189 * n + 8 + nr)
    assert circuits(10, 2).n_cons == 193598 and circuits(10, 2).domain_log2 == 18


def test_dual_ntt_oracle_satisfied_and_not(circuits):
    """circuits/falcon_dual_ntt.rs:141-169: a valid signature satisfies the system; corrupted ones do not"""
    from falcon_r1cs_b200 import synth
    for logn in (9, 10):
        c = circuits(logn, 2)
        sig, pk, hm = synth.make_signatures(logn, 2, seed=71)
        z, st, fu = c.witness(sig[0], pk[0], hm[0], construct_matrices=True)
        assert st == 0 and fu == -1 and c.r1cs_eval(z)[3] == -1
        bad = z.copy()
        bad[c.n_inst + 3] = bad[0]  # sig.pos[3] := 1
        assert c.r1cs_eval(bad)[3] >= 0
        hm2 = hm[1].copy()
        hm2[0] = (int(hm2[0]) + 1) % 12289  # v changes by one coefficient: still a consistent witness, larger or equal norm
        z2, st2, fu2 = c.witness(sig[1], pk[1], hm2, construct_matrices=True)
        assert fu2 == -1 or st2 != 0
        big = np.full(1 << logn, 6000, np.uint16)
        _, st3, fu3 = c.witness(big, pk[1], hm[1], construct_matrices=True)
        assert st3 == -2 and fu3 >= 0  # norm bound: the reference panics (range_proofs.rs:114-117 / 205-208)
