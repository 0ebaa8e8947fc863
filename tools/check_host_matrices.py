#!/usr/bin/env python
"""CPU-only: the circuit compiler's matrices (frcs_debug_host_matrix) against the oracle's generic synthesis."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_lib as O
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


if __name__ == "__main__":
    kinds = [int(x) for x in sys.argv[1:]] or [0, 2]
    for logn in (9, 10):
        for kind in kinds:
            c = O.Circuit(logn, kind)
            for which in range(3):
                cnt, rp, col, val = host_matrix(logn, kind, which)
                orp, ocol, oval = c.csr(which)
                ok = cnt[:3] == [c.n_inst, c.n_wit, c.n_cons] and rp.shape == orp.shape and (rp == orp).all() and \
                    (col == ocol).all() and (val == O.fr_to_canonical(oval)).all()
                print(logn, kind, which, ok, cnt)
