"""CPU checks of the gadget layer, no GPU needed: the stand-alone gadget circuits the CUDA entry points evaluate
(circuit::Builder::build_gadget through the host-only hook frcs_debug_host_matrix) hold on the oracle's assignment
exactly where the oracle's generic synthesis says they do -- same witness count, same row count, same first violated
row on the reference's bad-path cases (gadgets/arithmetics.rs:346-361, 480-494; range_proofs.rs:365-389, 442-474,
529-547)."""
import json
import os

import numpy as np
import pytest

from falcon_r1cs_b200 import gadgets as G
from test_host_matrices import host_matrix

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
KATS = json.load(open(os.path.join(GOLD, "gadget_kats.json")))
R = G.R_MOD


def first_unsat(mats, z):
    """cs.which_is_unsatisfied() of CSR matrices (canonical coefficients) on the integer assignment z"""
    def dot(m, r):
        _, rp, col, val = m
        s = 0
        for e in range(rp[r], rp[r + 1]):
            c = sum(int(val[e, j]) << (64 * j) for j in range(4))
            s += c * z[col[e]]
        return s % R
    for r in range(mats[0][0][2]):
        if dot(mats[0], r) * dot(mats[1], r) % R != dot(mats[2], r):
            return r
    return -1


def test_fr_conversions_round_trip():
    vals = [0, 1, 12289, (1 << 200) + 5, R - 1]
    assert G.to_int(G.to_fr(vals)) == vals


@pytest.mark.parametrize("gadget,which,name,logn", [(G.MOD_Q, 0, "mod_q", 10), (G.ADD_MOD, 1, "add_mod", 10),
                                                     (G.LESS_THAN_Q, 3, "less_than_q", 10),
                                                     (G.LESS_THAN_6144, 5, "less_than_6144", 9),
                                                     (G.NORM_BOUND, 4, "norm_bound_512", 9),
                                                     (G.NORM_BOUND, 4, "norm_bound_1024", 10)])
def test_gadget_circuits_on_the_oracle_assignment(oracle, gadget, which, name, logn):
    with_exp = gadget in (G.MOD_Q, G.ADD_MOD, G.LESS_THAN_6144)
    mats = [host_matrix(logn, 16 + gadget + (8 if with_exp else 0), w) for w in range(3)]
    for row in KATS[name]:
        sat, ins = row[-1], row[:-1]
        exp = 0
        if gadget in (G.MOD_Q, G.ADD_MOD):
            exp, ins = ins[-1], ins[:-1]
        elif gadget == G.LESS_THAN_6144:
            exp = int(sat)
        osat, oz, ofu = oracle.kat_z(which, logn, ins, exp)
        assert osat == sat
        assert mats[0][0][:3] == [1, oz.shape[0], mats[0][0][2]]  # same witness count as the oracle's system
        z = [1] + G.to_int(oz)
        assert first_unsat(mats, z) == ofu, (name, row)
