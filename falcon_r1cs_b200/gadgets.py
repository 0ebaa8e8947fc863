"""Gadget entry points of the reference under their own names (falcon-r1cs/src/gadgets/mod.rs:7-11), over the
C ABI's frcs_gadget_* exports (csrc/gadgets.cu): batched CUDA kernels that produce a gadget's witnesses in
arkworks' allocation order and evaluate the gadget's own rows.

In the reference a gadget takes `cs` and variables, allocates witnesses and enforces rows; the gadget tests then
assert `cs.is_satisfied()` and the output value.  Here a call takes operand *values* (ints) and returns a
GadgetResult: the output value(s), the witness block, and `satisfied` (= cs.is_satisfied() of a constraint system
holding only this gadget, as in the reference's #[cfg(test)] builds where range panics are compiled out) plus the
status a non-test build would panic with.  Nothing here computes on the CPU beyond int <-> Montgomery conversion.
"""
import ctypes as C

import numpy as np

from . import lib as L
from . import synth

Q = 12289
R_MOD = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
_R = (1 << 256) % R_MOD
_RINV = pow(1 << 256, -1, R_MOD)
L2_BOUND = {9: 34034726, 10: 70265242}  # gadgets/range_proofs.rs:104,196
MOD_Q, ADD_MOD, LESS_THAN_Q, LESS_THAN_6144, NORM_BOUND, NTT_CIRCUIT = range(6)


def to_int(fr):
    """canonical integers of Montgomery-form Fr images (..., 4) u64"""
    a = np.asarray(fr, dtype=np.uint64).reshape(-1, 4)
    return [(int(x[0]) | int(x[1]) << 64 | int(x[2]) << 128 | int(x[3]) << 192) * _RINV % R_MOD for x in a]


def to_fr(vals):
    """Montgomery images (n, 4) u64 of integers (F::from(x))"""
    out = np.zeros((len(vals), 4), dtype=np.uint64)
    for i, v in enumerate(vals):
        m = int(v) % R_MOD * _R % R_MOD
        for j in range(4):
            out[i, j] = (m >> (64 * j)) & 0xFFFFFFFFFFFFFFFF
    return out


class GadgetResult:
    def __init__(self, wit, first_unsat, status, out=None):
        self.wit = wit                      # (n, n_witness, 4) uint64, Montgomery
        self.first_unsat = first_unsat      # (n,) int64: first violated row of the gadget, -1 = none
        self.status = status                # (n,) int32: FRCS_OK / FRCS_E_COEFF_RANGE / FRCS_E_NORM_BOUND
        self.out = out                      # output values (ints) where the gadget has one

    @property
    def satisfied(self):
        return self.first_unsat == -1

    def wit_ints(self, i=0):
        return to_int(self.wit[i])


def ntt_param_var(logn):
    """the N twiddles 7^bitrev10(i) mod q of `ntt_param_var` (gadgets/misc.rs:67-77)"""
    return synth.ntt_table(1 << logn)


def shape(ctx, gadget):
    a, b, c = C.c_uint32(), C.c_uint32(), C.c_uint32()
    L.check(ctx._lib.frcs_gadget_shape(ctx.h, gadget, C.byref(a), C.byref(b), C.byref(c)), "frcs_gadget_shape")
    return a.value, b.value, c.value


def _p(a, t=L.u64p):
    return a.ctypes.data_as(t) if a is not None else None


def _scalar(ctx, gadget, fn, operands, expected=None, extra=()):
    ops = to_fr([x for row in operands for x in row])
    n = len(operands)
    _, n_wit, _ = shape(ctx, gadget)
    exp = to_fr(expected) if expected is not None else None
    wit = np.zeros((n, n_wit + (1 if exp is not None else 0), 4), dtype=np.uint64)
    fu = np.zeros(n, dtype=np.int64)
    st = np.zeros(n, dtype=np.int32)
    args = [ctx.h, n, _p(ops)] + ([_p(exp)] if fn in ("frcs_gadget_mod_q", "frcs_gadget_add_mod") else []) + list(extra)
    L.check(getattr(ctx._lib, fn)(*args, _p(wit), _p(fu, L.i64p), _p(st, L.i32p)), fn)
    return wit, fu, st


def mod_q(ctx, a, expected=None):
    """`mod_q(cs, &a, q)` (gadgets/arithmetics.rs:105-149) on integers a[i] < r; expected[i] (optional): the test
    macro's `b_var.enforce_equal(expected)` (arithmetics.rs:322-324).  out = a mod q."""
    wit, fu, st = _scalar(ctx, MOD_Q, "frcs_gadget_mod_q", [[x] for x in a], expected)
    return GadgetResult(wit, fu, st, to_int(wit[:, 1]))


def add_mod(ctx, a, b, expected=None):
    """`add_mod(cs, &a, &b, q)` (gadgets/arithmetics.rs:214-262); out = (a + b) mod q."""
    wit, fu, st = _scalar(ctx, ADD_MOD, "frcs_gadget_add_mod", list(zip(a, b)), expected)
    return GadgetResult(wit, fu, st, to_int(wit[:, 1]))


def enforce_less_than_q(ctx, a):
    """`enforce_less_than_q(cs, &a)` (gadgets/range_proofs.rs:42-94): satisfied iff a < 12289."""
    return GadgetResult(*_scalar(ctx, LESS_THAN_Q, "frcs_gadget_less_than_q", [[x] for x in a]))


def is_less_than_6144(ctx, a, enforce_true=False):
    """`is_less_than_6144(cs, &a)` (gadgets/range_proofs.rs:289-333); out = the returned Boolean's value;
    enforce_true adds `.enforce_equal(&Boolean::TRUE)` as the reference's test does (range_proofs.rs:512-513)."""
    wit, fu, st = _scalar(ctx, LESS_THAN_6144, "frcs_gadget_less_than_6144", [[x] for x in a], None,
                          (C.c_int32(int(enforce_true)),))
    return GadgetResult(wit, fu, st, to_int(wit[:, 15]))


def enforce_less_than_norm_bound(ctx, a):
    """`enforce_less_than_norm_bound(cs, &a)` (gadgets/range_proofs.rs:274-284) for the context's parameter set:
    satisfied iff a < SIG_L2_BOUND; status -17 where a non-test build panics (:114-117, :205-208)."""
    return GadgetResult(*_scalar(ctx, NORM_BOUND, "frcs_gadget_norm_bound", [[x] for x in a]))


class NTTPolyVar:
    """`NTTPolyVar::ntt_circuit(cs, &PolyVar, const_vars, param)` (gadgets/poly.rs:104-159)"""

    @staticmethod
    def ntt_circuit(ctx, polys, want_wit=True):
        """polys: (n, N) coefficients in [0, q).  out = (n, N) uint16 NTT values; wit (n, 29N, 4): per output the
        mod_q quotient t, the remainder b and the 27 range witnesses of b."""
        polys = np.ascontiguousarray(polys, dtype=np.uint16).reshape(-1, ctx.n)
        n = polys.shape[0]
        vals = np.zeros((n, ctx.n), dtype=np.uint16)
        wit = np.zeros((n, 29 * ctx.n, 4), dtype=np.uint64) if want_wit else None
        fu = np.zeros(n, dtype=np.int64)
        st = np.zeros(n, dtype=np.int32)
        L.check(ctx._lib.frcs_gadget_ntt_circuit(ctx.h, n, _p(polys, L.u16p), _p(vals, L.u16p), _p(wit), _p(fu, L.i64p),
                                                 _p(st, L.i32p)), "frcs_gadget_ntt_circuit")
        return GadgetResult(wit, fu, st, vals)


def l2_norm_var(ctx, v, sig):
    """`l2_norm_var(cs, &(v ++ sig), q)` (gadgets/misc.rs:30-51): sum over both polynomials of min(e, q - e)^2, read
    from the assignment of the full statement (the l2 elements are its last 36N witnesses before the norm bits);
    pk = 1 so that v = hm - sig."""
    n = ctx.n
    sig, v = np.asarray(sig, dtype=np.uint16).reshape(n), np.asarray(v, dtype=np.uint16).reshape(n)
    one = np.zeros(n, dtype=np.uint16)
    one[0] = 1
    hm = ((v.astype(np.uint32) + sig) % Q).astype(np.uint16)
    z, _ = ctx.witness_batch(sig[None], one[None], hm[None])
    w_l2 = ctx.n_inst + (2 + 27 + 58 + 30) * n
    blk = z[0, w_l2: w_l2 + 36 * n].reshape(2 * n, 18, 4)
    return sum(to_int(blk[:, 17]))  # the 18th witness of an element is its square p (App. A.5)
