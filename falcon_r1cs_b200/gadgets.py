"""Gadget entry points of the reference under their own names (falcon-r1cs/src/gadgets/mod.rs:7-11), as views of the
assignment the CUDA witness path produces.

In the reference a gadget takes `cs` and variables, allocates its witnesses and enforces its rows; here the whole
statement is one fused kernel (`witness_kernel`, csrc/witness.cu) writing z in arkworks' allocation order, so a gadget
call = run the kernel on inputs that carry the gadget's operands and return the gadget's own slice of z (outputs,
witnesses in allocation order) together with the satisfaction status of its rows.  Nothing here computes on the CPU
beyond slicing and converting the Montgomery images to integers.

Layout (witness index w -> z column 1 + 2N + w):
    sig[N] | v[N] | 27N range(v) | 29N ntt_circuit(sig) | 29N ntt_circuit(v) | 30N pointwise | 36N l2 | norm bits, chain
"""
import numpy as np

from . import synth

Q = 12289
R_MOD = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
_RINV = pow(1 << 256, -1, R_MOD)
L2_BOUND = {9: 34034726, 10: 70265242}  # gadgets/range_proofs.rs:104,196


class Layout:
    def __init__(self, logn):
        n = 1 << logn
        self.logn, self.n, self.n_inst = logn, n, 1 + 2 * n
        self.w_sig, self.w_v, self.w_vrange = 0, n, 2 * n
        self.w_nttsig = self.w_vrange + 27 * n
        self.w_nttv = self.w_nttsig + 29 * n
        self.w_pw = self.w_nttv + 29 * n
        self.w_l2 = self.w_pw + 30 * n
        self.w_norm = self.w_l2 + 36 * n
        self.r_vrange, self.r_nttsig = 0, 29 * n
        self.r_nttv = self.r_nttsig + 30 * n
        self.r_pw = self.r_nttv + 30 * n
        self.r_l2 = self.r_pw + 32 * n
        self.r_norm = self.r_l2 + 38 * n

    def col(self, w):
        return self.n_inst + w


def to_int(fr):
    """canonical integers of Montgomery-form Fr images (..., 4) u64"""
    a = np.asarray(fr, dtype=np.uint64).reshape(-1, 4)
    return [(int(x[0]) | int(x[1]) << 64 | int(x[2]) << 128 | int(x[3]) << 192) * _RINV % R_MOD for x in a]


def ntt_param_var(logn):
    """the N twiddles 7^bitrev10(i) mod q of `ntt_param_var` (gadgets/misc.rs:67-77)"""
    return synth.ntt_table(1 << logn)


def _assignment(ctx, sig, v):
    """z for a statement whose signature polynomial is `sig` and whose v = hm - sig * pk is `v` (pk = 1)"""
    n = ctx.n
    sig, v = np.asarray(sig, dtype=np.uint16).reshape(n), np.asarray(v, dtype=np.uint16).reshape(n)
    one = np.zeros(n, dtype=np.uint16)
    one[0] = 1  # pk(x) = 1, so hm = v + sig
    hm = ((v.astype(np.uint32) + sig) % Q).astype(np.uint16)
    z, st = ctx.witness_batch(sig[None], one[None], hm[None])  # (the kernel takes pk and hm as coefficient vectors)
    return z[0], int(st[0])


class NTTPolyVar:
    """`NTTPolyVar::ntt_circuit(cs, &PolyVar, const_vars, param)` (gadgets/poly.rs:104-159)"""

    @staticmethod
    def ntt_circuit(ctx, poly):
        """NTT of `poly` (N coefficients in [0, q)) through the circuit's lazy butterflies and mod_q reductions.
        Returns (values[N], witnesses (N, 29) as integers: t, b and the 27 range witnesses of every output)."""
        lay = Layout(ctx.logn)
        z, _ = _assignment(ctx, poly, np.zeros(ctx.n, np.uint16))
        blk = z[lay.col(lay.w_nttsig): lay.col(lay.w_nttsig) + 29 * ctx.n]
        wit = np.array(to_int(blk), dtype=object).reshape(ctx.n, 29)
        return [int(x) for x in wit[:, 1]], wit


def mod_q(ctx, poly):
    """`mod_q(cs, &a, q)` (gadgets/arithmetics.rs:105-149) on the N unreduced butterfly outputs a_k of `poly`:
    returns [(t_k, b_k)] with a_k = q t_k + b_k, 0 <= b_k < q"""
    _, wit = NTTPolyVar.ntt_circuit(ctx, poly)
    return [(int(w[0]), int(w[1])) for w in wit]


def l2_norm_var(ctx, v, sig):
    """`l2_norm_var(cs, &(v ++ sig), q)` (gadgets/misc.rs:30-51): sum over both polynomials of min(e, q - e)^2"""
    lay = Layout(ctx.logn)
    z, _ = _assignment(ctx, sig, v)
    blk = z[lay.col(lay.w_l2): lay.col(lay.w_l2) + 36 * ctx.n].reshape(2 * ctx.n, 18, 4)
    return sum(to_int(blk[:, 17]))  # the 18th witness of an element is its square p (App. A.5)


def enforce_less_than_norm_bound(ctx, v, sig):
    """`enforce_less_than_norm_bound(cs, &norm)` (gadgets/range_proofs.rs:274-284).  The reference panics outside
    tests when the bound fails (`:114-117,205-208`); here the kernel reports FRCS_E_NORM_BOUND (-17): returns True
    iff the bound holds."""
    _, st = _assignment(ctx, sig, v)
    return st != -17
