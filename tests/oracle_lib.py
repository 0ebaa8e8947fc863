"""ctypes loader for the CPU oracle (oracle/liboracle.so).  Test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = None

u64p = C.POINTER(C.c_uint64)
u32p = C.POINTER(C.c_uint32)
u16p = C.POINTER(C.c_uint16)
u8p = C.POINTER(C.c_uint8)


def build():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(ROOT, "oracle", "liboracle.so")
        if not os.path.exists(path):
            build()
        _LIB = _bind(path)
    return _LIB


def _bind(path):
    L = C.CDLL(path)
    L.orc_circuit_new.restype = C.c_void_p
    L.orc_setup.restype = C.c_void_p
    L.orc_pk_len.restype = C.c_uint64
    return L


def load_native():
    """bench.py's CPU arm: rebuild the oracle on this box with -march=native into oracle/_native/ and use that copy;
    returns a description of the build in use.  Must be called before the first lib()."""
    global _LIB
    flags = "-O3 -march=native -std=c++17 -fopenmp -fPIC -Wall -Wno-unused-function"
    out = os.path.join(ROOT, "oracle", "_native", "liboracle.so")
    try:
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.check_call(["make", "-s", "-B", "-C", os.path.join(ROOT, "oracle"), "native"], stdout=subprocess.DEVNULL,
                              stderr=subprocess.DEVNULL, timeout=300)
        _LIB = _bind(out)
        return "g++ " + flags + " (built on this box)"
    except Exception:
        lib()
        return "g++ -O3 -march=x86-64-v3 (shipped build; no compiler on this box)"


def ptr(a, t=u64p):
    return a.ctypes.data_as(t)


class Circuit:
    def __init__(self, logn, kind=0):
        self.logn, self.kind, self.n = logn, kind, 1 << logn
        self.h = C.c_void_p(lib().orc_circuit_new(logn, kind))
        assert self.h.value
        s = np.zeros(8, dtype=np.uint64)
        lib().orc_shape(self.h, ptr(s))
        (self.n_inst, self.n_wit, self.n_cons, self.nnz_a, self.nnz_b, self.nnz_c, self.domain_log2) = [int(x) for x in s[:7]]
        self.n_z = self.n_inst + self.n_wit

    def __del__(self):
        if getattr(self, "h", None) and self.h.value:
            lib().orc_circuit_free(self.h)
            self.h = None

    def csr(self, which):
        nnz = [self.nnz_a, self.nnz_b, self.nnz_c][which]
        rp = np.zeros(self.n_cons + 1, dtype=np.uint32)
        col = np.zeros(nnz, dtype=np.uint32)
        val = np.zeros((nnz, 4), dtype=np.uint64)
        lib().orc_get_csr(self.h, which, ptr(rp, u32p), ptr(col, u32p), ptr(val))
        return rp, col, val

    def witness(self, sig, pk, hm, construct_matrices=False, panic_on_range=True):
        z = np.zeros((self.n_z, 4), dtype=np.uint64)
        fu = C.c_int64(-2)
        sig, pk, hm = [np.ascontiguousarray(x, dtype=np.uint16) for x in (sig, pk, hm)]
        st = lib().orc_witness(self.h, ptr(sig, u16p), ptr(pk, u16p), ptr(hm, u16p), ptr(z),
                               int(construct_matrices), int(panic_on_range), C.byref(fu))
        return z, st, fu.value

    def r1cs_eval(self, z):
        az = np.zeros((self.n_cons, 4), dtype=np.uint64)
        bz = np.zeros_like(az)
        cz = np.zeros_like(az)
        fu = C.c_int64(0)
        z = np.ascontiguousarray(z, dtype=np.uint64)
        lib().orc_r1cs_eval(self.h, ptr(z), ptr(az), ptr(bz), ptr(cz), C.byref(fu))
        return az, bz, cz, fu.value

    def witness_map(self, z):
        h = np.zeros((1 << self.domain_log2, 4), dtype=np.uint64)
        z = np.ascontiguousarray(z, dtype=np.uint64)
        lib().orc_witness_map(self.h, ptr(z), ptr(h))
        return h

    def setup(self, seed=42, trapdoor=None):
        return Pk(self, seed, trapdoor)

    def prove(self, pk, z, r, s):
        pa = np.zeros(48, dtype=np.uint64)
        comp = np.zeros(192, dtype=np.uint8)
        z = np.ascontiguousarray(z, dtype=np.uint64)
        lib().orc_prove(self.h, pk.h, ptr(z), ptr(r), ptr(s), ptr(pa), ptr(comp, u8p))
        return pa, comp

    def verify_trapdoor(self, pk, z, r, s, proof_affine):
        z = np.ascontiguousarray(z, dtype=np.uint64)
        pa = np.ascontiguousarray(proof_affine, dtype=np.uint64)
        return bool(lib().orc_verify_trapdoor(self.h, pk.h, ptr(z), ptr(r), ptr(s), ptr(pa)))


PK_NAMES = ["a_query", "b_g1_query", "b_g2_query", "h_query", "l_query", "gamma_abc_g1", "g1_elems", "g2_elems"]


class Pk:
    def __init__(self, circ, seed, trapdoor=None):
        if trapdoor is not None:
            td = np.ascontiguousarray(trapdoor, dtype=np.uint64).reshape(7, 4)
            lib().orc_setup_trapdoor.restype = C.c_void_p
            self.h = C.c_void_p(lib().orc_setup_trapdoor(circ.h, ptr(td)))
        else:
            self.h = C.c_void_p(lib().orc_setup(circ.h, C.c_uint64(seed)))
        self._cache = {}

    def __del__(self):
        if getattr(self, "h", None) and self.h.value:
            lib().orc_pk_free(self.h)
            self.h = None

    def trapdoor(self):
        """alpha, beta, gamma, delta, tau, g1_scalar, g2_scalar (Montgomery limbs, 7 x 4)"""
        out = np.zeros((7, 4), dtype=np.uint64)
        lib().orc_pk_trapdoor(self.h, ptr(out))
        return out

    def export(self, name):
        if name not in self._cache:
            which = PK_NAMES.index(name)
            n = int(lib().orc_pk_len(self.h, which))
            w = 24 if which in (2, 7) else 12
            out = np.zeros((n, w), dtype=np.uint64)
            lib().orc_pk_export(self.h, which, ptr(out))
            self._cache[name] = out
        return self._cache[name]


def fr_from_canonical(x):
    x = np.ascontiguousarray(x, dtype=np.uint64).reshape(-1, 4)
    out = np.zeros_like(x)
    lib().orc_fr_from_canonical(ptr(x), ptr(out), C.c_uint64(x.shape[0]))
    return out


def fr_to_canonical(x):
    x = np.ascontiguousarray(x, dtype=np.uint64).reshape(-1, 4)
    out = np.zeros_like(x)
    lib().orc_fr_to_canonical(ptr(x), ptr(out), C.c_uint64(x.shape[0]))
    return out


def ints_to_limbs(vals, nl=4):
    out = np.zeros((len(vals), nl), dtype=np.uint64)
    for i, v in enumerate(vals):
        for j in range(nl):
            out[i, j] = (int(v) >> (64 * j)) & 0xFFFFFFFFFFFFFFFF
    return out


def limbs_to_ints(a):
    a = np.asarray(a, dtype=np.uint64).reshape(-1, a.shape[-1])
    return [sum(int(a[i, j]) << (64 * j) for j in range(a.shape[1])) for i in range(a.shape[0])]


def msm_g1(bases, scalars_canonical):
    out = np.zeros(12, dtype=np.uint64)
    b = np.ascontiguousarray(bases, dtype=np.uint64)
    s = np.ascontiguousarray(scalars_canonical, dtype=np.uint64)
    lib().orc_msm_g1(ptr(b), ptr(s), C.c_uint64(b.shape[0]), ptr(out))
    return out


def msm_g2(bases, scalars_canonical):
    out = np.zeros(24, dtype=np.uint64)
    b = np.ascontiguousarray(bases, dtype=np.uint64)
    s = np.ascontiguousarray(scalars_canonical, dtype=np.uint64)
    lib().orc_msm_g2(ptr(b), ptr(s), C.c_uint64(b.shape[0]), ptr(out))
    return out


def kat(which, logn, inputs, expected=0):
    a = np.array(inputs, dtype=np.uint64)
    ok = C.c_int(-1)
    counts = np.zeros(3, dtype=np.uint64)
    sat = lib().orc_kat(which, logn, ptr(a), len(inputs), C.c_uint64(expected), C.byref(ok), ptr(counts))
    return bool(sat), ok.value, [int(c) for c in counts]


def kat_z(which, logn, inputs, expected=0, n_wit=64):
    """orc_kat plus the witness assignment (n_wit x 4 Montgomery limbs, the gadget's operands first) and the first
    violated row (-1: satisfied)"""
    a = np.array(inputs, dtype=np.uint64)
    ok = C.c_int(-1)
    counts = np.zeros(3, dtype=np.uint64)
    z = np.zeros((n_wit, 4), dtype=np.uint64)
    fu = C.c_int64(-2)
    sat = lib().orc_kat_z(which, logn, ptr(a), len(inputs), C.c_uint64(expected), C.byref(ok), ptr(counts), ptr(z),
                          C.byref(fu))
    nw = int(counts[1])
    assert nw <= n_wit
    return bool(sat), z[:nw], fu.value


def kat_ntt_z(logn, poly):
    n = 1 << logn
    poly = np.ascontiguousarray(poly, dtype=np.uint16)
    out = np.zeros(n, dtype=np.uint16)
    counts = np.zeros(3, dtype=np.uint64)
    z = np.zeros((29 * n, 4), dtype=np.uint64)
    sat = lib().orc_kat_ntt_z(logn, ptr(poly, u16p), ptr(out, u16p), ptr(counts), ptr(z))
    return bool(sat), out, z, [int(c) for c in counts]
