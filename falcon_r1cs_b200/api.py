"""Host-side mirror of the reference's interface for the hot path, over the C ABI.

Reference interface (all Rust):
  FalconNTTVerificationCircuit::build_circuit(pk, msg, sig)      circuits/falcon_ntt.rs:15
  ConstraintSynthesizer::generate_constraints(self, cs)          circuits/falcon_ntt.rs:26
  ark_groth16::create_random_proof(circuit, &pk, rng)            examples/pok_sig.rs:32
  cs.is_satisfied(), cs.num_constraints(), cs.to_matrices()      circuits/falcon_ntt.rs:143-159

The same names are kept here; the arithmetic runs in libfalcon_r1cs_b200.so (CUDA,
sm_100a).  Nothing in this module computes on the CPU: it marshals numpy buffers.
"""
import ctypes as C
import hashlib

import numpy as np

from . import lib as L

Q = 12289


def _p(a, t=L.u64p):
    return a.ctypes.data_as(t) if a is not None else None


def _c(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


def hash_to_point(nonce: bytes, msg: bytes, n: int):
    """Polynomial::from_hash_of_message(msg, nonce) (falcon_ntt.rs:44, [EXT] falcon-rust):
    Falcon's hash-to-point, SHAKE256(nonce || msg) read as big-endian 16-bit words,
    rejection-sampled below 5*q = 61445 and reduced mod q.  Public data; done on the host."""
    sh = hashlib.shake_256(nonce + msg)
    out, need = [], n
    stream = sh.digest(2 * n * 2 + 512)
    pos = 0
    while len(out) < need:
        if pos + 2 > len(stream):
            stream = sh.digest(len(stream) * 2)
        w = (stream[pos] << 8) | stream[pos + 1]
        pos += 2
        if w < 61445:
            out.append(w % Q)
    return np.array(out, dtype=np.uint16)


class FalconNTTVerificationCircuit:
    """pk, sig: coefficient vectors in [0, q) (Polynomial::from(&PublicKey) /
    Polynomial::from(&Signature)); msg + nonce (or a precomputed hm polynomial)."""

    KIND = L.KIND_NTT

    def __init__(self, pk, msg, sig, nonce=b"", hm=None):
        self.pk = _c(pk, np.uint16)
        self.sig = _c(sig, np.uint16)
        self.msg = msg
        self.nonce = nonce
        self.hm = _c(hm, np.uint16) if hm is not None else hash_to_point(nonce, msg, self.pk.shape[-1])

    @classmethod
    def build_circuit(cls, pk, msg, sig, nonce=b"", hm=None):
        return cls(pk, msg, sig, nonce, hm)

    @classmethod
    def from_bytes(cls, pk_bytes, msg, sig_bytes):
        """build_circuit(pk, msg, sig) from Falcon wire formats (circuits/falcon_ntt.rs:15,27-28,44): the public key
        and the compressed signature are decoded, hm = hash_to_point(nonce || msg)."""
        from . import falcon_codec as fc
        logn_pk, h = fc.decode_public_key(pk_bytes)
        logn_sig, nonce, s2 = fc.decode_signature(sig_bytes)
        if logn_pk != logn_sig:
            raise fc.FalconFormatError("public key and signature are for different Falcon parameter sets")
        return cls(h, msg, s2, nonce)

    @property
    def logn(self):
        return int(self.pk.shape[-1]).bit_length() - 1


class FalconSchoolBookVerificationCircuit(FalconNTTVerificationCircuit):
    """circuits/falcon_schoolbook.rs:8-18: same (pk, msg, sig) statement, schoolbook product (context kind 1)."""
    KIND = L.KIND_SCHOOLBOOK


class FalconDualNTTVerificationCircuit(FalconNTTVerificationCircuit):
    """circuits/falcon_dual_ntt.rs:8-18: sig and v split into (pos, neg) halves (context kind 2)."""
    KIND = L.KIND_DUAL_NTT


class ProvingKey:
    """ark_groth16::ProvingKey<Bls12_381> as plain arrays (see frcs_pk_view)."""

    FIELDS = ["alpha_g1", "beta_g1", "delta_g1", "beta_g2", "delta_g2", "a_query", "b_g1_query", "b_g2_query",
              "h_query", "l_query"]

    def __init__(self, **kw):
        for f in self.FIELDS:
            setattr(self, f, _c(kw[f], np.uint64))

    def view(self):
        v = L.PkView()
        for f in ["alpha_g1", "beta_g1", "delta_g1", "beta_g2", "delta_g2"]:
            setattr(v, f, _p(getattr(self, f)))
        for f, ln in [("a_query", "a_len"), ("b_g1_query", "b_g1_len"), ("b_g2_query", "b_g2_len"),
                      ("h_query", "h_len"), ("l_query", "l_len")]:
            arr = getattr(self, f)
            setattr(v, f, _p(arr))
            setattr(v, ln, arr.shape[0])
        return v


class Context:
    """One (device, circuit) context: frcs_ctx."""

    def __init__(self, logn, kind=L.KIND_NTT, device=0):
        self._lib = L.load()
        h = C.c_void_p()
        L.check(self._lib.frcs_ctx_create(logn, kind, device, C.byref(h)), "frcs_ctx_create")
        self.h = h
        s = L.Shape()
        L.check(self._lib.frcs_shape_get(self.h, C.byref(s)), "frcs_shape_get")
        self.logn, self.kind, self.device = logn, kind, device
        self.n = 1 << logn
        self.n_inst, self.n_wit, self.n_cons = s.n_instance, s.n_witness, s.n_constraints
        self.n_z = self.n_inst + self.n_wit
        self.domain_log2 = s.domain_log2
        self.nnz = (s.nnz_a, s.nnz_b, s.nnz_c)
        self._pk = None

    def close(self):
        if self.h:
            self._lib.frcs_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- cs.to_matrices() ------------------------------------------------------------
    def get_matrix(self, which):
        rp = np.zeros(self.n_cons + 1, dtype=np.uint32)
        col = np.zeros(self.nnz[which], dtype=np.uint32)
        val = np.zeros((self.nnz[which], 4), dtype=np.uint64)
        L.check(self._lib.frcs_get_matrix(self.h, which, _p(rp, L.u32p), _p(col, L.u32p), _p(val)), "frcs_get_matrix")
        return rp, col, val

    # -- generate_constraints (Prove mode), batched --------------------------------
    def witness_batch(self, sig, pk, hm):
        sig, pk, hm = [_c(x, np.uint16).reshape(-1, self.n) for x in (sig, pk, hm)]
        n = sig.shape[0]
        z = np.empty((n, self.n_z, 4), dtype=np.uint64)
        st = np.zeros(n, dtype=np.int32)
        L.check(self._lib.frcs_witness_batch(self.h, n, _p(sig, L.u16p), _p(pk, L.u16p), _p(hm, L.u16p), _p(z),
                                             _p(st, L.i32p)), "frcs_witness_batch")
        return z, st

    def witness_check_batch(self, sig, pk, hm):
        """generate_constraints + cs.which_is_unsatisfied() for a batch; the assignments stay on the device.
        Returns (first_unsat, status): first_unsat[i] == -1 and status[i] == 0 for a valid signature."""
        sig, pk, hm = [_c(x, np.uint16).reshape(-1, self.n) for x in (sig, pk, hm)]
        n = sig.shape[0]
        fu = np.zeros(n, dtype=np.int64)
        st = np.zeros(n, dtype=np.int32)
        L.check(self._lib.frcs_witness_check_batch(self.h, n, _p(sig, L.u16p), _p(pk, L.u16p), _p(hm, L.u16p),
                                                   _p(fu, L.i64p), _p(st, L.i32p)), "frcs_witness_check_batch")
        return fu, st

    def generate_constraints(self, circuit):
        """Single-circuit form: returns (z, status)."""
        z, st = self.witness_batch(circuit.sig, circuit.pk, circuit.hm)
        return z[0], int(st[0])

    # -- A.z, B.z, C.z and cs.is_satisfied() -----------------------------------------
    def r1cs_eval_batch(self, z, want=True):
        z = _c(z, np.uint64).reshape(-1, self.n_z, 4)
        n = z.shape[0]
        az = np.empty((n, self.n_cons, 4), dtype=np.uint64) if want else None
        bz = np.empty_like(az) if want else None
        cz = np.empty_like(az) if want else None
        fu = np.zeros(n, dtype=np.int64)
        L.check(self._lib.frcs_r1cs_eval_batch(self.h, n, _p(z), _p(az), _p(bz), _p(cz), _p(fu, L.i64p)),
                "frcs_r1cs_eval_batch")
        return az, bz, cz, fu

    def is_satisfied(self, z):
        return bool((self.r1cs_eval_batch(z, want=False)[3] == -1).all())

    # -- R1CStoQAP::witness_map ----------------------------------------------------------
    def witness_map(self, z):
        z = _c(z, np.uint64).reshape(self.n_z, 4)
        h = np.empty((1 << self.domain_log2, 4), dtype=np.uint64)
        L.check(self._lib.frcs_witness_map(self.h, _p(z), _p(h)), "frcs_witness_map")
        return h

    def domain_op(self, log_size, op, data):
        data = _c(data, np.uint64).reshape(1 << log_size, 4).copy()
        L.check(self._lib.frcs_domain_op(self.h, log_size, op, _p(data)), "frcs_domain_op")
        return data

    # -- VariableBaseMSM::multi_scalar_mul ---------------------------------------------
    def msm_g1(self, bases, scalars, window_bits=None):
        """window_bits: None = frcs_msm_g1; 16 / 8 = the same sum through that window geometry (test hook)"""
        bases, scalars = _c(bases, np.uint64).reshape(-1, 12), _c(scalars, np.uint64).reshape(-1, 4)
        out = np.zeros(12, dtype=np.uint64)
        if window_bits is None:
            L.check(self._lib.frcs_msm_g1(self.h, bases.shape[0], _p(bases), _p(scalars), _p(out)), "frcs_msm_g1")
        else:
            L.check(self._lib.frcs_debug_msm_g1(self.h, window_bits, bases.shape[0], _p(bases), _p(scalars), _p(out)),
                    "frcs_debug_msm_g1")
        return out

    def msm_g2(self, bases, scalars, window_bits=None):
        bases, scalars = _c(bases, np.uint64).reshape(-1, 24), _c(scalars, np.uint64).reshape(-1, 4)
        out = np.zeros(24, dtype=np.uint64)
        if window_bits is None:
            L.check(self._lib.frcs_msm_g2(self.h, bases.shape[0], _p(bases), _p(scalars), _p(out)), "frcs_msm_g2")
        else:
            L.check(self._lib.frcs_debug_msm_g2(self.h, window_bits, bases.shape[0], _p(bases), _p(scalars), _p(out)),
                    "frcs_debug_msm_g2")
        return out

    # -- proving --------------------------------------------------------------------------
    def load_pk(self, pk: ProvingKey):
        self._pk = pk  # keep the host arrays alive during the call
        v = pk.view()
        L.check(self._lib.frcs_load_pk(self.h, C.byref(v)), "frcs_load_pk")

    def setup(self, trapdoor, shard=0, n_shards=1):
        """Groth16::circuit_specific_setup on the device from explicit toxic waste (7 x 4 uint64 Montgomery:
        alpha, beta, gamma, delta, tau, g1_scalar, g2_scalar).  Installs the proving key (or its base-range shard
        `shard` of `n_shards`, for a proof split over several GPUs) in this context and returns the verifying key as a
        dict of affine points."""
        td = _c(trapdoor, np.uint64).reshape(7, 4)
        a1 = np.zeros(12, dtype=np.uint64)
        g2 = np.zeros((3, 24), dtype=np.uint64)
        ic = np.zeros((self.n_inst, 12), dtype=np.uint64)
        L.check(self._lib.frcs_setup_shard(self.h, _p(td), shard, n_shards, _p(a1), _p(g2), _p(ic)), "frcs_setup_shard")
        return {"alpha_g1": a1, "beta_g2": g2[0], "gamma_g2": g2[1], "delta_g2": g2[2], "gamma_abc_g1": ic}

    def export_pk(self, name):
        which = ["a_query", "b_g1_query", "b_g2_query", "h_query", "l_query"].index(name)
        n = [self.n_z, self.n_z, self.n_z, (1 << self.domain_log2) - 1, self.n_wit][which]
        out = np.zeros((n, 24 if which == 2 else 12), dtype=np.uint64)
        L.check(self._lib.frcs_export_pk(self.h, which, _p(out)), "frcs_export_pk")
        return out

    def load_pk_shard(self, pk: ProvingKey, shard, n_shards):
        """Base-range shard `shard` of `n_shards` of the proving key (single proof over several GPUs)."""
        self._pk = pk
        v = pk.view()
        L.check(self._lib.frcs_load_pk_shard(self.h, C.byref(v), shard, n_shards), "frcs_load_pk_shard")

    def prove_partial_dev(self, n, d_sig, d_pk, d_hm, d_r, d_s, d_partials, d_status, stream=0):
        """Device pointers (ints); writes n x 144 u64 MSM sums of this context's key shard to d_partials."""
        args = [C.c_void_p(int(x)) for x in (d_sig, d_pk, d_hm, d_r, d_s, d_partials, d_status, stream)]
        L.check(self._lib.frcs_prove_partial_dev(self.h, n, *args), "frcs_prove_partial_dev")

    def prove_split_begin_dev(self, d_sig, d_pk, d_hm, d_r, d_s, d_abc, d_status, stream=0):
        """One proof, key shard: enqueues everything up to this shard's coset vectors in d_abc (3 x 2^domain x 4 u64)."""
        args = [C.c_void_p(int(x)) for x in (d_sig, d_pk, d_hm, d_r, d_s, d_abc, d_status, stream)]
        L.check(self._lib.frcs_prove_split_begin_dev(self.h, *args), "frcs_prove_split_begin_dev")

    def prove_split_finish_dev(self, d_abc, d_partials, stream=0):
        """After the shards exchanged their coset vectors: enqueues the rest, 144 u64 of MSM sums to d_partials."""
        args = [C.c_void_p(int(x)) for x in (d_abc, d_partials, stream)]
        L.check(self._lib.frcs_prove_split_finish_dev(self.h, *args), "frcs_prove_split_finish_dev")

    def prove_batch(self, sig, pk, hm, r, s):
        sig, pk, hm = [_c(x, np.uint16).reshape(-1, self.n) for x in (sig, pk, hm)]
        r, s = _c(r, np.uint64).reshape(-1, 4), _c(s, np.uint64).reshape(-1, 4)
        n = sig.shape[0]
        proofs = np.zeros((n, 48), dtype=np.uint64)
        st = np.zeros(n, dtype=np.int32)
        L.check(self._lib.frcs_prove_batch(self.h, n, _p(sig, L.u16p), _p(pk, L.u16p), _p(hm, L.u16p), _p(r), _p(s),
                                           _p(proofs), _p(st, L.i32p)), "frcs_prove_batch")
        return proofs, st

    def prove_from_z(self, z, r, s):
        z = _c(z, np.uint64).reshape(-1, self.n_z, 4)
        r, s = _c(r, np.uint64).reshape(-1, 4), _c(s, np.uint64).reshape(-1, 4)
        proofs = np.zeros((z.shape[0], 48), dtype=np.uint64)
        L.check(self._lib.frcs_prove_from_z(self.h, z.shape[0], _p(z), _p(r), _p(s), _p(proofs)), "frcs_prove_from_z")
        return proofs

    PROF = {"witness": 0, "r1cs": 1, "witness_map": 2, "msm_h_accum": 3, "msm_h": 4, "msm_a": 5, "msm_b_g1": 6,
            "msm_l": 7, "msm_b_g2": 8, "host_tail": 9, "ntt": 10, "sort_z": 11, "sort_lh": 12, "group": 13,
            "sort_digits": 14, "sort_scatter": 15}

    def profile_enable(self, on=True):
        L.check(self._lib.frcs_profile_enable(self.h, int(on)), "frcs_profile_enable")

    def profile_get(self, name, reset=True):
        """(total device ms, launches, work counter) of one stage since the last reset"""
        ms, cnt, work = C.c_double(0), C.c_uint64(0), C.c_uint64(0)
        L.check(self._lib.frcs_profile_get(self.h, self.PROF[name], C.byref(ms), C.byref(cnt), C.byref(work),
                                           int(reset)), "frcs_profile_get")
        return ms.value, cnt.value, work.value

    def launch_count(self):
        return int(self._lib.frcs_launch_count(self.h))

    def imad_peak(self):
        v = C.c_double(0)
        L.check(self._lib.frcs_imad_peak(self.h, C.byref(v)), "frcs_imad_peak")
        return v.value


def shard_ranges(n_inst, n_wit, domain_log2, shard, n_shards):
    """The contiguous base ranges (lo, hi) shard `shard` of `n_shards` holds of the a/b queries, l_query and
    h_query: the same arithmetic as frcs_load_pk_shard."""
    nv, nh = n_inst + n_wit, (1 << domain_log2) - 1

    def rng(length):
        return length * shard // n_shards, length * (shard + 1) // n_shards
    return {"z": rng(nv), "l": rng(n_wit), "h": rng(nh)}


PARTIAL_WORDS = 144  # u64 per proof and shard: A | B1 | L+H | unused (G1 XYZZ), B2 (G2 XYZZ)


def combine_partials(partials, r, s):
    """partials: [n_shards, n, 144] uint64 (the gathered MSM sums of every shard) -> [n, 48] affine proofs.
    Host code only (no GPU needed): the O(1) tail of create_proof."""
    partials = _c(partials, np.uint64)
    n_shards, n = partials.shape[0], partials.shape[1]
    r, s = _c(r, np.uint64).reshape(-1, 4), _c(s, np.uint64).reshape(-1, 4)
    proofs = np.zeros((n, 48), dtype=np.uint64)
    L.check(L.load().frcs_combine_partials(n_shards, n, _p(partials), _p(r), _p(s), _p(proofs)), "frcs_combine_partials")
    return proofs


def proof_compress(proof_affine):
    """ark-serialize compressed Proof bytes (A 48 | B 96 | C 48)."""
    out = np.zeros(192, dtype=np.uint8)
    pa = _c(proof_affine, np.uint64)
    L.check(L.load().frcs_proof_compress(_p(pa), _p(out, L.u8p)), "frcs_proof_compress")
    return bytes(out)


def verify_proof(vk, proof, public_inputs):
    """ark_groth16::verify_proof(&pvk, &proof, &public_inputs) (examples/pok_sig.rs:45-47).  vk: the dict returned
    by Context.setup (alpha_g1, beta_g2, gamma_g2, delta_g2, gamma_abc_g1); proof: 48 uint64 affine;
    public_inputs: [n, 4] Montgomery Fr (pk_ntt then hm_ntt, i.e. z[1:n_instance]).  Host code, no GPU needed."""
    g2 = np.ascontiguousarray(np.stack([vk["beta_g2"], vk["gamma_g2"], vk["delta_g2"]]), dtype=np.uint64)
    ic = _c(vk["gamma_abc_g1"], np.uint64)
    x = _c(public_inputs, np.uint64).reshape(-1, 4)
    if x.shape[0] + 1 != ic.shape[0]:
        raise ValueError("verify_proof: %d public inputs for %d gamma_abc_g1 points" % (x.shape[0], ic.shape[0]))
    rc = L.load().frcs_verify_proof(_p(_c(vk["alpha_g1"], np.uint64)), _p(g2), _p(ic), x.shape[0], _p(x),
                                    _p(_c(proof, np.uint64)))
    if rc < 0:
        raise L.FrcsError(rc, "frcs_verify_proof")
    return bool(rc)


class SystemRng:
    """OS entropy (`secrets`), the default source for the Groth16 toxic waste and the (r, s) blinders: the reference
    uses a ChaCha20 CSPRNG (examples/pok_sig.rs:13); a seeded numpy Generator (PCG64 is not a CSPRNG) is for
    reproducible tests and benchmarks only."""

    def integers(self, lo, hi, dtype=np.uint64):
        import secrets
        return lo + secrets.randbelow(hi - lo)


def random_trapdoor(rng=None):
    """The toxic waste ark-groth16's generate_random_parameters draws (alpha, beta, gamma, delta, then the two
    generators as scalars of the standard ones, then tau), 7 x 4 uint64 Montgomery, in frcs_setup's order.
    rng = None: OS entropy; pass a seeded numpy Generator only for tests."""
    rng = rng if rng is not None else SystemRng()
    alpha, beta, gamma, delta, g1s, g2s, tau = [fr_rand(rng, nonzero=True) for _ in range(7)]
    return np.stack([alpha, beta, gamma, delta, tau, g1s, g2s])


def fr_rand(rng=None, nonzero=False):
    """Fr::rand (ark-ff 0.3.0): 4 x u64 from the rng, top limb >> 1,
    accept if < r; the limbs are used as the Montgomery representation directly.  rng = None: OS entropy."""
    rng = rng if rng is not None else SystemRng()
    R = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
    while True:
        limbs = [int(rng.integers(0, 1 << 64, dtype=np.uint64)) for _ in range(4)]
        limbs[3] >>= 1
        v = sum(x << (64 * i) for i, x in enumerate(limbs))
        if v < R and (v or not nonzero):
            return np.array(limbs, dtype=np.uint64)


def create_random_proof(ctx: Context, circuit: FalconNTTVerificationCircuit, rng=None):
    """ark_groth16::create_random_proof(circuit, &pk, rng): r then s are drawn with
    Fr::rand, then create_proof(circuit, pk, r, s).  ctx must hold the proving key.
    rng = None draws (r, s) from OS entropy (predictable blinders break zero-knowledge)."""
    r = fr_rand(rng)
    s = fr_rand(rng)
    proofs, st = ctx.prove_batch(circuit.sig, circuit.pk, circuit.hm, r, s)
    if st[0] != 0:
        raise ValueError("witness generation failed with status %d (the reference panics here)" % st[0])
    return proofs[0]
