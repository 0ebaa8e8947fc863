//! Rust binding of `libfalcon_r1cs_b200.so` (include/falcon_r1cs_b200.h) that keeps the call shape of
//! `falcon-r1cs/examples/pok_sig.rs`:
//!
//! ```ignore
//! let circuit = FalconNTTVerificationCircuit::build_circuit(pk, msg.to_vec(), sig);     // unchanged
//! let (pp, vk) = Groth16::<Bls12_381>::circuit_specific_setup(circuit.clone(), &mut rng)?; // unchanged
//! let gpu = GpuProver::new(&pp, 0)?;                                                     // once per key and device
//! let stmt = FalconStatement::new(pk, msg.to_vec(), sig);          // the same three arguments as build_circuit
//! let proof = gpu.create_random_proof(&stmt, &mut rng)?;           // was: create_random_proof(circuit, &pp, &mut rng)
//! assert!(verify_proof(&pvk, &proof, &public_inputs)?);                                  // unchanged
//! ```
//!
//! The fields of `FalconNTTVerificationCircuit` are private in the reference (circuits/falcon_ntt.rs:8-12 exposes
//! only `build_circuit`), so this crate does not reach into that struct: it takes the `(PublicKey, msg, Signature)`
//! triple in a type of its own, `FalconStatement` (`PublicKey` and `Signature` are `Copy`, examples/pok_sig.rs:25-35
//! uses them after the move).  Nothing of the reference has to be patched.
//!
//! STATUS: UNCOMPILED.  The image this repository is built in has no cargo / rustc and no vendored arkworks, so
//! this crate has never been through a compiler; the C ABI it binds is exercised by tests/ through ctypes, and
//! tests/differential.rs is the harness to run first once a toolchain exists.  arkworks 0.3 field elements are `Fp256(BigInteger256([u64; 4]))` /
//! `Fp384(BigInteger384([u64; 6]))` in Montgomery form, which is exactly the layout the library expects, so
//! limbs are copied, never converted.
#![allow(non_camel_case_types)]

use ark_bls12_381::{Bls12_381, Fq, Fq2, Fr, G1Affine, G2Affine};
use ark_ff::{PrimeField, UniformRand, Zero};
use ark_groth16::{Proof, ProvingKey};
use ark_relations::r1cs::SynthesisError;
use falcon_rust::{Polynomial, PublicKey, Signature, LOG_N, N};
use rand::Rng;
use std::os::raw::c_char;

#[repr(C)]
pub struct frcs_ctx {
    _private: [u8; 0],
}

#[repr(C)]
pub struct frcs_pk_view {
    pub alpha_g1: *const u64,
    pub beta_g1: *const u64,
    pub delta_g1: *const u64,
    pub beta_g2: *const u64,
    pub delta_g2: *const u64,
    pub a_query: *const u64,
    pub a_len: u64,
    pub b_g1_query: *const u64,
    pub b_g1_len: u64,
    pub b_g2_query: *const u64,
    pub b_g2_len: u64,
    pub h_query: *const u64,
    pub h_len: u64,
    pub l_query: *const u64,
    pub l_len: u64,
}

#[repr(C)]
#[derive(Default, Clone, Copy, Debug)]
pub struct frcs_shape {
    pub logn: u32,
    pub kind: u32,
    pub n_instance: u32,
    pub n_witness: u32,
    pub n_constraints: u32,
    pub domain_log2: u32,
    pub nnz_a: u64,
    pub nnz_b: u64,
    pub nnz_c: u64,
}

pub const FRCS_OK: i32 = 0;
pub const FRCS_E_COEFF_RANGE: i32 = -16;
pub const FRCS_E_NORM_BOUND: i32 = -17;

#[link(name = "falcon_r1cs_b200")]
extern "C" {
    pub fn frcs_ctx_create(logn: u32, kind: u32, device: i32, out: *mut *mut frcs_ctx) -> i32;
    pub fn frcs_ctx_destroy(ctx: *mut frcs_ctx);
    pub fn frcs_last_error() -> *const c_char;
    pub fn frcs_load_pk(ctx: *mut frcs_ctx, pk: *const frcs_pk_view) -> i32;
    pub fn frcs_prove_batch(
        ctx: *mut frcs_ctx, n: u64, sig: *const u16, pk: *const u16, hm: *const u16, r: *const u64,
        s: *const u64, proofs_out: *mut u64, status: *mut i32,
    ) -> i32;
    pub fn frcs_witness_batch(
        ctx: *mut frcs_ctx, n: u64, sig: *const u16, pk: *const u16, hm: *const u16, z_out: *mut u64,
        status: *mut i32,
    ) -> i32;
    /// generate_constraints + cs.which_is_unsatisfied() for a batch; the assignments stay on the device
    pub fn frcs_witness_check_batch(
        ctx: *mut frcs_ctx, n: u64, sig: *const u16, pk: *const u16, hm: *const u16, first_unsat: *mut i64,
        status: *mut i32,
    ) -> i32;
    pub fn frcs_shape_get(ctx: *const frcs_ctx, out: *mut frcs_shape) -> i32;
    pub fn frcs_get_matrix(ctx: *mut frcs_ctx, which: i32, row_ptr: *mut u32, col: *mut u32, val: *mut u64) -> i32;
    pub fn frcs_proof_compress(proof_affine: *const u64, out192: *mut u8) -> i32;
    pub fn frcs_gadget_mod_q(
        ctx: *mut frcs_ctx, n: u64, a: *const u64, expected: *const u64, wit: *mut u64, first_unsat: *mut i64,
        status: *mut i32,
    ) -> i32;
    /// A z, B z, C z (any of them null to skip) and the first violated row (-1: satisfied) of n assignments
    pub fn frcs_r1cs_eval_batch(
        ctx: *mut frcs_ctx, n: u64, z: *const u64, az: *mut u64, bz: *mut u64, cz: *mut u64, first_unsat: *mut i64,
    ) -> i32;
}

fn last_error() -> String {
    unsafe { std::ffi::CStr::from_ptr(frcs_last_error()).to_string_lossy().into_owned() }
}

fn fq_limbs(x: &Fq, out: &mut Vec<u64>) {
    out.extend_from_slice(&x.0 .0); // Montgomery limbs, as stored
}
fn g1_limbs(v: &[G1Affine]) -> Vec<u64> {
    let mut o = Vec::with_capacity(v.len() * 12);
    for p in v {
        if p.infinity {
            o.extend_from_slice(&[0u64; 12]); // the library encodes infinity as all-zero coordinates
        } else {
            fq_limbs(&p.x, &mut o);
            fq_limbs(&p.y, &mut o);
        }
    }
    o
}
fn g2_limbs(v: &[G2Affine]) -> Vec<u64> {
    let mut o = Vec::with_capacity(v.len() * 24);
    for p in v {
        if p.infinity {
            o.extend_from_slice(&[0u64; 24]);
        } else {
            fq_limbs(&p.x.c0, &mut o);
            fq_limbs(&p.x.c1, &mut o);
            fq_limbs(&p.y.c0, &mut o);
            fq_limbs(&p.y.c1, &mut o);
        }
    }
    o
}
fn fq_from(l: &[u64]) -> Fq {
    let mut b = [0u64; 6];
    b.copy_from_slice(l);
    ark_ff::Fp384::new(ark_ff::BigInteger384(b)) // limbs are already Montgomery: `new` stores them as is
}
fn g1_from(l: &[u64]) -> G1Affine {
    if l.iter().all(|w| *w == 0) {
        return G1Affine::zero();
    }
    G1Affine::new(fq_from(&l[0..6]), fq_from(&l[6..12]), false)
}
fn g2_from(l: &[u64]) -> G2Affine {
    if l.iter().all(|w| *w == 0) {
        return G2Affine::zero();
    }
    G2Affine::new(
        Fq2::new(fq_from(&l[0..6]), fq_from(&l[6..12])),
        Fq2::new(fq_from(&l[12..18]), fq_from(&l[18..24])),
        false,
    )
}

/// The statement of `FalconNTTVerificationCircuit::build_circuit(pk, msg, sig)` (circuits/falcon_ntt.rs:15), held in a
/// type whose fields this crate can read.
#[derive(Clone, Debug)]
pub struct FalconStatement {
    pub pk: PublicKey,
    pub msg: Vec<u8>,
    pub sig: Signature,
}

impl FalconStatement {
    pub fn new(pk: PublicKey, msg: Vec<u8>, sig: Signature) -> Self {
        Self { pk, msg, sig }
    }
    /// The three coefficient vectors the circuit is synthesised from (circuits/falcon_ntt.rs:27-28, 44).
    pub fn polynomials(&self) -> (Polynomial, Polynomial, Polynomial) {
        let sig: Polynomial = (&self.sig).into();
        let pk: Polynomial = (&self.pk).into();
        let hm = Polynomial::from_hash_of_message(self.msg.as_ref(), self.sig.nonce());
        (sig, pk, hm)
    }
}

/// One GPU context holding the circuit matrices and the pre-processed proving key.
pub struct GpuProver {
    ctx: *mut frcs_ctx,
}

impl GpuProver {
    /// Uploads `pk` (from `Groth16::circuit_specific_setup`, examples/pok_sig.rs:30-31) to CUDA device `device`.
    pub fn new(pk: &ProvingKey<Bls12_381>, device: i32) -> Result<Self, String> {
        let mut ctx = std::ptr::null_mut();
        if unsafe { frcs_ctx_create(LOG_N as u32, 0, device, &mut ctx) } != FRCS_OK {
            return Err(last_error());
        }
        let (alpha, beta1, delta1) = (g1_limbs(&[pk.vk.alpha_g1]), g1_limbs(&[pk.beta_g1]), g1_limbs(&[pk.delta_g1]));
        let (beta2, delta2) = (g2_limbs(&[pk.vk.beta_g2]), g2_limbs(&[pk.vk.delta_g2]));
        let (a, b1, b2) = (g1_limbs(&pk.a_query), g1_limbs(&pk.b_g1_query), g2_limbs(&pk.b_g2_query));
        let (h, l) = (g1_limbs(&pk.h_query), g1_limbs(&pk.l_query));
        let view = frcs_pk_view {
            alpha_g1: alpha.as_ptr(), beta_g1: beta1.as_ptr(), delta_g1: delta1.as_ptr(),
            beta_g2: beta2.as_ptr(), delta_g2: delta2.as_ptr(),
            a_query: a.as_ptr(), a_len: pk.a_query.len() as u64,
            b_g1_query: b1.as_ptr(), b_g1_len: pk.b_g1_query.len() as u64,
            b_g2_query: b2.as_ptr(), b_g2_len: pk.b_g2_query.len() as u64,
            h_query: h.as_ptr(), h_len: pk.h_query.len() as u64,
            l_query: l.as_ptr(), l_len: pk.l_query.len() as u64,
        };
        let rc = unsafe { frcs_load_pk(ctx, &view) };
        if rc != FRCS_OK {
            unsafe { frcs_ctx_destroy(ctx) };
            return Err(last_error());
        }
        Ok(Self { ctx })
    }

    /// Same call shape as `ark_groth16::create_random_proof(circuit, &pk, rng)` (examples/pok_sig.rs:32):
    /// `r` then `s` are drawn with `Fr::rand`, as ark-groth16 0.3.0 does, so a seeded rng yields the same proof.
    pub fn create_random_proof<R: Rng>(
        &self, stmt: &FalconStatement, rng: &mut R,
    ) -> Result<Proof<Bls12_381>, SynthesisError> {
        let r = Fr::rand(rng);
        let s = Fr::rand(rng);
        self.create_proof(stmt, r, s)
    }

    /// `ark_groth16::create_proof(circuit, &pk, r, s)`
    pub fn create_proof(
        &self, stmt: &FalconStatement, r: Fr, s: Fr,
    ) -> Result<Proof<Bls12_381>, SynthesisError> {
        let (sig, pk, hm) = stmt.polynomials();
        debug_assert_eq!(sig.coeff().len(), N);
        let mut proof = [0u64; 48];
        let mut status = 0i32;
        let rc = unsafe {
            frcs_prove_batch(
                self.ctx, 1, sig.coeff().as_ptr(), pk.coeff().as_ptr(), hm.coeff().as_ptr(),
                r.0 .0.as_ptr(), s.0 .0.as_ptr(), proof.as_mut_ptr(), &mut status,
            )
        };
        if rc != FRCS_OK {
            eprintln!("falcon_r1cs_b200: {}", last_error());
            return Err(SynthesisError::Unsatisfiable);
        }
        if status == FRCS_E_COEFF_RANGE || status == FRCS_E_NORM_BOUND {
            // the reference panics here in non-test builds (gadgets/range_proofs.rs:58-60,114-117,205-208)
            panic!("invalid input: range proof failed (status {})", status);
        }
        Ok(Proof { a: g1_from(&proof[0..12]), b: g2_from(&proof[12..36]), c: g1_from(&proof[36..48]) })
    }

    /// Batched form: proofs for many (pk, msg, sig) triples in one call.
    pub fn create_random_proofs<R: Rng>(
        &self, circuits: &[FalconStatement], rng: &mut R,
    ) -> Result<Vec<Proof<Bls12_381>>, SynthesisError> {
        let n = circuits.len();
        let (mut sig, mut pk, mut hm) = (Vec::with_capacity(n * N), Vec::with_capacity(n * N), Vec::with_capacity(n * N));
        let (mut r, mut s) = (Vec::with_capacity(n * 4), Vec::with_capacity(n * 4));
        for c in circuits {
            let (ps, pp, ph) = c.polynomials();
            sig.extend_from_slice(ps.coeff());
            pk.extend_from_slice(pp.coeff());
            hm.extend_from_slice(ph.coeff());
            r.extend_from_slice(&Fr::rand(rng).0 .0);
            s.extend_from_slice(&Fr::rand(rng).0 .0);
        }
        let mut proofs = vec![0u64; 48 * n];
        let mut status = vec![0i32; n];
        let rc = unsafe {
            frcs_prove_batch(self.ctx, n as u64, sig.as_ptr(), pk.as_ptr(), hm.as_ptr(), r.as_ptr(), s.as_ptr(),
                             proofs.as_mut_ptr(), status.as_mut_ptr())
        };
        if rc != FRCS_OK || status.iter().any(|x| *x != 0) {
            return Err(SynthesisError::Unsatisfiable);
        }
        Ok(proofs.chunks(48).map(|p| Proof { a: g1_from(&p[0..12]), b: g2_from(&p[12..36]), c: g1_from(&p[36..48]) }).collect())
    }
}

impl GpuProver {
    /// `cs.which_is_unsatisfied()` after `generate_constraints` (circuits/falcon_ntt.rs:143-159) for many statements at
    /// once: `None` = satisfied, `Some(row)` = first violated constraint.  An input the reference would panic on (a
    /// coefficient or the norm out of range) is reported as `Err(status)` for that statement.
    pub fn which_is_unsatisfied_batch(
        &self, circuits: &[FalconStatement],
    ) -> Result<Vec<Result<Option<usize>, i32>>, String> {
        let n = circuits.len();
        let (mut sig, mut pk, mut hm) = (Vec::with_capacity(n * N), Vec::with_capacity(n * N), Vec::with_capacity(n * N));
        for c in circuits {
            let (ps, pp, ph) = c.polynomials();
            sig.extend_from_slice(ps.coeff());
            pk.extend_from_slice(pp.coeff());
            hm.extend_from_slice(ph.coeff());
        }
        let mut first_unsat = vec![0i64; n];
        let mut status = vec![0i32; n];
        let rc = unsafe {
            frcs_witness_check_batch(self.ctx, n as u64, sig.as_ptr(), pk.as_ptr(), hm.as_ptr(),
                                     first_unsat.as_mut_ptr(), status.as_mut_ptr())
        };
        if rc != FRCS_OK {
            return Err(last_error());
        }
        Ok((0..n)
            .map(|i| {
                if status[i] != FRCS_OK {
                    Err(status[i])
                } else if first_unsat[i] < 0 {
                    Ok(None)
                } else {
                    Ok(Some(first_unsat[i] as usize))
                }
            })
            .collect())
    }
}

impl GpuProver {
    /// raw context handle, for the differential tests (tests/differential.rs)
    pub fn raw(&self) -> *mut frcs_ctx {
        self.ctx
    }
}

impl Drop for GpuProver {
    fn drop(&mut self) {
        unsafe { frcs_ctx_destroy(self.ctx) }
    }
}

// The context is used from one thread at a time (as the reference's Rc<RefCell> constraint system is).
unsafe impl Send for GpuProver {}
