/*
 * falcon_r1cs_b200 — C ABI of the B200-native prover backend for the Falcon
 * signature-verification circuit of zhenfeizhang/falcon-r1cs.
 *
 * The reference has no FFI of its own: its boundary is the Rust API driven by
 * falcon-r1cs/examples/pok_sig.rs:11-48.  Each entry point below names the
 * reference interface it stands in for; INTEGRATION.md shows the Rust binding.
 *
 * Conventions
 *  - every function returns int32_t: FRCS_OK (0) or a negative FRCS_E_* code;
 *    nothing unwinds across the boundary; frcs_last_error() gives a thread-local
 *    message for the last failure on the calling thread.
 *  - the caller owns every buffer it passes; the library owns only frcs_ctx.
 *  - Fr elements are 4 x uint64 little-endian limbs in Montgomery form
 *    (x * 2^256 mod r): the in-memory image of ark_ff::Fp256.
 *    Fq elements are 6 x uint64, Montgomery (x * 2^384 mod p): ark_ff::Fp384.
 *  - G1 affine = x | y (12 x uint64); G2 affine = x.c0 | x.c1 | y.c0 | y.c1
 *    (24 x uint64).  The point at infinity is encoded as all-zero coordinates.
 *  - "host" entry points take host pointers, do the host<->device copies and return when
 *    their outputs are written.  "_dev" entry points take device pointers and run on `stream`
 *    (a cudaStream_t passed as void*): frcs_witness_batch_dev, frcs_r1cs_eval_batch_dev,
 *    frcs_witness_check_batch_dev and frcs_witness_map_dev only enqueue work and do not
 *    synchronise (they may synchronise the device once, the first time a larger batch needs
 *    bigger internal buffers).  frcs_prove_batch_dev and frcs_prove_partial_dev synchronise
 *    with `stream`: the last step of a proof (three inversions and two scalar multiplications)
 *    runs on host threads, overlapped with the next group of proofs on the device, so the
 *    call returns with the proofs written.
 *  - calls on one context must be serialised by the caller (the reference's
 *    ConstraintSystemRef is Rc<RefCell>, i.e. single-threaded as well).
 *  - there is no CPU fallback: without a usable CUDA device every call fails
 *    with FRCS_E_CUDA.
 */
#ifndef FALCON_R1CS_B200_H
#define FALCON_R1CS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FRCS_OK 0
#define FRCS_E_INVALID_ARG (-1)
#define FRCS_E_CUDA (-2)
#define FRCS_E_NO_PK (-3)
#define FRCS_E_ALLOC (-4)
#define FRCS_E_INVALID_POINT (-5) /* a group element with out-of-range coordinates, off the curve or outside the subgroup */
/* per-signature status codes: where the reference panics during synthesis */
#define FRCS_E_COEFF_RANGE (-16) /* gadgets/range_proofs.rs:58-60 (value >= 12289) */
#define FRCS_E_NORM_BOUND (-17)  /* gadgets/range_proofs.rs:114-117, 205-208 */

#define FRCS_KIND_NTT 0        /* circuits/falcon_ntt.rs:26-123 */
#define FRCS_KIND_SCHOOLBOOK 1 /* circuits/falcon_schoolbook.rs:26-132 */
#define FRCS_KIND_DUAL_NTT 2   /* circuits/falcon_dual_ntt.rs:26-132 + gadgets/dual_poly.rs:8-52 */

typedef struct frcs_ctx frcs_ctx;

typedef struct frcs_shape {
  uint32_t logn;        /* 9 = Falcon-512, 10 = Falcon-1024 (cargo features, Cargo.toml:28-32) */
  uint32_t kind;        /* FRCS_KIND_* */
  uint32_t n_instance;  /* cs.num_instance_variables(), incl. the constant One */
  uint32_t n_witness;   /* cs.num_witness_variables() */
  uint32_t n_constraints; /* cs.num_constraints()  (README.md:41-56) */
  uint32_t domain_log2; /* log2 of the Radix2EvaluationDomain size */
  uint64_t nnz_a, nnz_b, nnz_c; /* non-zeros of cs.to_matrices() */
} frcs_shape;

/* Proving-key view: what ark_groth16::ProvingKey<Bls12_381> holds, as plain arrays
 * (do not rely on Rust struct layout).  query lengths: a, b_g1, b_g2 = n_instance +
 * n_witness; h = domain - 1; l = n_witness. */
typedef struct frcs_pk_view {
  const uint64_t* alpha_g1;  /* 12 */
  const uint64_t* beta_g1;   /* 12 */
  const uint64_t* delta_g1;  /* 12 */
  const uint64_t* beta_g2;   /* 24 */
  const uint64_t* delta_g2;  /* 24 */
  const uint64_t* a_query;    uint64_t a_len;     /* G1 */
  const uint64_t* b_g1_query; uint64_t b_g1_len;  /* G1 */
  const uint64_t* b_g2_query; uint64_t b_g2_len;  /* G2 */
  const uint64_t* h_query;    uint64_t h_len;     /* G1 */
  const uint64_t* l_query;    uint64_t l_len;     /* G1 */
} frcs_pk_view;

/* ---- context ------------------------------------------------------------------
 * Builds the circuit (A/B/C in CSR on the device, witness layout, NTT tables) for
 * FalconNTTVerificationCircuit / FalconSchoolBookVerificationCircuit / FalconDualNTTVerificationCircuit
 * (circuits/falcon_ntt.rs:8-18, circuits/falcon_schoolbook.rs:8-18, circuits/falcon_dual_ntt.rs:8-18) on CUDA
 * device `device`.  One context per (device, circuit). */
int32_t frcs_ctx_create(uint32_t logn, uint32_t kind, int32_t device, frcs_ctx** out);
void frcs_ctx_destroy(frcs_ctx* ctx);
const char* frcs_last_error(void);
/* cs.num_instance_variables() / num_witness_variables() / num_constraints() */
int32_t frcs_shape_get(const frcs_ctx* ctx, frcs_shape* out);
/* cs.to_matrices(): CSR of matrix `which` (0=A,1=B,2=C): row_ptr[n_constraints+1],
 * col[nnz], val[nnz*4] (Montgomery).  For differential tests against arkworks. */
int32_t frcs_get_matrix(frcs_ctx* ctx, int32_t which, uint32_t* row_ptr, uint32_t* col, uint64_t* val);

/* ---- (1) witness generation: ConstraintSynthesizer::generate_constraints in Prove
 * mode (circuits/falcon_ntt.rs:26-123), batched.  Inputs are the coefficient
 * vectors the reference derives from (pk, msg, sig): sig = Polynomial::from(&sig),
 * pk = Polynomial::from(&pk), hm = Polynomial::from_hash_of_message(msg, nonce),
 * each n x N uint16 in [0, 12289).  z_out: n x (n_instance+n_witness) x 4 uint64 =
 * instance_assignment ++ witness_assignment.  status[i]: FRCS_OK or the
 * FRCS_E_COEFF_RANGE / FRCS_E_NORM_BOUND of the first panic site the reference
 * would hit (z is still written, as in the reference's #[cfg(test)] build). */
int32_t frcs_witness_batch(frcs_ctx* ctx, uint64_t n, const uint16_t* sig, const uint16_t* pk, const uint16_t* hm,
                           uint64_t* z_out, int32_t* status);
int32_t frcs_witness_batch_dev(frcs_ctx* ctx, uint64_t n, const uint16_t* d_sig, const uint16_t* d_pk,
                               const uint16_t* d_hm, uint64_t* d_z, int32_t* d_status, void* stream);

/* ---- (2) R1CS evaluation: evaluate_constraint over cs.to_matrices() (ark-groth16
 * r1cs_to_qap.rs witness_map) and cs.which_is_unsatisfied().  az/bz/cz: n x
 * n_constraints x 4 (any may be NULL); first_unsat[i] = first row with
 * <A,z><B,z> != <C,z>, or -1. */
int32_t frcs_r1cs_eval_batch(frcs_ctx* ctx, uint64_t n, const uint64_t* z, uint64_t* az, uint64_t* bz, uint64_t* cz,
                             int64_t* first_unsat);
int32_t frcs_r1cs_eval_batch_dev(frcs_ctx* ctx, uint64_t n, const uint64_t* d_z, uint64_t* d_az, uint64_t* d_bz,
                                 uint64_t* d_cz, int64_t* d_first_unsat, void* stream);

/* (1) + (2) fused for batches that only need the verdict (BASELINE configs[2]: witness generation + R1CS
 * satisfaction for 65,536 signatures): generate_constraints followed by cs.which_is_unsatisfied()
 * (circuits/falcon_ntt.rs:143-159), the assignments never leave the device.  first_unsat[i] = -1 if satisfied. */
int32_t frcs_witness_check_batch(frcs_ctx* ctx, uint64_t n, const uint16_t* sig, const uint16_t* pk, const uint16_t* hm,
                                 int64_t* first_unsat, int32_t* status);
int32_t frcs_witness_check_batch_dev(frcs_ctx* ctx, uint64_t n, const uint16_t* d_sig, const uint16_t* d_pk,
                                     const uint16_t* d_hm, int64_t* d_first_unsat, int32_t* d_status, void* stream);

/* ---- (3) R1CStoQAP::witness_map (ark-groth16 0.3.0): z -> h, 2^domain_log2 x 4 */
int32_t frcs_witness_map(frcs_ctx* ctx, const uint64_t* z, uint64_t* h_out);
int32_t frcs_witness_map_dev(frcs_ctx* ctx, const uint64_t* d_z, uint64_t* d_h, void* stream);
/* Radix2EvaluationDomain primitives on 2^log_size elements (parity tests):
 * op 0 fft_in_place, 1 ifft_in_place, 2 coset_fft_in_place, 3 coset_ifft_in_place */
int32_t frcs_domain_op(frcs_ctx* ctx, uint32_t log_size, int32_t op, uint64_t* data);

/* ---- (4) VariableBaseMSM::multi_scalar_mul (ark-ec 0.3.0).  scalars are canonical
 * integers (into_repr()), n x 4 uint64; result affine (12 / 24 uint64). */
int32_t frcs_msm_g1(frcs_ctx* ctx, uint64_t n, const uint64_t* bases, const uint64_t* scalars, uint64_t* out);
int32_t frcs_msm_g2(frcs_ctx* ctx, uint64_t n, const uint64_t* bases, const uint64_t* scalars, uint64_t* out);

/* ---- gadget entry points (falcon-r1cs/src/gadgets/mod.rs:7-11), batched over n independent instances.
 * In the reference a gadget call allocates its witnesses in `cs` and enforces its rows, and the gadget tests then
 * read cs.is_satisfied() (range panics are compiled out under #[cfg(test)], range_proofs.rs:55-60).  Here: operand
 * values in (Montgomery Fr, 4 x uint64 each), out: the gadget's witnesses in allocation order (`wit`, n x n_witness
 * x 4), first_unsat[i] = first violated row of the gadget's own rows or -1 (cs.which_is_unsatisfied()), status[i] =
 * FRCS_OK or the code of the panic a non-test build would hit.  Gadget ids for frcs_gadget_shape: 0 mod_q, 1 add_mod,
 * 2 enforce_less_than_q, 3 is_less_than_6144, 4 enforce_less_than_norm_bound, 5 ntt_circuit. */
int32_t frcs_gadget_shape(const frcs_ctx* ctx, int32_t gadget, uint32_t* n_operands, uint32_t* n_witness,
                          uint32_t* n_rows);
/* mod_q(cs, &a, q) (gadgets/arithmetics.rs:105-149): wit = t, b, 27 range witnesses of b (29; the output b is wit[1]).
 * expected (may be NULL): as in test_mod_q (arithmetics.rs:322-324) one more witness with this value and the row
 * b.enforce_equal(expected); wit then has 30 entries per instance. */
int32_t frcs_gadget_mod_q(frcs_ctx* ctx, uint64_t n, const uint64_t* a, const uint64_t* expected, uint64_t* wit,
                          int64_t* first_unsat, int32_t* status);
/* add_mod(cs, &a, &b, q) (gadgets/arithmetics.rs:214-262): ab = n x 2 operands (a_i, b_i); wit = t, c, 27 range
 * witnesses of c; expected as above (arithmetics.rs:461-463). */
int32_t frcs_gadget_add_mod(frcs_ctx* ctx, uint64_t n, const uint64_t* ab, const uint64_t* expected, uint64_t* wit,
                            int64_t* first_unsat, int32_t* status);
/* enforce_less_than_q(cs, &a) (gadgets/range_proofs.rs:42-94): wit = 14 bits, 11 or-results, 2 and-results (27);
 * status FRCS_E_COEFF_RANGE where the reference panics (range_proofs.rs:58-60). */
int32_t frcs_gadget_less_than_q(frcs_ctx* ctx, uint64_t n, const uint64_t* a, uint64_t* wit, int64_t* first_unsat,
                                int32_t* status);
/* is_less_than_6144(cs, &a) (gadgets/range_proofs.rs:289-333): wit = 14 bits, y1, y2 (16); the returned Boolean is
 * wit[15].  enforce_true != 0 adds .enforce_equal(&Boolean::TRUE) as in test_range_proof_half_q (:512-513). */
int32_t frcs_gadget_less_than_6144(frcs_ctx* ctx, uint64_t n, const uint64_t* a, int32_t enforce_true, uint64_t* wit,
                                   int64_t* first_unsat, int32_t* status);
/* enforce_less_than_norm_bound(cs, &a) (gadgets/range_proofs.rs:274-284 -> :100-186 / :192-272 by the context's
 * parameter set): wit = 26 / 27 bits then the k-ary and chain results (50 / 52); status FRCS_E_NORM_BOUND where the
 * reference panics (:114-117, :205-208). */
int32_t frcs_gadget_norm_bound(frcs_ctx* ctx, uint64_t n, const uint64_t* a, uint64_t* wit, int64_t* first_unsat,
                               int32_t* status);
/* NTTPolyVar::ntt_circuit(cs, &poly, const_vars, param) (gadgets/poly.rs:104-159): poly = n x N coefficients in
 * [0, q); values (may be NULL) = the N outputs (== NTTPolynomial::from(&poly), poly.rs:292-297); wit (may be NULL) =
 * n x 29N x 4: per output t, b and the 27 range witnesses of b; first_unsat over the gadget's 30N rows. */
int32_t frcs_gadget_ntt_circuit(frcs_ctx* ctx, uint64_t n, const uint16_t* poly, uint16_t* values, uint64_t* wit,
                                int64_t* first_unsat, int32_t* status);

/* ---- proving key (ark_groth16::ProvingKey, produced by circuit_specific_setup,
 * pok_sig.rs:30-31): uploaded once, bases pre-processed on the device. */
int32_t frcs_load_pk(frcs_ctx* ctx, const frcs_pk_view* pk);

/* Groth16::circuit_specific_setup (pok_sig.rs:30-31) -> ark_groth16::generate_parameters for the context's
 * circuit, on the device, from explicit toxic waste: trapdoor = 7 x 4 uint64 Montgomery Fr
 * (alpha, beta, gamma, delta, tau, g1_scalar, g2_scalar; the group generators are g1_scalar*G1 and
 * g2_scalar*G2, which ark-groth16 draws at random).  The proving key stays on the device and is installed
 * as frcs_load_pk would; the verifying key is written to the host: vk_alpha_g1 (12), vk_g2 = beta_g2 |
 * gamma_g2 | delta_g2 (3 x 24), gamma_abc_g1 (n_instance x 12).  Any output may be NULL. */
int32_t frcs_setup(frcs_ctx* ctx, const uint64_t* trapdoor, uint64_t* vk_alpha_g1, uint64_t* vk_g2,
                   uint64_t* gamma_abc_g1);
/* the same parameter generation, keeping only base-range shard `shard` of `n_shards` of the proving key in this context
 * (see frcs_load_pk_shard): every GPU of a split proof runs it with the same trapdoor and its own shard index. */
int32_t frcs_setup_shard(frcs_ctx* ctx, const uint64_t* trapdoor, uint32_t shard, uint32_t n_shards, uint64_t* vk_alpha_g1,
                         uint64_t* vk_g2, uint64_t* gamma_abc_g1);
/* the queries of the proving key held by the context (which: 0 a, 1 b_g1, 2 b_g2, 3 h, 4 l), affine */
int32_t frcs_export_pk(frcs_ctx* ctx, int32_t which, uint64_t* out);

/* ---- whole path: ark_groth16::create_proof(circuit, pk, r, s) (pok_sig.rs:32 calls
 * create_random_proof, which draws r then s with Fr::rand and calls this).
 * r, s: n x 4 Montgomery.  proofs_out: n x 48 uint64 = A (G1) | B (G2) | C (G1)
 * affine.  status as for frcs_witness_batch. */
int32_t frcs_prove_batch(frcs_ctx* ctx, uint64_t n, const uint16_t* sig, const uint16_t* pk, const uint16_t* hm,
                         const uint64_t* r, const uint64_t* s, uint64_t* proofs_out, int32_t* status);
/* same, starting from full assignments z (n x (n_instance+n_witness) x 4) */
int32_t frcs_prove_from_z(frcs_ctx* ctx, uint64_t n, const uint64_t* z, const uint64_t* r, const uint64_t* s,
                          uint64_t* proofs_out);
/* device-resident variant: inputs already in HBM, proofs written to HBM */
int32_t frcs_prove_batch_dev(frcs_ctx* ctx, uint64_t n, const uint16_t* d_sig, const uint16_t* d_pk,
                             const uint16_t* d_hm, const uint64_t* d_r, const uint64_t* d_s, uint64_t* d_proofs,
                             int32_t* d_status, void* stream);
/* ---- one proof over several GPUs: the proving key is split by base range (shard k of n holds the
 * k-th contiguous slice of every query; the constant bases alpha/beta/delta live on shard 0).  Every
 * shard recomputes z and h (cheap) and runs its slice of the MSMs; the only exchange is one gather of
 * 144 u64 per proof and shard (A | B1 | L+H | unused in G1 XYZZ, B2 in G2 XYZZ), e.g. ncclAllGather,
 * after which frcs_combine_partials (host, any rank) adds the shards and finishes create_proof.
 * This is the multi-GPU form of VariableBaseMSM::multi_scalar_mul inside create_proof (pok_sig.rs:32);
 * the reference has no counterpart (single process, rayon). */
int32_t frcs_load_pk_shard(frcs_ctx* ctx, const frcs_pk_view* pk, uint32_t shard, uint32_t n_shards);
int32_t frcs_prove_partial_dev(frcs_ctx* ctx, uint64_t n, const uint16_t* d_sig, const uint16_t* d_pk,
                               const uint16_t* d_hm, const uint64_t* d_r, const uint64_t* d_s, uint64_t* d_partials,
                               int32_t* d_status, void* stream);
/* Lower latency for ONE proof: the witness map is shared between the shards too.  begin: witness generation, this
 * shard's A / B chains, and ifft + coset_fft of the vectors v of (a, b, c) with v % n_shards == shard, left in
 * d_abc[v] (d_abc: device, 3 x 2^domain_log2 x 4 u64, caller-owned so that it can be handed to NCCL).  The caller then
 * broadcasts d_abc[v] from shard v % n_shards to all shards (ncclBroadcast on `stream` or ordered after it), and
 * finish does (a.b - c)/Z, coset_ifft, this shard's L + H chain and writes the 144 u64 of MSM sums to d_partials.
 * Both calls only enqueue work; the results are ordered on `stream`.  One proof in flight per context. */
int32_t frcs_prove_split_begin_dev(frcs_ctx* ctx, const uint16_t* d_sig, const uint16_t* d_pk, const uint16_t* d_hm,
                                   const uint64_t* d_r, const uint64_t* d_s, uint64_t* d_abc, int32_t* d_status,
                                   void* stream);
int32_t frcs_prove_split_finish_dev(frcs_ctx* ctx, uint64_t* d_abc, uint64_t* d_partials, void* stream);
/* partials: [n_shards][n][144] u64 (host); r, s: n x 4 Montgomery; proofs_out: n x 48 u64 */
int32_t frcs_combine_partials(uint32_t n_shards, uint64_t n, const uint64_t* partials, const uint64_t* r,
                              const uint64_t* s, uint64_t* proofs_out);
/* ark-serialize 0.3 compressed Proof (48 + 96 + 48 bytes) from the affine form */
int32_t frcs_proof_compress(const uint64_t* proof_affine, uint8_t* out192);

/* ---- ark_groth16::verify_proof(&pvk, &proof, &public_inputs) (pok_sig.rs:45-47), on the host (no GPU needed):
 * 1 = the proof verifies, 0 = it does not, negative = error.  vk as written by frcs_setup; public_inputs:
 * n_inputs x 4 Montgomery Fr without the leading One (pok_sig.rs:33-44: pk_ntt then hm_ntt coefficients);
 * proof: A (12) | B (24) | C (12) affine. */
int32_t frcs_verify_proof(const uint64_t* vk_alpha_g1, const uint64_t* vk_g2, const uint64_t* gamma_abc_g1,
                          uint64_t n_inputs, const uint64_t* public_inputs, const uint64_t* proof);
/* Point validation (what ark-ec does when it deserialises a Proof / VerifyingKey): coordinates below p, on the curve,
 * in the prime-order subgroup; (0, 0) = infinity is accepted.  1 = valid, 0 = not.  frcs_verify_proof validates the
 * three proof points itself (FRCS_E_INVALID_POINT otherwise) and range-checks the key; validate a key once with
 * frcs_vk_validate before trusting it. */
int32_t frcs_g1_validate(const uint64_t* p);
int32_t frcs_g2_validate(const uint64_t* p);
int32_t frcs_vk_validate(const uint64_t* vk_alpha_g1, const uint64_t* vk_g2, const uint64_t* gamma_abc_g1, uint64_t n_inputs);
/* pairing identities for tests: e(p1, q1) == e(p2, q2) and e(p, q) == 1 (1 / 0) */
int32_t frcs_pairing_eq(const uint64_t* p1, const uint64_t* q1, const uint64_t* p2, const uint64_t* q2);
int32_t frcs_pairing_is_one(const uint64_t* p, const uint64_t* q);

/* ---- instrumentation -------------------------------------------------------------
 * number of kernels this library has launched on ctx since creation */
uint64_t frcs_launch_count(const frcs_ctx* ctx);
/* per-stage device timing (CUDA events on the launching streams).  ids: 0 witness gen,
 * 1 R1CS evaluation, 2 witness map (incl. 1), 3 bucket-accumulation kernel of the h MSM
 * (work = number of point additions of the last launch), 4..8 whole MSMs h, a, b_g1, l,
 * b_g2, 9 host tail (wall clock), 10 NTT launches of the witness map, 11 / 12 digit sort of the z scalars / of the
 * (w, -rs, h) scalars, 13 one whole proof group (witness map through the join of the MSM streams). */
int32_t frcs_profile_enable(frcs_ctx* ctx, int32_t on);
int32_t frcs_profile_get(frcs_ctx* ctx, int32_t id, double* ms_total, uint64_t* count, uint64_t* work, int32_t reset);
/* self-tests of the field / curve code: op selects the operation, see csrc/selftest.cu.
 * on_device = 0 runs the host build of the same source (no GPU needed). */
int32_t frcs_selftest(int32_t op, int32_t on_device, const uint64_t* in, uint64_t n, uint64_t* out);
/* test hooks: the pre-processed MSM table of n bases, 16 windows x n affine points
 * (window k holds 2^(16k) P_i) */
int32_t frcs_debug_windows_g1(frcs_ctx* ctx, uint64_t n, const uint64_t* bases, uint64_t* out);
int32_t frcs_debug_windows_g2(frcs_ctx* ctx, uint64_t n, const uint64_t* bases, uint64_t* out);
/* test hook, host only (no GPU needed): cs.to_matrices() as emitted by the circuit compiler for (logn, kind), matrix
 * `which` (0 = A, 1 = B, 2 = C); val = canonical integers, 4 x uint64 per entry.  counts (may be NULL): n_instance,
 * n_witness, n_constraints, nnz.  row_ptr / col / val may be NULL (query the counts first).  kind = 16 + g (+ 8)
 * selects the stand-alone circuit of gadget g (frcs_gadget_shape ids), + 8 with the expected-output row. */
int32_t frcs_debug_host_matrix(uint32_t logn, uint32_t kind, int32_t which, uint32_t* row_ptr, uint32_t* col,
                               uint64_t* val, uint64_t* counts);
/* test hooks: frcs_msm_g1 / frcs_msm_g2 through either window geometry of the MSM subsystem (window_bits = 16: 16
 * windows sharing 32768 buckets, used for the dense h_query scalars and by frcs_msm_g1/g2; window_bits = 8: 32 windows
 * sharing 128 buckets, used for the 0/1-heavy assignment against a_query / b_g1_query / b_g2_query).  Same semantics
 * as VariableBaseMSM::multi_scalar_mul (ark-ec 0.3.0). */
int32_t frcs_debug_msm_g1(frcs_ctx* ctx, int32_t window_bits, uint64_t n, const uint64_t* bases, const uint64_t* scalars,
                          uint64_t* out);
int32_t frcs_debug_msm_g2(frcs_ctx* ctx, int32_t window_bits, uint64_t n, const uint64_t* bases, const uint64_t* scalars,
                          uint64_t* out);
/* test hook, host only: the balanced signed digits (5 per coefficient, base 2^bits; bits = 28 or 32, or 0 for the
 * 32-bit records of the warp-per-row kernel) that the long-row R1CS kernels keep for the integer behind each
 * canonical Fr coefficient c (c itself or c - r); ok[i] = 0 when the integer does not fit.  No reference counterpart:
 * arkworks stores these coefficients (products of NTT twiddles, gadgets/poly.rs:115-149) as field elements. */
int32_t frcs_debug_digits(const uint64_t* coeffs, uint64_t n, int32_t bits, int64_t* digits, int32_t* ok);
/* IMAD.WIDE limb-product peak microbenchmark: returns limb-products per second */
int32_t frcs_imad_peak(frcs_ctx* ctx, double* lp_per_s);

#ifdef __cplusplus
}
#endif
#endif
