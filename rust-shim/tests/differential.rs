//! Differential harness: arkworks 0.3 itself against `libfalcon_r1cs_b200.so`, on a real Falcon signature.
//!
//! STATUS: UNCOMPILED (no cargo / rustc in the image this repository is built in).  It is the first thing to run once
//! a toolchain and the arkworks / falcon-rust sources are available: every "parity unpinned" note in DESIGN.md and
//! SURVEY.md App. F is decided by these four checks.
//!
//!   cargo test --release --test differential -- --nocapture       (with LD_LIBRARY_PATH at libfalcon_r1cs_b200.so)
//!
//! Reference call sites mirrored: circuits/falcon_ntt.rs:143-159 (synthesis + is_satisfied),
//! examples/constraint_counts.rs:49-72 (counts), examples/pok_sig.rs:30-47 (setup, proof, verification).
use ark_bls12_381::{Bls12_381, Fr};
use ark_ff::UniformRand;
use ark_groth16::{create_proof, generate_random_parameters, prepare_verifying_key, verify_proof};
use ark_relations::r1cs::{ConstraintSynthesizer, ConstraintSystem, OptimizationGoal};
use ark_serialize::CanonicalSerialize;
use falcon_r1cs::FalconNTTVerificationCircuit;
use falcon_r1cs_b200::*;
use falcon_rust::{KeyPair, NTTPolynomial, Polynomial, LOG_N, N};
use rand::SeedableRng;
use rand_chacha::ChaCha20Rng;

fn statement() -> (FalconNTTVerificationCircuit, FalconStatement) {
    let keypair = KeyPair::keygen();
    let msg = "testing message".as_bytes();
    let sig = keypair.secret_key.sign_with_seed("test seed".as_ref(), msg);
    assert!(keypair.public_key.verify(msg, &sig));
    (
        FalconNTTVerificationCircuit::build_circuit(keypair.public_key, msg.to_vec(), sig),
        FalconStatement::new(keypair.public_key, msg.to_vec(), sig),
    )
}

/// 1 + 2: cs.num_*() and cs.to_matrices() against frcs_shape_get / frcs_get_matrix, entry by entry.
/// Decides the row-form questions of SURVEY.md App. F.1 (Boolean::or arm bindings, enforce_equal sign).
#[test]
fn matrices_equal_arkworks() {
    let (circuit, _) = statement();
    let cs = ConstraintSystem::<Fr>::new_ref();
    cs.set_optimization_goal(OptimizationGoal::Constraints); // what ark-groth16 sets
    circuit.generate_constraints(cs.clone()).unwrap();
    cs.finalize();
    assert!(cs.is_satisfied().unwrap());
    let m = cs.to_matrices().unwrap();

    let mut ctx = std::ptr::null_mut();
    assert_eq!(unsafe { frcs_ctx_create(LOG_N as u32, 0, 0, &mut ctx) }, FRCS_OK);
    let mut sh = frcs_shape::default();
    assert_eq!(unsafe { frcs_shape_get(ctx, &mut sh) }, FRCS_OK);
    assert_eq!(sh.n_instance as usize, cs.num_instance_variables());
    assert_eq!(sh.n_witness as usize, cs.num_witness_variables());
    assert_eq!(sh.n_constraints as usize, cs.num_constraints());
    for (which, (rows, nnz)) in [(&m.a, sh.nnz_a), (&m.b, sh.nnz_b), (&m.c, sh.nnz_c)].iter().enumerate() {
        let mut row_ptr = vec![0u32; sh.n_constraints as usize + 1];
        let mut col = vec![0u32; *nnz as usize];
        let mut val = vec![0u64; 4 * *nnz as usize];
        assert_eq!(
            unsafe { frcs_get_matrix(ctx, which as i32, row_ptr.as_mut_ptr(), col.as_mut_ptr(), val.as_mut_ptr()) },
            FRCS_OK
        );
        assert_eq!(rows.iter().map(|r| r.len()).sum::<usize>(), *nnz as usize, "nnz of matrix {}", which);
        for (i, row) in rows.iter().enumerate() {
            // arkworks keeps a row as Vec<(F, usize)> in LC order; the library sorts by column: compare as sets
            let mut want: Vec<(usize, [u64; 4])> = row.iter().map(|(c, j)| (*j, c.0 .0)).collect();
            want.sort();
            let (lo, hi) = (row_ptr[i] as usize, row_ptr[i + 1] as usize);
            let got: Vec<(usize, [u64; 4])> =
                (lo..hi).map(|k| (col[k] as usize, [val[4 * k], val[4 * k + 1], val[4 * k + 2], val[4 * k + 3]])).collect();
            assert_eq!(got, want, "matrix {} row {}", which, i);
        }
    }
    unsafe { frcs_ctx_destroy(ctx) };
}

/// 3: the full assignment z = instance ++ witness, limb for limb.
#[test]
fn assignment_equals_arkworks() {
    let (circuit, stmt) = statement();
    let cs = ConstraintSystem::<Fr>::new_ref();
    circuit.generate_constraints(cs.clone()).unwrap();
    let inner = cs.borrow().unwrap();
    let want: Vec<u64> = inner
        .instance_assignment
        .iter()
        .chain(inner.witness_assignment.iter())
        .flat_map(|x| x.0 .0.to_vec())
        .collect();

    let mut ctx = std::ptr::null_mut();
    assert_eq!(unsafe { frcs_ctx_create(LOG_N as u32, 0, 0, &mut ctx) }, FRCS_OK);
    let (sig, pk, hm) = stmt.polynomials();
    let mut z = vec![0u64; want.len()];
    let mut status = 0i32;
    assert_eq!(
        unsafe {
            frcs_witness_batch(ctx, 1, sig.coeff().as_ptr(), pk.coeff().as_ptr(), hm.coeff().as_ptr(), z.as_mut_ptr(), &mut status)
        },
        FRCS_OK
    );
    assert_eq!(status, FRCS_OK);
    assert_eq!(z, want);
    // public inputs in the order pok_sig.rs:33-44 rebuilds them: pk_ntt then hm_ntt
    let pk_ntt = NTTPolynomial::from(&pk);
    let hm_ntt = NTTPolynomial::from(&hm);
    for i in 0..N {
        assert_eq!(&z[4 * (1 + i)..4 * (2 + i)], &Fr::from(pk_ntt.coeff()[i]).0 .0[..]);
        assert_eq!(&z[4 * (1 + N + i)..4 * (2 + N + i)], &Fr::from(hm_ntt.coeff()[i]).0 .0[..]);
    }
    unsafe { frcs_ctx_destroy(ctx) };
}

/// 4: create_proof(circuit, &pk, r, s) byte-identical (192 compressed bytes), and verify_proof accepts it.
#[test]
fn proof_bytes_equal_arkworks() {
    let (circuit, stmt) = statement();
    let mut rng = ChaCha20Rng::from_seed([0u8; 32]); // examples/pok_sig.rs:13
    let pp = generate_random_parameters::<Bls12_381, _, _>(circuit.clone(), &mut rng).unwrap();
    let (r, s) = (Fr::rand(&mut rng), Fr::rand(&mut rng));
    let want = create_proof(circuit, &pp, r, s).unwrap();
    let gpu = GpuProver::new(&pp, 0).unwrap();
    let got = gpu.create_proof(&stmt, r, s).unwrap();
    let (mut wb, mut gb) = (Vec::new(), Vec::new());
    want.serialize(&mut wb).unwrap();
    got.serialize(&mut gb).unwrap();
    assert_eq!(gb, wb);
    let (_, pk, hm) = stmt.polynomials();
    let inputs: Vec<Fr> = NTTPolynomial::from(&pk)
        .coeff()
        .iter()
        .chain(NTTPolynomial::from(&hm).coeff().iter())
        .map(|x| Fr::from(*x))
        .collect();
    assert!(verify_proof(&prepare_verifying_key(&pp.vk), &got, &inputs).unwrap());
}

/// Gadget entry point on the GPU against the reference's own KATs (gadgets/arithmetics.rs:346-361).
#[test]
fn mod_q_known_answers() {
    let mut ctx = std::ptr::null_mut();
    assert_eq!(unsafe { frcs_ctx_create(LOG_N as u32, 0, 0, &mut ctx) }, FRCS_OK);
    for (a, b, sat) in [(6u64, 6u64, true), (0, 0, true), (12289, 0, true), (12290, 1, true), (6, 7, false), (5, 12288, false)] {
        let (av, bv) = (Fr::from(a).0 .0, Fr::from(b).0 .0);
        let mut wit = [0u64; 30 * 4];
        let (mut fu, mut st) = (0i64, 0i32);
        assert_eq!(
            unsafe { frcs_gadget_mod_q(ctx, 1, av.as_ptr(), bv.as_ptr(), wit.as_mut_ptr(), &mut fu, &mut st) },
            FRCS_OK
        );
        assert_eq!(fu == -1, sat, "mod_q({}) == {}", a, b);
    }
    unsafe { frcs_ctx_destroy(ctx) };
}
