// ORACLE — TEST INFRASTRUCTURE ONLY.  ctypes-facing C API of the CPU oracle.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs may load this library; the product (falcon_r1cs_b200/) never does.
#include <chrono>
#include <cstdio>

#include "gadgets.hpp"
#include "groth16.hpp"

using namespace orc;

namespace {

struct Circuit {
  int logn, kind, n;
  Matrices m;
  Domain* dom = nullptr;
  ~Circuit() { delete dom; }
};
struct PkHandle {
  ProvingKey pk;
  Trapdoor td;
};

void synth(Gadgets& g, int kind, const std::vector<uint32_t>& sig, const std::vector<uint32_t>& pk,
           const std::vector<uint32_t>& hm) {
  if (kind == KIND_NTT)
    g.falcon_ntt_circuit(sig, pk, hm);
  else if (kind == KIND_SCHOOLBOOK)
    g.falcon_schoolbook_circuit(sig, pk, hm);
  else if (kind == KIND_DUAL_NTT)
    g.falcon_dual_ntt_circuit(sig, pk, hm);
  else
    throw std::runtime_error("unknown circuit kind");
}
void store_fr(uint64_t* out, const Fr& x) { memcpy(out, x.v, 32); }
G1A load_g1(const uint64_t* p) {
  G1A a;
  a.x = Fq::from_raw(p);
  a.y = Fq::from_raw(p + 6);
  a.inf = a.x.is_zero() && a.y.is_zero();
  return a;
}
void store_g1(uint64_t* p, const G1A& a) {
  if (a.inf) {
    memset(p, 0, 96);
    return;
  }
  memcpy(p, a.x.v, 48);
  memcpy(p + 6, a.y.v, 48);
}
G2A load_g2(const uint64_t* p) {
  G2A a;
  a.x = {Fq::from_raw(p), Fq::from_raw(p + 6)};
  a.y = {Fq::from_raw(p + 12), Fq::from_raw(p + 18)};
  a.inf = a.x.is_zero() && a.y.is_zero();
  return a;
}
void store_g2(uint64_t* p, const G2A& a) {
  if (a.inf) {
    memset(p, 0, 192);
    return;
  }
  memcpy(p, a.x.c0.v, 48);
  memcpy(p + 6, a.x.c1.v, 48);
  memcpy(p + 12, a.y.c0.v, 48);
  memcpy(p + 18, a.y.c1.v, 48);
}

}  // namespace

extern "C" {

// Build the circuit in Setup mode (as ark-groth16's generator does) and keep A/B/C.
void* orc_circuit_new(int logn, int kind) {
  try {
    Circuit* c = new Circuit;
    c->logn = logn;
    c->kind = kind;
    c->n = 1 << logn;
    ConstraintSystem cs;
    cs.mode = MODE_SETUP;
    Gadgets g(&cs, logn);
    std::vector<uint32_t> dummy(c->n, 0);
    synth(g, kind, dummy, dummy, dummy);
    c->m.num_instance = cs.num_instance;
    c->m.num_witness = cs.num_witness;
    c->m.num_constraints = cs.num_constraints;
    c->m.a = cs.to_csr(cs.a);
    c->m.b = cs.to_csr(cs.b);
    c->m.c = cs.to_csr(cs.c);
    c->dom = new Domain((size_t)cs.num_constraints + cs.num_instance);
    return c;
  } catch (std::exception& e) {
    fprintf(stderr, "orc_circuit_new: %s\n", e.what());
    return nullptr;
  }
}
void orc_circuit_free(void* h) { delete (Circuit*)h; }
// out: n_inst, n_wit, n_cons, nnzA, nnzB, nnzC, domain_log2
void orc_shape(void* h, uint64_t* out) {
  Circuit* c = (Circuit*)h;
  out[0] = c->m.num_instance;
  out[1] = c->m.num_witness;
  out[2] = c->m.num_constraints;
  out[3] = c->m.a.col.size();
  out[4] = c->m.b.col.size();
  out[5] = c->m.c.col.size();
  out[6] = c->dom->log_size;
}
void orc_get_csr(void* h, int which, uint32_t* row_ptr, uint32_t* col, uint64_t* val) {
  Circuit* c = (Circuit*)h;
  const CSR& m = which == 0 ? c->m.a : which == 1 ? c->m.b : c->m.c;
  memcpy(row_ptr, m.row_ptr.data(), m.row_ptr.size() * 4);
  memcpy(col, m.col.data(), m.col.size() * 4);
  for (size_t i = 0; i < m.val.size(); i++) store_fr(val + 4 * i, m.val[i]);
}
// generate_constraints in Prove mode.  z_out = instance ++ witness (Montgomery limbs).
// Returns the status (0, or the code of the first reference panic site hit).
// first_unsat: only filled when construct_matrices != 0 (cs.which_is_unsatisfied()).
int orc_witness(void* h, const uint16_t* sig, const uint16_t* pk, const uint16_t* hm, uint64_t* z_out,
                int construct_matrices, int panic_on_range, int64_t* first_unsat) {
  Circuit* c = (Circuit*)h;
  try {
    ConstraintSystem cs;
    cs.mode = MODE_PROVE;
    cs.construct_matrices = construct_matrices != 0;
    Gadgets g(&cs, c->logn, panic_on_range != 0);
    std::vector<uint32_t> s(sig, sig + c->n), p(pk, pk + c->n), m(hm, hm + c->n);
    synth(g, c->kind, s, p, m);
    if (cs.num_instance != c->m.num_instance || cs.num_witness != c->m.num_witness) return -100;
    for (size_t i = 0; i < cs.instance.size(); i++) store_fr(z_out + 4 * i, cs.instance[i]);
    for (size_t i = 0; i < cs.witness.size(); i++) store_fr(z_out + 4 * (cs.num_instance + i), cs.witness[i]);
    if (first_unsat) *first_unsat = construct_matrices ? cs.first_unsatisfied() : -2;
    return cs.status;
  } catch (std::exception& e) {
    fprintf(stderr, "orc_witness: %s\n", e.what());
    return -101;
  }
}
// A.z, B.z, C.z with the stored matrices; first_unsat = first row with az*bz != cz, else -1
void orc_r1cs_eval(void* h, const uint64_t* z, uint64_t* az, uint64_t* bz, uint64_t* cz, int64_t* first_unsat) {
  Circuit* c = (Circuit*)h;
  const Fr* zz = (const Fr*)z;
  int64_t bad = -1;
  size_t nc = c->m.num_constraints;
  std::vector<uint8_t> viol(nc, 0);
#pragma omp parallel for schedule(dynamic, 256)
  for (size_t i = 0; i < nc; i++) {
    Fr a = csr_row_dot(c->m.a, i, zz), b = csr_row_dot(c->m.b, i, zz), cc = csr_row_dot(c->m.c, i, zz);
    if (az) store_fr(az + 4 * i, a);
    if (bz) store_fr(bz + 4 * i, b);
    if (cz) store_fr(cz + 4 * i, cc);
    viol[i] = a * b != cc;
  }
  for (size_t i = 0; i < nc; i++)
    if (viol[i]) {
      bad = (int64_t)i;
      break;
    }
  if (first_unsat) *first_unsat = bad;
}
void orc_witness_map(void* h, const uint64_t* z, uint64_t* h_out) {
  Circuit* c = (Circuit*)h;
  std::vector<Fr> r = witness_map(c->m, *c->dom, (const Fr*)z);
  memcpy(h_out, r.data(), r.size() * 32);
}
// Domain primitives exposed for NTT parity tests.  op: 0 fft, 1 ifft, 2 coset_fft, 3 coset_ifft
void orc_domain_op(uint32_t log_size, int op, uint64_t* data) {
  Domain d((size_t)1 << log_size);
  std::vector<Fr> v((const Fr*)data, (const Fr*)data + d.size);
  if (op == 0) d.fft_in_place(v);
  if (op == 1) d.ifft_in_place(v);
  if (op == 2) d.coset_fft_in_place(v);
  if (op == 3) d.coset_ifft_in_place(v);
  memcpy(data, v.data(), d.size * 32);
}

void* orc_setup(void* h, uint64_t seed) {
  Circuit* c = (Circuit*)h;
  PkHandle* p = new PkHandle;
  p->pk = setup(c->m, *c->dom, seed, &p->td);
  return p;
}
// setup from explicit toxic waste: 7 x 4 u64 Montgomery (alpha, beta, gamma, delta, tau, g1_scalar, g2_scalar)
void* orc_setup_trapdoor(void* h, const uint64_t* td7) {
  Circuit* c = (Circuit*)h;
  PkHandle* p = new PkHandle;
  Trapdoor t;
  Fr* f[7] = {&t.alpha, &t.beta, &t.gamma, &t.delta, &t.tau, &t.g1_scalar, &t.g2_scalar};
  for (int k = 0; k < 7; k++) memcpy(f[k]->v, td7 + 4 * k, 32);
  p->pk = setup_from_trapdoor(c->m, *c->dom, t, &p->td);
  return p;
}
void orc_pk_free(void* p) { delete (PkHandle*)p; }
// the toxic waste of a setup (Montgomery limbs): alpha, beta, gamma, delta, tau, g1_scalar, g2_scalar
void orc_pk_trapdoor(void* p, uint64_t* out) {
  const Trapdoor& t = ((PkHandle*)p)->td;
  const Fr* f[7] = {&t.alpha, &t.beta, &t.gamma, &t.delta, &t.tau, &t.g1_scalar, &t.g2_scalar};
  for (int k = 0; k < 7; k++) memcpy(out + 4 * k, f[k]->v, 32);
}
// which: 0 a_query, 1 b_g1_query, 2 b_g2_query, 3 h_query, 4 l_query, 5 gamma_abc_g1,
//        6 [alpha_g1, beta_g1, delta_g1], 7 [beta_g2, delta_g2, gamma_g2]
uint64_t orc_pk_len(void* p, int which) {
  ProvingKey& pk = ((PkHandle*)p)->pk;
  switch (which) {
    case 0: return pk.a_query.size();
    case 1: return pk.b_g1_query.size();
    case 2: return pk.b_g2_query.size();
    case 3: return pk.h_query.size();
    case 4: return pk.l_query.size();
    case 5: return pk.gamma_abc_g1.size();
    case 6: return 3;
    case 7: return 3;
  }
  return 0;
}
void orc_pk_export(void* p, int which, uint64_t* out) {
  ProvingKey& pk = ((PkHandle*)p)->pk;
  auto g1v = [&](const std::vector<G1A>& v) {
    for (size_t i = 0; i < v.size(); i++) store_g1(out + 12 * i, v[i]);
  };
  switch (which) {
    case 0: g1v(pk.a_query); break;
    case 1: g1v(pk.b_g1_query); break;
    case 2:
      for (size_t i = 0; i < pk.b_g2_query.size(); i++) store_g2(out + 24 * i, pk.b_g2_query[i]);
      break;
    case 3: g1v(pk.h_query); break;
    case 4: g1v(pk.l_query); break;
    case 5: g1v(pk.gamma_abc_g1); break;
    case 6:
      store_g1(out, pk.alpha_g1);
      store_g1(out + 12, pk.beta_g1);
      store_g1(out + 24, pk.delta_g1);
      break;
    case 7:
      store_g2(out, pk.beta_g2);
      store_g2(out + 24, pk.delta_g2);
      store_g2(out + 48, pk.gamma_g2);
      break;
  }
}
// scalars: canonical integers (into_repr()), 4 x u64 each
void orc_msm_g1(const uint64_t* bases, const uint64_t* scalars, uint64_t n, uint64_t* out) {
  std::vector<G1A> b(n);
  for (size_t i = 0; i < n; i++) b[i] = load_g1(bases + 12 * i);
  store_g1(out, msm_pippenger<Fq>(b.data(), (const Big<4>*)scalars, n).to_affine());
}
void orc_msm_g2(const uint64_t* bases, const uint64_t* scalars, uint64_t n, uint64_t* out) {
  std::vector<G2A> b(n);
  for (size_t i = 0; i < n; i++) b[i] = load_g2(bases + 24 * i);
  store_g2(out, msm_pippenger<Fq2>(b.data(), (const Big<4>*)scalars, n).to_affine());
}
// Montgomery <-> canonical helpers for tests
void orc_fr_from_canonical(const uint64_t* in, uint64_t* out, uint64_t n) {
  for (size_t i = 0; i < n; i++) store_fr(out + 4 * i, Fr::from_big(*(const Big<4>*)(in + 4 * i)));
}
void orc_fr_to_canonical(const uint64_t* in, uint64_t* out, uint64_t n) {
  for (size_t i = 0; i < n; i++) {
    Big<4> b = ((const Fr*)in)[i].to_big();
    memcpy(out + 4 * i, b.l, 32);
  }
}
// proof_affine: a (12 u64) | b (24 u64) | c (12 u64), Montgomery; compressed: 192 bytes
void orc_prove(void* h, void* p, const uint64_t* z, const uint64_t* r, const uint64_t* s, uint64_t* proof_affine,
               uint8_t* compressed) {
  Circuit* c = (Circuit*)h;
  Proof pf = create_proof(c->m, *c->dom, ((PkHandle*)p)->pk, (const Fr*)z, Fr::from_raw(r), Fr::from_raw(s));
  if (proof_affine) {
    store_g1(proof_affine, pf.a);
    store_g2(proof_affine + 12, pf.b);
    store_g1(proof_affine + 36, pf.c);
  }
  if (compressed) {
    memset(compressed, 0, 192);
    ser_proof(pf, compressed);
  }
}
int orc_verify_trapdoor(void* h, void* p, const uint64_t* z, const uint64_t* r, const uint64_t* s,
                        const uint64_t* proof_affine) {
  Circuit* c = (Circuit*)h;
  Proof pf{load_g1(proof_affine), load_g2(proof_affine + 12), load_g1(proof_affine + 36)};
  return verify_with_trapdoor(c->m, *c->dom, ((PkHandle*)p)->td, (const Fr*)z, Fr::from_raw(r), Fr::from_raw(s), pf);
}
void orc_compress_proof(const uint64_t* proof_affine, uint8_t* out) {
  Proof pf{load_g1(proof_affine), load_g2(proof_affine + 12), load_g1(proof_affine + 36)};
  memset(out, 0, 192);
  ser_proof(pf, out);
}
int orc_point_check(const uint64_t* p, int g2) {
  return g2 ? g2_on_curve(load_g2(p)) : g1_on_curve(load_g1(p));
}
// scalar multiplication of a single point (tests of the host curve code)
void orc_g1_mul(const uint64_t* p, const uint64_t* k_canonical, uint64_t* out) {
  store_g1(out, G1J::from_affine(load_g1(p)).mul(*(const Big<4>*)k_canonical).to_affine());
}
void orc_g2_mul(const uint64_t* p, const uint64_t* k_canonical, uint64_t* out) {
  store_g2(out, G2J::from_affine(load_g2(p)).mul(*(const Big<4>*)k_canonical).to_affine());
}
void orc_generators(uint64_t* g1, uint64_t* g2) {
  store_g1(g1, g1_generator());
  store_g2(g2, g2_generator());
}

// ---- gadget known-answer harness (mirrors the reference's #[cfg(test)] macros) ----
// which: 0 mod_q(a)=exp  1 add_mod(a,b)=exp  2 mul_mod(a,b)=exp  3 enforce_less_than_q(a)
//        4 enforce_less_than_norm_bound(a)  5 is_less_than_6144(a).enforce_equal(TRUE)
//        6 inner_product_mod(a[0..k), b[0..k)) = exp   7 enforce_less_than_1024(a)
// in: canonical u64 inputs.  Returns is_satisfied; *value_ok = (gadget value == exp);
// counts = {num_instance, num_witness, num_constraints}.
// z_out (may be NULL): the whole witness assignment, cs.num_witness x 4 Montgomery limbs; first_unsat (may be NULL):
// cs.which_is_unsatisfied() or -1.
int orc_kat_z(int which, int logn, const uint64_t* in, int n_in, uint64_t expected, int* value_ok, uint64_t* counts,
              uint64_t* z_out, int64_t* first_unsat);
int orc_kat(int which, int logn, const uint64_t* in, int n_in, uint64_t expected, int* value_ok, uint64_t* counts) {
  return orc_kat_z(which, logn, in, n_in, expected, value_ok, counts, nullptr, nullptr);
}
int orc_kat_z(int which, int logn, const uint64_t* in, int n_in, uint64_t expected, int* value_ok, uint64_t* counts,
              uint64_t* z_out, int64_t* first_unsat) {
  ConstraintSystem cs;
  Gadgets g(&cs, logn, /*panic_on_range=*/false);  // #[cfg(test)] build: range panics compiled out
  FpVar q = FpVar::constant(Fr::from_u64(FALCON_Q));
  Fr exp = Fr::from_u64(expected);
  FpVar out;
  bool has_out = false;
  if (which == 0 || which == 1 || which == 2) {
    FpVar a = g.c.new_witness(Fr::from_u64(in[0]));
    if (which == 0)
      out = g.mod_q(a, q);
    else {
      FpVar b = g.c.new_witness(Fr::from_u64(in[1]));
      out = which == 1 ? g.add_mod(a, b, q) : g.mul_mod(a, b, q);
    }
    has_out = true;
  } else if (which == 3) {
    g.enforce_less_than_q(g.c.new_witness(Fr::from_u64(in[0])));
  } else if (which == 4) {
    g.enforce_less_than_norm_bound(g.c.new_witness(Fr::from_u64(in[0])));
  } else if (which == 5) {
    // test_range_proof_half_q (range_proofs.rs:506-522): is_less.enforce_equal(TRUE)
    Boolean r = g.is_less_than_6144(g.c.new_witness(Fr::from_u64(in[0])));
    g.b.enforce_true(r);
    if (value_ok) *value_ok = r.value() == (expected != 0);
  } else if (which == 6) {
    int k = n_in / 2;
    std::vector<FpVar> a, b;
    for (int i = 0; i < k; i++) a.push_back(g.c.new_witness(Fr::from_u64(in[i])));
    for (int i = 0; i < k; i++) b.push_back(g.c.new_witness(Fr::from_u64(in[k + i])));
    g.n = k;
    out = g.inner_product_mod(a, b.data(), q);
    has_out = true;
  } else if (which == 7) {
    g.enforce_less_than_1024(g.c.new_witness(Fr::from_u64(in[0])));
  }
  if (has_out) {
    FpVar e = g.c.new_witness(exp);
    g.c.enforce_equal(out, e);
    if (value_ok) *value_ok = out.value() == exp;
  }
  if (counts) {
    counts[0] = cs.num_instance;
    counts[1] = cs.num_witness;
    counts[2] = cs.num_constraints;
  }
  if (z_out)
    for (size_t i = 0; i < cs.witness.size(); i++) store_fr(z_out + 4 * i, cs.witness[i]);
  const int64_t fu = cs.first_unsatisfied();
  if (first_unsat) *first_unsat = fu;
  return fu < 0;
}
// ntt_circuit on a polynomial (test_ntt_mul_circuit, poly.rs:252-301): writes the N
// output values (canonical, as u16) and returns is_satisfied; counts as above.
int orc_kat_ntt_z(int logn, const uint16_t* poly, uint16_t* out, uint64_t* counts, uint64_t* z_out);
int orc_kat_ntt(int logn, const uint16_t* poly, uint16_t* out, uint64_t* counts) {
  return orc_kat_ntt_z(logn, poly, out, counts, nullptr);
}
// z_out (may be NULL): the gadget's 29N witnesses (after the N input witnesses), Montgomery limbs
int orc_kat_ntt_z(int logn, const uint16_t* poly, uint16_t* out, uint64_t* counts, uint64_t* z_out) {
  ConstraintSystem cs;
  Gadgets g(&cs, logn, false);
  std::vector<uint32_t> p(poly, poly + g.n);
  std::vector<FpVar> vars = g.alloc_vars(p, false);
  uint32_t ni = cs.num_instance, nw = cs.num_witness, nc = cs.num_constraints;
  std::vector<FpVar> r = g.ntt_circuit(vars, g.const_q_power_vars(), g.ntt_param_var());
  for (int i = 0; i < g.n; i++) out[i] = (uint16_t)r[i].value().to_big().l[0];
  if (counts) {
    counts[0] = cs.num_instance - ni;
    counts[1] = cs.num_witness - nw;
    counts[2] = cs.num_constraints - nc;
  }
  if (z_out)
    for (size_t i = nw; i < cs.witness.size(); i++) store_fr(z_out + 4 * (i - nw), cs.witness[i]);
  return cs.first_unsatisfied() < 0;
}
void orc_ntt_clear(int logn, const uint16_t* in, uint16_t* out) {
  std::vector<uint32_t> p(in, in + (1 << logn));
  std::vector<uint32_t> r = ntt_clear(p, logn);
  for (size_t i = 0; i < r.size(); i++) out[i] = (uint16_t)r[i];
}
int orc_num_threads() { return omp_get_max_threads(); }
void orc_set_num_threads(int n) { omp_set_num_threads(n); }

}  // extern "C"
