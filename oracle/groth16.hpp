// ORACLE — TEST INFRASTRUCTURE ONLY.  CPU restatement of the parts of
// ark-poly 0.3.0 (Radix2EvaluationDomain), ark-ec 0.3.0 (VariableBaseMSM) and
// ark-groth16 0.3.0 (R1CStoQAP::witness_map, create_proof, generate_parameters)
// reached from examples/pok_sig.rs:30-32.  All [EXT]; restated from SURVEY.md
// App. B.4-B.7.  "parity unpinned" beyond verify_proof == true (pok_sig.rs:47):
// the reference holds no golden h / MSM / proof bytes (keygen is unseeded).
#pragma once
#include <omp.h>

#include <vector>

#include "ec.hpp"
#include "r1cs.hpp"

namespace orc {

static inline Fr fr_two_adic_root() {  // 7^((r-1)/2^32)
  Big<4> e;
  Big<4> rm1;
  memcpy(rm1.l, FrTag::MOD, sizeof rm1.l);
  rm1.l[0] -= 1;
  // (r-1) >> 32
  for (int i = 0; i < 4; i++) e.l[i] = (rm1.l[i] >> 32) | (i < 3 ? rm1.l[i + 1] << 32 : 0);
  return Fr::from_u64(7).pow(e);
}

// ark-poly 0.3.0 Radix2EvaluationDomain
struct Domain {
  uint32_t log_size;
  size_t size;
  Fr group_gen, group_gen_inv, size_inv, gen, gen_inv;  // gen = multiplicative generator 7 (coset offset)
  std::vector<Fr> tw, tw_inv;                           // w^i, w^-i for i < size/2

  explicit Domain(size_t num_coeffs) {
    log_size = 0;
    while (((size_t)1 << log_size) < num_coeffs) log_size++;
    size = (size_t)1 << log_size;
    group_gen = fr_two_adic_root();
    for (uint32_t i = log_size; i < 32; i++) group_gen = group_gen.square();
    group_gen_inv = group_gen.inverse();
    size_inv = Fr::from_u64(size).inverse();
    gen = Fr::from_u64(7);
    gen_inv = gen.inverse();
    tw.resize(size / 2);
    tw_inv.resize(size / 2);
    Fr a = Fr::one(), b = Fr::one();
    for (size_t i = 0; i < size / 2; i++) {
      tw[i] = a;
      tw_inv[i] = b;
      a *= group_gen;
      b *= group_gen_inv;
    }
  }
  void fft_core(std::vector<Fr>& x, const std::vector<Fr>& w) const {
    size_t n = size;
    for (size_t i = 0; i < n; i++) {  // bit reversal
      size_t j = 0;
      for (uint32_t b = 0; b < log_size; b++) j |= ((i >> b) & 1) << (log_size - 1 - b);
      if (i < j) std::swap(x[i], x[j]);
    }
    for (size_t len = 2; len <= n; len <<= 1) {
      size_t half = len / 2, step = n / len;
#pragma omp parallel for schedule(static)
      for (size_t k = 0; k < n / 2; k++) {
        size_t blk = k / half, j = k % half;
        size_t i0 = blk * len + j, i1 = i0 + half;
        Fr t = x[i1] * w[j * step];
        x[i1] = x[i0] - t;
        x[i0] = x[i0] + t;
      }
    }
  }
  void fft_in_place(std::vector<Fr>& x) const { fft_core(x, tw); }
  void ifft_in_place(std::vector<Fr>& x) const {
    fft_core(x, tw_inv);
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < size; i++) x[i] *= size_inv;
  }
  static void distribute_powers(std::vector<Fr>& x, const Fr& g) {
    Fr p = Fr::one();
    for (auto& e : x) {
      e *= p;
      p *= g;
    }
  }
  void coset_fft_in_place(std::vector<Fr>& x) const {
    distribute_powers(x, gen);
    fft_in_place(x);
  }
  void coset_ifft_in_place(std::vector<Fr>& x) const {
    ifft_in_place(x);
    distribute_powers(x, gen_inv);
  }
  Fr evaluate_vanishing_polynomial(const Fr& tau) const { return tau.pow_u64(size) - Fr::one(); }
  void divide_by_vanishing_poly_on_coset_in_place(std::vector<Fr>& x) const {
    Fr i = evaluate_vanishing_polynomial(gen).inverse();
#pragma omp parallel for schedule(static)
    for (size_t k = 0; k < x.size(); k++) x[k] *= i;
  }
  // L_i(tau) = Z(tau) * w^i / (n * (tau - w^i)), tau outside the domain
  std::vector<Fr> evaluate_all_lagrange_coefficients(const Fr& tau) const {
    Fr z = evaluate_vanishing_polynomial(tau) * size_inv;
    std::vector<Fr> den(size), u(size);
    Fr w = Fr::one();
    for (size_t i = 0; i < size; i++) {
      den[i] = tau - w;
      u[i] = w;
      w *= group_gen;
    }
    // batch inversion
    std::vector<Fr> pre(size);
    Fr acc = Fr::one();
    for (size_t i = 0; i < size; i++) {
      pre[i] = acc;
      acc *= den[i];
    }
    Fr inv = acc.inverse();
    for (size_t i = size; i-- > 0;) {
      Fr di = inv * pre[i];
      inv *= den[i];
      u[i] = u[i] * z * di;
    }
    return u;
  }
};

static inline Fr csr_row_dot(const CSR& m, size_t row, const Fr* z) {  // evaluate_constraint
  Fr s = Fr::zero();
  for (uint32_t k = m.row_ptr[row]; k < m.row_ptr[row + 1]; k++) s += m.val[k] * z[m.col[k]];
  return s;
}

struct Matrices {
  uint32_t num_instance, num_witness, num_constraints;
  CSR a, b, c;
};

// ark-groth16 0.3.0 r1cs_to_qap.rs: R1CStoQAP::witness_map.  z = instance ++ witness.
static inline std::vector<Fr> witness_map(const Matrices& m, const Domain& d, const Fr* z) {
  size_t n = d.size, nc = m.num_constraints, ni = m.num_instance;
  std::vector<Fr> a(n, Fr::zero()), b(n, Fr::zero()), c(n, Fr::zero());
#pragma omp parallel for schedule(dynamic, 256)
  for (size_t i = 0; i < nc; i++) {
    a[i] = csr_row_dot(m.a, i, z);
    b[i] = csr_row_dot(m.b, i, z);
  }
  for (size_t i = 0; i < ni; i++) a[nc + i] = z[i];
  d.ifft_in_place(a);
  d.ifft_in_place(b);
  d.coset_fft_in_place(a);
  d.coset_fft_in_place(b);
  std::vector<Fr> ab(n);
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < n; i++) ab[i] = a[i] * b[i];
#pragma omp parallel for schedule(dynamic, 256)
  for (size_t i = 0; i < nc; i++) c[i] = csr_row_dot(m.c, i, z);
  d.ifft_in_place(c);
  d.coset_fft_in_place(c);
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < n; i++) ab[i] -= c[i];
  d.divide_by_vanishing_poly_on_coset_in_place(ab);
  d.coset_ifft_in_place(ab);
  return ab;
}

// ark-ec 0.3.0 VariableBaseMSM::multi_scalar_mul (Pippenger)
template <class F>
static Jac<F> msm_pippenger(const Affine<F>* bases, const Big<4>* scalars, size_t size) {
  auto log2c = [](size_t x) {
    if (x == 0) return 0u;
    uint32_t l = 0;
    while (((size_t)1 << l) < x) l++;
    return l;
  };
  size_t c = size < 32 ? 3 : (log2c(size) * 69 / 100) + 2;
  const int num_bits = 255;
  Big<4> one;
  memset(one.l, 0, sizeof one.l);
  one.l[0] = 1;
  std::vector<size_t> starts;
  for (size_t w = 0; w < (size_t)num_bits; w += c) starts.push_back(w);
  std::vector<Jac<F>> window_sums(starts.size());
#pragma omp parallel for schedule(dynamic, 1)
  for (size_t wi = 0; wi < starts.size(); wi++) {
    size_t w_start = starts[wi];
    Jac<F> res = Jac<F>::infinity();
    std::vector<Jac<F>> buckets(((size_t)1 << c) - 1, Jac<F>::infinity());
    for (size_t i = 0; i < size; i++) {
      const Big<4>& s = scalars[i];
      if (s.is_zero()) continue;
      if (s == one) {
        if (w_start == 0) res = res.add_mixed(bases[i]);
      } else {
        // scalar.divn(w_start); scalar.as_ref()[0] % (1 << c)
        size_t limb = w_start / 64, off = w_start % 64;
        uint64_t v = s.l[limb] >> off;
        if (off && limb + 1 < 4) v |= s.l[limb + 1] << (64 - off);
        v &= (((uint64_t)1 << c) - 1);
        if (v) buckets[v - 1] = buckets[v - 1].add_mixed(bases[i]);
      }
    }
    Jac<F> running = Jac<F>::infinity();
    for (size_t k = buckets.size(); k-- > 0;) {
      running = running.add(buckets[k]);
      res = res.add(running);
    }
    window_sums[wi] = res;
  }
  Jac<F> total = Jac<F>::infinity();
  for (size_t wi = window_sums.size(); wi-- > 1;) {
    total = total.add(window_sums[wi]);
    for (size_t k = 0; k < c; k++) total = total.dbl();
  }
  return window_sums[0].add(total);
}

struct ProvingKey {
  G1A alpha_g1, beta_g1, delta_g1;
  G2A beta_g2, delta_g2, gamma_g2;
  std::vector<G1A> a_query, b_g1_query, h_query, l_query, gamma_abc_g1;
  std::vector<G2A> b_g2_query;
};
struct Trapdoor {
  Fr alpha, beta, gamma, delta, tau, g1_scalar, g2_scalar;
};
struct Proof {
  G1A a;
  G2A b;
  G1A c;
};

// Fixed-base scalar multiplication table (8-bit windows)
template <class F>
struct FixedBase {
  std::vector<Affine<F>> tab;  // [32][256]
  explicit FixedBase(const Jac<F>& g) {
    std::vector<Jac<F>> t(32 * 256);
    Jac<F> base = g;
    for (int w = 0; w < 32; w++) {
      Jac<F> acc = Jac<F>::infinity();
      for (int k = 0; k < 256; k++) {
        t[w * 256 + k] = acc;
        acc = acc.add(base);
      }
      base = acc;  // 256 * previous base
    }
    tab = batch_to_affine(t);
  }
  Jac<F> mul(const Fr& s) const {
    Big<4> k = s.to_big();
    Jac<F> r = Jac<F>::infinity();
    for (int w = 0; w < 32; w++) {
      uint32_t d = (k.l[w / 8] >> (8 * (w % 8))) & 0xff;
      if (d) r = r.add_mixed(tab[w * 256 + d]);
    }
    return r;
  }
  std::vector<Affine<F>> mul_all(const std::vector<Fr>& s) const {
    std::vector<Jac<F>> out(s.size());
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < s.size(); i++) out[i] = mul(s[i]);
    // chunked batch normalisation, parallel
    std::vector<Affine<F>> aff(s.size());
    size_t chunk = 4096, nchunks = (s.size() + chunk - 1) / chunk;
#pragma omp parallel for schedule(static)
    for (size_t ci = 0; ci < nchunks; ci++) {
      size_t lo = ci * chunk, hi = std::min(s.size(), lo + chunk);
      std::vector<Jac<F>> part(out.begin() + lo, out.begin() + hi);
      std::vector<Affine<F>> r = batch_to_affine(part);
      std::copy(r.begin(), r.end(), aff.begin() + lo);
    }
    return aff;
  }
};

// splitmix64 stream -> Fr (value taken mod r by rejection on the top limb mask)
struct Rng {
  uint64_t s;
  uint64_t next() {
    uint64_t z = (s += 0x9e3779b97f4a7c15ULL);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
  }
  Fr fr() {
    for (;;) {
      Big<4> b;
      for (int i = 0; i < 4; i++) b.l[i] = next();
      b.l[3] >>= 1;
      if (cmp_n<4>(b.l, FrTag::MOD) < 0) return Fr::from_big(b);
    }
  }
};

// ark-groth16 0.3.0 generator.rs: generate_parameters, with the toxic waste kept
// (seeded) so proofs can also be checked in the exponent.
static inline ProvingKey setup_from_trapdoor(const Matrices& m, const Domain& d, const Trapdoor& td, Trapdoor* td_out);
static inline ProvingKey setup(const Matrices& m, const Domain& d, uint64_t seed, Trapdoor* td_out) {
  Rng rng{seed};
  Trapdoor td;
  td.alpha = rng.fr();
  td.beta = rng.fr();
  td.gamma = rng.fr();
  td.delta = rng.fr();
  td.g1_scalar = rng.fr();
  td.g2_scalar = rng.fr();
  td.tau = rng.fr();  // domain.sample_element_outside_domain: w.h.p. outside
  return setup_from_trapdoor(m, d, td, td_out);
}
// the same with explicit toxic waste (parity tests of the product's frcs_setup)
static inline ProvingKey setup_from_trapdoor(const Matrices& m, const Domain& d, const Trapdoor& td, Trapdoor* td_out) {
  size_t ni = m.num_instance, nw = m.num_witness, nc = m.num_constraints, nv = ni + nw;
  Fr zt = d.evaluate_vanishing_polynomial(td.tau);
  std::vector<Fr> u = d.evaluate_all_lagrange_coefficients(td.tau);
  std::vector<Fr> a(nv, Fr::zero()), b(nv, Fr::zero()), c(nv, Fr::zero());
  for (size_t i = 0; i < ni; i++) a[i] = u[nc + i];
  for (size_t i = 0; i < nc; i++) {
    for (uint32_t k = m.a.row_ptr[i]; k < m.a.row_ptr[i + 1]; k++) a[m.a.col[k]] += u[i] * m.a.val[k];
    for (uint32_t k = m.b.row_ptr[i]; k < m.b.row_ptr[i + 1]; k++) b[m.b.col[k]] += u[i] * m.b.val[k];
    for (uint32_t k = m.c.row_ptr[i]; k < m.c.row_ptr[i + 1]; k++) c[m.c.col[k]] += u[i] * m.c.val[k];
  }
  Fr gamma_inv = td.gamma.inverse(), delta_inv = td.delta.inverse();
  std::vector<Fr> gabc(ni), l(nw);
  for (size_t i = 0; i < ni; i++) gabc[i] = (td.beta * a[i] + td.alpha * b[i] + c[i]) * gamma_inv;
  for (size_t i = 0; i < nw; i++) l[i] = (td.beta * a[ni + i] + td.alpha * b[ni + i] + c[ni + i]) * delta_inv;
  std::vector<Fr> hs(d.size - 1);
  Fr p = zt * delta_inv;
  for (size_t i = 0; i + 1 < d.size; i++) {
    hs[i] = p;
    p *= td.tau;
  }
  G1J g1 = G1J::from_affine(g1_generator()).mul(td.g1_scalar.to_big());
  G2J g2 = G2J::from_affine(g2_generator()).mul(td.g2_scalar.to_big());
  FixedBase<Fq> t1(g1);
  FixedBase<Fq2> t2(g2);
  ProvingKey pk;
  pk.alpha_g1 = t1.mul(td.alpha).to_affine();
  pk.beta_g1 = t1.mul(td.beta).to_affine();
  pk.delta_g1 = t1.mul(td.delta).to_affine();
  pk.beta_g2 = t2.mul(td.beta).to_affine();
  pk.delta_g2 = t2.mul(td.delta).to_affine();
  pk.gamma_g2 = t2.mul(td.gamma).to_affine();
  pk.a_query = t1.mul_all(a);
  pk.b_g1_query = t1.mul_all(b);
  pk.b_g2_query = t2.mul_all(b);
  pk.h_query = t1.mul_all(hs);
  pk.l_query = t1.mul_all(l);
  pk.gamma_abc_g1 = t1.mul_all(gabc);
  if (td_out) *td_out = td;
  return pk;
}

template <class F>
static Jac<F> calculate_coeff(const Jac<F>& initial, const std::vector<Affine<F>>& query, const Affine<F>& vk_param,
                              const std::vector<Big<4>>& assignment) {
  Jac<F> acc = msm_pippenger<F>(query.data() + 1, assignment.data(), std::min(query.size() - 1, assignment.size()));
  Jac<F> res = initial.add_mixed(query[0]);
  res = res.add(acc);
  return res.add_mixed(vk_param);
}

// ark-groth16 0.3.0 prover.rs: create_proof(circuit, pk, r, s) after synthesis;
// z = instance ++ witness, h = witness_map(z).
static inline Proof create_proof(const Matrices& m, const Domain& d, const ProvingKey& pk, const Fr* z, const Fr& r,
                                 const Fr& s) {
  std::vector<Fr> h = witness_map(m, d, z);
  std::vector<Big<4>> h_assign(h.size());
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < h.size(); i++) h_assign[i] = h[i].to_big();
  G1J h_acc = msm_pippenger<Fq>(pk.h_query.data(), h_assign.data(), std::min(pk.h_query.size(), h_assign.size()));
  size_t ni = m.num_instance, nw = m.num_witness;
  std::vector<Big<4>> aux(nw), assignment(ni - 1 + nw);
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < nw; i++) aux[i] = z[ni + i].to_big();
  G1J l_aux_acc = msm_pippenger<Fq>(pk.l_query.data(), aux.data(), std::min(pk.l_query.size(), aux.size()));
  G1J delta_g1 = G1J::from_affine(pk.delta_g1);
  G1J r_s_delta_g1 = delta_g1.mul(r.to_big()).mul(s.to_big());
  for (size_t i = 1; i < ni; i++) assignment[i - 1] = z[i].to_big();
  for (size_t i = 0; i < nw; i++) assignment[ni - 1 + i] = aux[i];
  G1J r_g1 = delta_g1.mul(r.to_big());
  G1J g_a = calculate_coeff<Fq>(r_g1, pk.a_query, pk.alpha_g1, assignment);
  G1J s_g_a = g_a.mul(s.to_big());
  G1J g1_b = G1J::infinity();
  if (!r.is_zero()) {
    G1J s_g1 = delta_g1.mul(s.to_big());
    g1_b = calculate_coeff<Fq>(s_g1, pk.b_g1_query, pk.beta_g1, assignment);
  }
  G2J s_g2 = G2J::from_affine(pk.delta_g2).mul(s.to_big());
  G2J g2_b = calculate_coeff<Fq2>(s_g2, pk.b_g2_query, pk.beta_g2, assignment);
  G1J r_g1_b = g1_b.mul(r.to_big());
  G1J g_c = s_g_a.add(r_g1_b).add(r_s_delta_g1.neg()).add(l_aux_acc).add(h_acc);
  return {g_a.to_affine(), g2_b.to_affine(), g_c.to_affine()};
}

// Check a proof "in the exponent" with the trapdoor: A = g1^a, B = g2^b, C = g1^c
// where a = alpha + sum z_i u_i(tau) + r delta, b = beta + sum z_i v_i(tau) + s delta,
// c = (sum_wit z_j (beta u_j + alpha v_j + w_j) + h(tau) Z(tau))/delta + s a + r b - r s delta.
// Equivalent to the pairing equation of verify_proof (pok_sig.rs:45-47) for honest keys.
static inline bool verify_with_trapdoor(const Matrices& m, const Domain& d, const Trapdoor& td, const Fr* z,
                                        const Fr& r, const Fr& s, const Proof& pf) {
  size_t ni = m.num_instance, nw = m.num_witness, nc = m.num_constraints, nv = ni + nw;
  std::vector<Fr> u = d.evaluate_all_lagrange_coefficients(td.tau);
  Fr A = Fr::zero(), B = Fr::zero(), C = Fr::zero();
  for (size_t i = 0; i < nc; i++) {
    A += u[i] * csr_row_dot(m.a, i, z);
    B += u[i] * csr_row_dot(m.b, i, z);
    C += u[i] * csr_row_dot(m.c, i, z);
  }
  for (size_t i = 0; i < ni; i++) A += u[nc + i] * z[i];
  Fr zt = d.evaluate_vanishing_polynomial(td.tau);
  // h(tau) = (A*B - C)/Z(tau): exists iff the R1CS is satisfied
  Fr htau_zt = A * B - C;
  // public part: sum_{i<ni} z_i (beta u_i + alpha v_i + w_i): needs per-variable evaluations
  std::vector<Fr> ua(nv, Fr::zero()), ub(nv, Fr::zero()), uc(nv, Fr::zero());
  for (size_t i = 0; i < ni; i++) ua[i] = u[nc + i];
  for (size_t i = 0; i < nc; i++) {
    for (uint32_t k = m.a.row_ptr[i]; k < m.a.row_ptr[i + 1]; k++) ua[m.a.col[k]] += u[i] * m.a.val[k];
    for (uint32_t k = m.b.row_ptr[i]; k < m.b.row_ptr[i + 1]; k++) ub[m.b.col[k]] += u[i] * m.b.val[k];
    for (uint32_t k = m.c.row_ptr[i]; k < m.c.row_ptr[i + 1]; k++) uc[m.c.col[k]] += u[i] * m.c.val[k];
  }
  Fr lw = Fr::zero();
  for (size_t j = ni; j < nv; j++) lw += z[j] * (td.beta * ua[j] + td.alpha * ub[j] + uc[j]);
  Fr a = td.alpha + A + r * td.delta;
  Fr b = td.beta + B + s * td.delta;
  Fr c = (lw + htau_zt) * td.delta.inverse() + s * a + r * b - r * s * td.delta;
  (void)zt;
  G1J g1 = G1J::from_affine(g1_generator()).mul(td.g1_scalar.to_big());
  G2J g2 = G2J::from_affine(g2_generator()).mul(td.g2_scalar.to_big());
  return g1.mul(a.to_big()).to_affine() == pf.a && g2.mul(b.to_big()).to_affine() == pf.b &&
         g1.mul(c.to_big()).to_affine() == pf.c;
}

// ark-serialize 0.3 compressed form: x little-endian, flags in
// the top bits of the last byte: bit7 = y > -y, bit6 = infinity.
static inline bool fq_gt(const Fq& a, const Fq& b) {
  Big<6> x = a.to_big(), y = b.to_big();
  return cmp_n<6>(x.l, y.l) > 0;
}
static inline void ser_fq(const Fq& a, uint8_t* out) {
  Big<6> x = a.to_big();
  memcpy(out, x.l, 48);
}
static inline void ser_g1(const G1A& p, uint8_t* out) {
  if (p.inf) {
    memset(out, 0, 48);
    out[47] |= 1 << 6;
    return;
  }
  ser_fq(p.x, out);
  if (fq_gt(p.y, -p.y)) out[47] |= 1 << 7;
}
static inline void ser_g2(const G2A& p, uint8_t* out) {
  if (p.inf) {
    memset(out, 0, 96);
    out[95] |= 1 << 6;
    return;
  }
  ser_fq(p.x.c0, out);
  ser_fq(p.x.c1, out + 48);
  Fq2 ny = -p.y;
  bool gt = p.y.c1 != ny.c1 ? fq_gt(p.y.c1, ny.c1) : fq_gt(p.y.c0, ny.c0);
  if (gt) out[95] |= 1 << 7;
}
static inline void ser_proof(const Proof& pf, uint8_t* out) {
  ser_g1(pf.a, out);
  ser_g2(pf.b, out + 48);
  ser_g1(pf.c, out + 144);
}

}  // namespace orc
