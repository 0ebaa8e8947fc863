// ORACLE — TEST INFRASTRUCTURE ONLY.  Restatement of the reference's gadgets and
// circuits on top of r1cs.hpp; every function cites the reference lines it follows.
// N / LOG_N are run-time here (the reference selects them with a cargo feature,
// falcon-r1cs/Cargo.toml:28-32).
#pragma once
#include "falcon.hpp"
#include "r1cs.hpp"

namespace orc {

enum CircuitKind { KIND_NTT = 0, KIND_SCHOOLBOOK = 1, KIND_DUAL_NTT = 2 };

// BigUint a / q, a % q on the canonical integer (arithmetics.rs:127-134)
static inline void divmod_q(const Fr& a, Fr* t, Fr* rem) {
  Big<4> x = a.to_big(), qt;
  u128 r = 0;
  for (int i = 3; i >= 0; i--) {
    u128 cur = (r << 64) | x.l[i];
    qt.l[i] = (uint64_t)(cur / FALCON_Q);
    r = cur % FALCON_Q;
  }
  *t = Fr::from_big(qt);
  *rem = Fr::from_u64((uint64_t)r);
}

struct Gadgets {
  Cs c;
  BoolOps b;
  int logn, n;
  bool panic_on_range;  // true = non-test build behaviour (record status where the reference panics)
  Gadgets(ConstraintSystem* cs, int logn_, bool panic_ = true) : c{cs}, b{{cs}}, logn(logn_), n(1 << logn_), panic_on_range(panic_) {}
  ConstraintSystem* cs() const { return c.cs; }

  Fr val_or_one(const FpVar& a) const { return cs()->is_in_setup_mode() ? Fr::one() : a.value(); }

  // gadgets/misc.rs:9-24
  void enforce_decompose(const FpVar& a, const std::vector<Boolean>& bits) const {
    FpVar res = b.to_fp(bits.back());
    for (size_t i = bits.size() - 1; i-- > 0;) res = c.add(c.dbl(res), b.to_fp(bits[i]));
    c.enforce_equal(res, a);
  }
  // a_val.into_repr().to_bits_le().take(k) -> Boolean::new_witness each
  std::vector<Boolean> alloc_bits(const Fr& a_val, int k) const {
    Big<4> r = a_val.to_big();
    std::vector<Boolean> v;
    for (int i = 0; i < k; i++) v.push_back(b.new_witness(r.bit(i)));
    return v;
  }
  static bool ge_u64(const Fr& a, uint64_t bound) {
    Big<4> r = a.to_big();
    return r.l[1] || r.l[2] || r.l[3] || r.l[0] >= bound;
  }
  // gadgets/range_proofs.rs:13-37
  void enforce_less_than_1024(const FpVar& a) const {
    Fr a_val = val_or_one(a);
    enforce_decompose(a, alloc_bits(a_val, 10));
  }
  // gadgets/range_proofs.rs:42-94
  void enforce_less_than_q(const FpVar& a) const {
    Fr a_val = val_or_one(a);
    if (panic_on_range && ge_u64(a_val, FALCON_Q)) cs()->fail(ORC_E_COEFF_RANGE);
    std::vector<Boolean> bits = alloc_bits(a_val, 14);
    enforce_decompose(a, bits);
    Boolean inner = b.kary_or(bits, 0, 12).not_();             // kary_or(a[0..12]).is_eq(FALSE)
    Boolean mid = b.or_(bits[12].not_(), inner);               // a[12]==0 .or(..)
    b.enforce_true(b.or_(bits[13].not_(), mid));               // a[13]==0 .or(..) == TRUE
  }
  // gadgets/range_proofs.rs:289-333
  Boolean is_less_than_6144(const FpVar& a) const {
    Fr a_val = val_or_one(a);
    std::vector<Boolean> bits = alloc_bits(a_val, 14);
    enforce_decompose(a, bits);
    Boolean inner = b.or_(bits[12].not_(), bits[11].not_());
    return b.and_(bits[13].not_(), inner);  // .is_eq(TRUE) is the identity
  }
  // gadgets/range_proofs.rs:100-186 (feature falcon-512)
  void enforce_less_than_norm_bound_512(const FpVar& a) const {
    Fr a_val = val_or_one(a);
    if (panic_on_range && ge_u64(a_val, sig_l2_bound(9))) cs()->fail(ORC_E_NORM_BOUND);
    std::vector<Boolean> x = alloc_bits(a_val, 26);
    enforce_decompose(a, x);
    // receivers are evaluated before arguments: kary results outermost-first,
    // binary results innermost-first.
    Boolean k19 = b.kary_or(x, 19, 25).not_();
    Boolean k16 = b.kary_and(x, 16, 19).not_();
    Boolean k6 = b.kary_or(x, 6, 10).not_();
    Boolean k3 = b.kary_or(x, 3, 5).not_();
    Boolean k1 = b.kary_and(x, 1, 3).not_();
    Boolean t = b.and_(k3, k1);
    t = b.or_(x[5].not_(), t);
    t = b.and_(k6, t);
    t = b.or_(x[10].not_(), t);
    t = b.and_(x[11].not_(), t);
    t = b.or_(x[12].not_(), t);
    t = b.and_(x[13].not_(), t);
    t = b.or_(x[14].not_(), t);
    t = b.and_(x[15].not_(), t);
    t = b.or_(k16, t);
    t = b.and_(k19, t);
    t = b.or_(x[25].not_(), t);
    b.enforce_true(t);
  }
  // gadgets/range_proofs.rs:192-272 (feature falcon-1024)
  void enforce_less_than_norm_bound_1024(const FpVar& a) const {
    Fr a_val = val_or_one(a);
    if (panic_on_range && ge_u64(a_val, sig_l2_bound(10))) cs()->fail(ORC_E_NORM_BOUND);
    std::vector<Boolean> x = alloc_bits(a_val, 27);
    enforce_decompose(a, x);
    Boolean k22 = b.kary_or(x, 22, 26).not_();
    Boolean k20 = b.kary_and(x, 20, 22).not_();
    Boolean k14 = b.kary_or(x, 14, 20).not_();
    Boolean k9 = b.kary_or(x, 9, 11).not_();
    Boolean k7 = b.kary_and(x, 7, 9).not_();
    Boolean k5 = b.kary_or(x, 5, 7).not_();
    Boolean k3 = b.kary_and(x, 3, 5).not_();
    Boolean k1 = b.kary_or(x, 1, 3).not_();
    Boolean t = b.or_(k3, k1);
    t = b.and_(k5, t);
    t = b.or_(k7, t);
    t = b.and_(k9, t);
    t = b.or_(x[11].not_(), t);
    t = b.and_(x[12].not_(), t);
    t = b.or_(x[13].not_(), t);
    t = b.and_(k14, t);
    t = b.or_(k20, t);
    t = b.and_(k22, t);
    t = b.or_(x[26].not_(), t);
    b.enforce_true(t);
  }
  // gadgets/range_proofs.rs:274-284
  void enforce_less_than_norm_bound(const FpVar& a) const {
    if (logn == 9)
      enforce_less_than_norm_bound_512(a);
    else
      enforce_less_than_norm_bound_1024(a);
  }
  // gadgets/arithmetics.rs:105-149
  FpVar mod_q(const FpVar& a, const FpVar& modulus_var) const {
    Fr a_val = val_or_one(a), t_val, b_val;
    divmod_q(a_val, &t_val, &b_val);
    FpVar t_var = c.new_witness(t_val);
    FpVar b_var = c.new_witness(b_val);
    FpVar left = c.sub(a, c.mul(t_var, modulus_var));
    c.enforce_equal(left, b_var);
    enforce_less_than_q(b_var);
    return b_var;
  }
  // gadgets/arithmetics.rs:214-262
  FpVar add_mod(const FpVar& a, const FpVar& bb, const FpVar& modulus_var) const {
    Fr ab = val_or_one(a) + val_or_one(bb), t_val, c_val;
    divmod_q(ab, &t_val, &c_val);  // t = (ab - c)/q == ab/q
    FpVar t_var = c.new_witness(t_val);
    FpVar c_var = c.new_witness(c_val);
    FpVar left = c.sub(c.add(a, bb), c.mul(t_var, modulus_var));
    c.enforce_equal(left, c_var);
    enforce_less_than_q(c_var);
    return c_var;
  }
  // gadgets/arithmetics.rs:157-205 (dead code in the reference; kept for its KATs :412-435)
  FpVar mul_mod(const FpVar& a, const FpVar& bb, const FpVar& modulus_var) const {
    Fr ab = val_or_one(a) * val_or_one(bb), t_val, c_val;
    divmod_q(ab, &t_val, &c_val);
    FpVar t_var = c.new_witness(t_val);
    FpVar c_var = c.new_witness(c_val);
    FpVar left = c.sub(c.mul(a, bb), c.mul(t_var, modulus_var));
    c.enforce_equal(left, c_var);
    enforce_less_than_q(c_var);
    return c_var;
  }
  // gadgets/arithmetics.rs:34-100
  FpVar inner_product_mod(const std::vector<FpVar>& a, const FpVar* bb, const FpVar& modulus_var) const {
    Fr ab = Fr::zero();
    if (cs()->is_in_setup_mode()) {
      ab = Fr::from_u64((uint64_t)n);  // vec![F::one(); N] . vec![F::one(); N]
    } else {
      for (size_t i = 0; i < a.size(); i++) ab += a[i].value() * bb[i].value();
    }
    Fr t_val, c_val;
    divmod_q(ab, &t_val, &c_val);
    FpVar t_var = c.new_witness(t_val);
    FpVar c_var = c.new_witness(c_val);
    FpVar ab_var = c.mul(a[0], bb[0]);
    for (size_t i = 1; i < a.size(); i++) ab_var = c.add(ab_var, c.mul(a[i], bb[i]));
    FpVar left = c.sub(ab_var, c.mul(t_var, modulus_var));
    c.enforce_equal(left, c_var);
    enforce_less_than_q(c_var);
    return c_var;
  }
  // gadgets/misc.rs:30-51
  FpVar l2_norm_var(const std::vector<FpVar>& input, const FpVar& modulus_var) const {
    FpVar res;
    for (size_t i = 0; i < input.size(); i++) {
      Boolean lt = is_less_than_6144(input[i]);
      FpVar tmp = b.select(lt, input[i], c.sub(modulus_var, input[i]));
      FpVar sq = c.mul(tmp, tmp);
      res = i == 0 ? sq : c.add(res, sq);
    }
    return res;
  }
  // gadgets/misc.rs:55-65
  FpVar l2_norm_var_without_range_check(const std::vector<FpVar>& input) const {
    FpVar res = c.mul(input[0], input[0]);
    for (size_t i = 1; i < input.size(); i++) res = c.add(res, c.mul(input[i], input[i]));
    return res;
  }
  // gadgets/misc.rs:67-77
  std::vector<FpVar> ntt_param_var() const {
    std::vector<FpVar> r;
    for (uint32_t e : ntt_table(n)) r.push_back(FpVar::constant(Fr::from_u64(e)));
    return r;
  }
  // falcon_ntt.rs:31-39: [q, 2 q^2, 4 q^3, ...], x = 1 .. LOG_N+1
  std::vector<FpVar> const_q_power_vars() const {
    std::vector<FpVar> r;
    for (int x = 1; x < logn + 2; x++)
      r.push_back(FpVar::constant(Fr::from_u64(1u << (x - 1)) * Fr::from_u64(FALCON_Q).pow_u64((uint64_t)x)));
    return r;
  }
  // gadgets/poly.rs:104-159
  std::vector<FpVar> ntt_circuit(const std::vector<FpVar>& input, const std::vector<FpVar>& const_vars,
                                 const std::vector<FpVar>& param) const {
    if ((int)input.size() != n) throw std::runtime_error("input length is not N");
    std::vector<FpVar> output(input);
    int t = n;
    for (int l = 0; l < logn; l++) {
      int m = 1 << l, ht = t / 2, j1 = 0;
      for (int i = 0; i < m; i++) {
        const FpVar& s = param[m + i];
        for (int j = j1; j < j1 + ht; j++) {
          FpVar u = output[j];
          FpVar v = c.mul(output[j + ht], s);
          FpVar neg_v = c.sub(const_vars[l + 1], v);
          output[j] = c.add(u, v);
          output[j + ht] = c.add(u, neg_v);
        }
        j1 += t;
      }
      t = ht;
    }
    for (auto& e : output) e = mod_q(e, const_vars[0]);
    return output;
  }
  std::vector<FpVar> alloc_vars(const std::vector<uint32_t>& coeff, bool input) const {  // poly.rs:47-63,195-211
    std::vector<FpVar> v;
    for (uint32_t x : coeff) v.push_back(input ? c.new_input(Fr::from_u64(x)) : c.new_witness(Fr::from_u64(x)));
    return v;
  }

  // circuits/falcon_ntt.rs:26-123.  sig, pk, hm are coefficient vectors in [0,q):
  // sig = Polynomial::from(&Signature), pk = Polynomial::from(&PublicKey),
  // hm = Polynomial::from_hash_of_message(msg, nonce) (hash-to-point stays outside).
  void falcon_ntt_circuit(const std::vector<uint32_t>& sig, const std::vector<uint32_t>& pk,
                          const std::vector<uint32_t>& hm) const {
    std::vector<FpVar> const_q = const_q_power_vars();
    std::vector<FpVar> param = ntt_param_var();
    std::vector<uint32_t> hm_ntt = ntt_clear(hm, logn);
    std::vector<uint32_t> v = poly_sub(hm, poly_mul(sig, pk));
    std::vector<uint32_t> pk_ntt = ntt_clear(pk, logn);
    std::vector<FpVar> sig_vars = alloc_vars(sig, false);
    std::vector<FpVar> pk_ntt_vars = alloc_vars(pk_ntt, true);
    std::vector<FpVar> hm_ntt_vars = alloc_vars(hm_ntt, true);
    std::vector<FpVar> v_vars = alloc_vars(v, false);
    for (auto& e : v_vars) enforce_less_than_q(e);
    std::vector<FpVar> sig_ntt = ntt_circuit(sig_vars, const_q, param);
    std::vector<FpVar> v_ntt = ntt_circuit(v_vars, const_q, param);
    for (int i = 0; i < n; i++) {
      FpVar prod = c.mul(sig_ntt[i], pk_ntt_vars[i]);
      c.enforce_equal(hm_ntt_vars[i], add_mod(v_ntt[i], prod, const_q[0]));
    }
    std::vector<FpVar> cat(v_vars);
    cat.insert(cat.end(), sig_vars.begin(), sig_vars.end());
    enforce_less_than_norm_bound(l2_norm_var(cat, const_q[0]));
  }
  // DualPolynomial::from(&Polynomial) ([EXT] falcon-rust, floating git dependency; restated: a coefficient below
  // (q-1)/2 = 6144 goes to `pos`, any other e to `neg` as q - e -- the same split is_less_than_6144 makes in
  // l2_norm_var, misc.rs:35-39, so both NTT circuits bound the same norm).  Parity unpinned.
  static void dual_split(const std::vector<uint32_t>& p, std::vector<uint32_t>* pos, std::vector<uint32_t>* neg) {
    pos->assign(p.size(), 0);
    neg->assign(p.size(), 0);
    for (size_t i = 0; i < p.size(); i++) {
      if (p[i] < 6144)
        (*pos)[i] = p[i];
      else
        (*neg)[i] = FALCON_Q - p[i];
    }
  }
  struct DualVars {
    std::vector<FpVar> pos, neg;
  };
  // DualPolyVar::alloc_vars (gadgets/dual_poly.rs:14-33): pos, neg, then sum_i pos_i * neg_i == 0 through
  // acc.is_zero()?.enforce_equal(TRUE).  FpVar::is_zero = self.is_eq(&zero()); the (Var, Constant) arm of
  // FpVar::is_eq turns the constant into an AllocatedFp and calls c.is_eq(v), i.e. x - y = 0 - acc.
  DualVars alloc_dual(const std::vector<uint32_t>& pos, const std::vector<uint32_t>& neg) const {
    DualVars d{alloc_vars(pos, false), alloc_vars(neg, false)};
    FpVar acc = c.mul(d.pos[0], d.neg[0]);
    for (int i = 1; i < n; i++) acc = c.add(acc, c.mul(d.pos[i], d.neg[i]));
    b.enforce_true(b.is_eq(FpVar::constant(Fr::zero()), acc));
    return d;
  }
  // circuits/falcon_dual_ntt.rs:26-132
  void falcon_dual_ntt_circuit(const std::vector<uint32_t>& sig, const std::vector<uint32_t>& pk,
                               const std::vector<uint32_t>& hm) const {
    std::vector<FpVar> const_q = const_q_power_vars();
    std::vector<FpVar> param = ntt_param_var();
    std::vector<uint32_t> hm_ntt = ntt_clear(hm, logn);
    // v = hm - sig.pos * pk + sig.neg * pk = hm - sig * pk   (:47-51)
    std::vector<uint32_t> v = poly_sub(hm, poly_mul(sig, pk));
    std::vector<uint32_t> pk_ntt = ntt_clear(pk, logn);
    std::vector<uint32_t> sp, sn, vp, vn;
    dual_split(sig, &sp, &sn);
    dual_split(v, &vp, &vn);
    DualVars sig_vars = alloc_dual(sp, sn);                         // :60-61
    std::vector<FpVar> pk_ntt_vars = alloc_vars(pk_ntt, true);      // :65
    std::vector<FpVar> hm_ntt_vars = alloc_vars(hm_ntt, true);      // :69
    DualVars v_vars = alloc_dual(vp, vn);                           // :73
    // DualNTTPolyVar::ntt_circuit (dual_poly.rs:42-52): pos then neg
    std::vector<FpVar> sig_ntt_pos = ntt_circuit(sig_vars.pos, const_q, param);
    std::vector<FpVar> sig_ntt_neg = ntt_circuit(sig_vars.neg, const_q, param);
    std::vector<FpVar> v_ntt_pos = ntt_circuit(v_vars.pos, const_q, param);
    std::vector<FpVar> v_ntt_neg = ntt_circuit(v_vars.neg, const_q, param);
    for (int i = 0; i < n; i++) {  // :96-116
      FpVar lsum = c.add(hm_ntt_vars[i], v_ntt_neg[i]);
      FpVar left = mod_q(c.add(lsum, c.mul(sig_ntt_neg[i], pk_ntt_vars[i])), const_q[0]);
      FpVar right = mod_q(c.add(v_ntt_pos[i], c.mul(sig_ntt_pos[i], pk_ntt_vars[i])), const_q[0]);
      c.enforce_equal(left, right);
    }
    std::vector<FpVar> cat(v_vars.pos);  // :121-129
    cat.insert(cat.end(), v_vars.neg.begin(), v_vars.neg.end());
    cat.insert(cat.end(), sig_vars.pos.begin(), sig_vars.pos.end());
    cat.insert(cat.end(), sig_vars.neg.begin(), sig_vars.neg.end());
    enforce_less_than_norm_bound(l2_norm_var_without_range_check(cat));
  }
  // circuits/falcon_schoolbook.rs:26-132
  void falcon_schoolbook_circuit(const std::vector<uint32_t>& sig, const std::vector<uint32_t>& pk,
                                 const std::vector<uint32_t>& hm) const {
    FpVar const_q = FpVar::constant(Fr::from_u64(FALCON_Q));
    std::vector<uint32_t> v = poly_sub(hm, poly_mul(sig, pk));
    std::vector<FpVar> sig_vars = alloc_vars(sig, false);
    std::vector<FpVar> pk_vars, neg_pk_vars;
    for (uint32_t e : pk) {
      FpVar tmp = c.new_input(Fr::from_u64(e));
      neg_pk_vars.push_back(c.sub(const_q, tmp));
      pk_vars.push_back(tmp);
    }
    std::vector<FpVar> hm_vars = alloc_vars(hm, true);
    std::vector<FpVar> v_vars;
    for (uint32_t e : v) {
      FpVar tmp = c.new_witness(Fr::from_u64(e));
      enforce_less_than_q(tmp);
      v_vars.push_back(tmp);
    }
    std::vector<FpVar> buf(neg_pk_vars);
    buf.insert(buf.end(), pk_vars.begin(), pk_vars.end());
    std::reverse(buf.begin(), buf.end());
    for (int i = 0; i < n; i++) {
      FpVar col = inner_product_mod(sig_vars, &buf[n - 1 - i], const_q);
      FpVar rhs = c.sub(c.add(hm_vars[i], const_q), col);
      Boolean e1 = b.is_eq(rhs, v_vars[i]);
      Boolean e2 = b.is_eq(rhs, c.add(v_vars[i], const_q));
      b.enforce_true(b.or_(e1, e2));
    }
    std::vector<FpVar> cat(v_vars);
    cat.insert(cat.end(), sig_vars.begin(), sig_vars.end());
    enforce_less_than_norm_bound(l2_norm_var(cat, const_q));
  }
};

}  // namespace orc
