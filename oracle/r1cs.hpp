// ORACLE — TEST INFRASTRUCTURE ONLY.  A small CPU model of the parts of
// ark-relations 0.3.0 (ConstraintSystem) and ark-r1cs-std 0.3.1 (FpVar, Boolean,
// AllocatedBool) that the reference's gadgets call.  Those crates are [EXT] to
// /root/reference (falcon-r1cs/Cargo.toml:14-19); their rules are restated from
// SURVEY.md App. B.1/B.2.  Parity pins available in-tree: variable/constraint
// counts (README.md:41-56), gadget known answers (arithmetics.rs:346-361,
// range_proofs.rs:365-389 ...), satisfiability.  Row *forms* are otherwise
// "parity unpinned" (no Rust toolchain here).
//
// Differences from arkworks that do not change any output: symbolic LCs are
// inlined eagerly (arkworks inlines them in finalize()); an LC is kept sorted by
// variable with duplicates merged (arkworks: compactify()).
#pragma once
#include <algorithm>
#include <memory>
#include <stdexcept>
#include <vector>

#include "ff.hpp"

namespace orc {

// Variable encoding: 0 = One, i>=1 = Instance(i), 0x80000000|j = Witness(j).
// Integer order == arkworks' Variable order (One < Instance < Witness).
typedef uint32_t Var;
static const Var VAR_ONE = 0;
static inline Var var_witness(uint32_t j) { return 0x80000000u | j; }
static inline bool var_is_witness(Var v) { return v >> 31; }

struct LC {
  std::vector<std::pair<Var, Fr>> t;  // sorted by Var, no duplicates
};
typedef std::shared_ptr<const LC> LCP;

static inline LCP lc_var(Var v, const Fr& c) {
  auto p = std::make_shared<LC>();
  p->t.push_back({v, c});
  return p;
}
static inline LCP lc_empty() { return std::make_shared<LC>(); }
static inline LCP lc_scale(const LCP& a, const Fr& c) {
  auto p = std::make_shared<LC>();
  p->t.reserve(a->t.size());
  for (auto& e : a->t) p->t.push_back({e.first, e.second * c});
  return p;
}
// a + s*b
static inline LCP lc_axpy(const LCP& a, const LCP& b, bool negate_b) {
  auto p = std::make_shared<LC>();
  p->t.reserve(a->t.size() + b->t.size());
  size_t i = 0, j = 0;
  while (i < a->t.size() || j < b->t.size()) {
    if (j == b->t.size() || (i < a->t.size() && a->t[i].first < b->t[j].first)) {
      p->t.push_back(a->t[i++]);
    } else if (i == a->t.size() || b->t[j].first < a->t[i].first) {
      p->t.push_back({b->t[j].first, negate_b ? -b->t[j].second : b->t[j].second});
      j++;
    } else {
      Fr c = negate_b ? a->t[i].second - b->t[j].second : a->t[i].second + b->t[j].second;
      p->t.push_back({a->t[i].first, c});
      i++;
      j++;
    }
  }
  return p;
}
static inline LCP lc_add_const(const LCP& a, const Fr& c) { return lc_axpy(a, lc_var(VAR_ONE, c), false); }

enum Mode { MODE_SETUP = 0, MODE_PROVE = 1 };

// Status codes recorded where the reference would panic (non-test builds).
enum {
  ORC_OK = 0,
  ORC_E_COEFF_RANGE = -1,  // range_proofs.rs:58-60
  ORC_E_NORM_BOUND = -2,   // range_proofs.rs:114-117, 205-208
};

struct CSR {
  std::vector<uint32_t> row_ptr, col;
  std::vector<Fr> val;
};

struct ConstraintSystem {
  Mode mode = MODE_PROVE;
  bool construct_matrices = true;  // ark-relations SynthesisMode::Prove{construct_matrices}
  std::vector<Fr> instance;        // instance_assignment (index 0 = One)
  std::vector<Fr> witness;         // witness_assignment
  uint32_t num_instance = 1, num_witness = 0, num_constraints = 0;
  std::vector<LCP> a, b, c;
  int status = ORC_OK;

  ConstraintSystem() { instance.push_back(Fr::one()); }
  bool is_in_setup_mode() const { return mode == MODE_SETUP; }
  bool build_lc() const { return mode == MODE_SETUP || construct_matrices; }
  void fail(int code) {
    if (status == ORC_OK) status = code;
  }

  Var new_input_variable(const Fr* val) {  // ConstraintSystem::new_input_variable
    Var v = num_instance++;
    if (mode != MODE_SETUP) instance.push_back(*val);
    return v;
  }
  Var new_witness_variable(const Fr* val) {  // ConstraintSystem::new_witness_variable
    Var v = var_witness(num_witness++);
    if (mode != MODE_SETUP) witness.push_back(*val);
    return v;
  }
  void enforce_constraint(const LCP& la, const LCP& lb, const LCP& lc) {
    num_constraints++;
    if (build_lc()) {
      a.push_back(la);
      b.push_back(lb);
      c.push_back(lc);
    }
  }
  uint32_t col_of(Var v) const { return var_is_witness(v) ? num_instance + (v & 0x7fffffffu) : v; }
  // to_matrices()/make_row: zero coefficients dropped, column = get_index_unchecked
  CSR to_csr(const std::vector<LCP>& m) const {
    CSR r;
    r.row_ptr.push_back(0);
    for (auto& lc : m) {
      for (auto& e : lc->t)
        if (!e.second.is_zero()) {
          r.col.push_back(col_of(e.first));
          r.val.push_back(e.second);
        }
      r.row_ptr.push_back((uint32_t)r.col.size());
    }
    return r;
  }
  Fr eval_lc(const LCP& lc) const {
    Fr s = Fr::zero();
    for (auto& e : lc->t) {
      const Fr& z = var_is_witness(e.first) ? witness[e.first & 0x7fffffffu] : instance[e.first];
      s += e.second * z;
    }
    return s;
  }
  // which_is_unsatisfied(): index of the first violated row, -1 if none
  int64_t first_unsatisfied() const {
    for (size_t i = 0; i < a.size(); i++)
      if (eval_lc(a[i]) * eval_lc(b[i]) != eval_lc(c[i])) return (int64_t)i;
    return -1;
  }
};

// ---- ark-r1cs-std: FpVar -----------------------------------------------------
struct FpVar {
  bool is_const = true;
  Fr cval = Fr::zero();  // constant value, or assigned value (prove mode)
  bool has_value = true;
  LCP lc;  // Var(..): the (inlined) linear combination

  static FpVar constant(const Fr& c) {
    FpVar v;
    v.cval = c;
    return v;
  }
  Fr value() const {
    if (!has_value) throw std::runtime_error("AssignmentMissing");
    return cval;
  }
};

struct Cs {  // thin handle playing ConstraintSystemRef
  ConstraintSystem* cs;

  FpVar make_var(const LCP& lc, const Fr* val) const {
    FpVar v;
    v.is_const = false;
    v.has_value = val != nullptr;
    if (val) v.cval = *val;
    v.lc = lc;
    return v;
  }
  // FpVar::new_witness / new_input (AllocVar): value closure evaluated only outside setup
  FpVar new_witness(const Fr& val) const {
    Var x = cs->new_witness_variable(&val);
    return make_var(cs->build_lc() ? lc_var(x, Fr::one()) : lc_empty(), cs->is_in_setup_mode() ? nullptr : &val);
  }
  FpVar new_input(const Fr& val) const {
    Var x = cs->new_input_variable(&val);
    return make_var(cs->build_lc() ? lc_var(x, Fr::one()) : lc_empty(), cs->is_in_setup_mode() ? nullptr : &val);
  }
  LCP as_lc(const FpVar& a) const { return a.is_const ? lc_var(VAR_ONE, a.cval) : a.lc; }

  FpVar add(const FpVar& a, const FpVar& b) const {
    if (a.is_const && b.is_const) return FpVar::constant(a.cval + b.cval);
    FpVar r;
    r.is_const = false;
    r.has_value = a.has_value && b.has_value;
    if (r.has_value) r.cval = a.cval + b.cval;
    if (cs->build_lc()) r.lc = lc_axpy(as_lc(a), as_lc(b), false);
    return r;
  }
  FpVar sub(const FpVar& a, const FpVar& b) const {
    if (a.is_const && b.is_const) return FpVar::constant(a.cval - b.cval);
    FpVar r;
    r.is_const = false;
    r.has_value = a.has_value && b.has_value;
    if (r.has_value) r.cval = a.cval - b.cval;
    if (cs->build_lc()) r.lc = lc_axpy(as_lc(a), as_lc(b), true);
    return r;
  }
  FpVar dbl(const FpVar& a) const { return add(a, a); }
  // FpVar * FpVar: Var*Constant is an LC; Var*Var allocates `product` and one row
  FpVar mul(const FpVar& a, const FpVar& b) const {
    if (a.is_const && b.is_const) return FpVar::constant(a.cval * b.cval);
    if (a.is_const || b.is_const) {
      const FpVar& v = a.is_const ? b : a;
      const Fr& k = a.is_const ? a.cval : b.cval;
      FpVar r;
      r.is_const = false;
      r.has_value = v.has_value;
      if (r.has_value) r.cval = v.cval * k;
      if (cs->build_lc()) r.lc = lc_scale(v.lc, k);
      return r;
    }
    Fr pv = Fr::zero();
    bool hv = a.has_value && b.has_value;
    if (hv) pv = a.cval * b.cval;
    FpVar p = new_witness(pv);
    cs->enforce_constraint(a.lc, b.lc, p.lc);
    return p;
  }
  // EqGadget for FpVar: <self - other | 1 | 0>
  void enforce_equal(const FpVar& a, const FpVar& b) const {
    if (a.is_const && b.is_const) return;
    // (Constant(c), Var(v)) | (Var(v), Constant(c)) => c.conditional_enforce_equal(v): <c - v | 1 | 0>
    bool swap = !a.is_const && b.is_const;
    LCP d = cs->build_lc() ? (swap ? lc_axpy(as_lc(b), as_lc(a), true) : lc_axpy(as_lc(a), as_lc(b), true)) : LCP();
    cs->enforce_constraint(d, cs->build_lc() ? lc_var(VAR_ONE, Fr::one()) : LCP(), cs->build_lc() ? lc_empty() : LCP());
  }
};

// ---- ark-r1cs-std: Boolean / AllocatedBool ---------------------------------
struct Bit {  // AllocatedBool
  Var var = 0;
  bool val = false;
  bool has_value = false;
};
struct Boolean {
  enum Kind { IS, NOT, CONST } kind = CONST;
  Bit bit;
  bool cst = false;
  static Boolean constant(bool b) {
    Boolean r;
    r.kind = CONST;
    r.cst = b;
    return r;
  }
  static Boolean is(const Bit& b) {
    Boolean r;
    r.kind = IS;
    r.bit = b;
    return r;
  }
  Boolean not_() const {
    Boolean r = *this;
    if (kind == CONST)
      r.cst = !cst;
    else
      r.kind = kind == IS ? NOT : IS;
    return r;
  }
  bool value() const {
    if (kind == CONST) return cst;
    if (!bit.has_value) throw std::runtime_error("AssignmentMissing");
    return kind == IS ? bit.val : !bit.val;
  }
};

struct BoolOps {
  Cs c;
  ConstraintSystem* cs() const { return c.cs; }
  LCP one() const { return lc_var(VAR_ONE, Fr::one()); }
  LCP v(const Bit& b) const { return lc_var(b.var, Fr::one()); }
  LCP not_v(const Bit& b) const { return lc_axpy(one(), v(b), true); }  // 1 - b

  Bit alloc_unchecked(bool val) const {  // new_witness_without_booleanity_check
    Bit b;
    Fr f = val ? Fr::one() : Fr::zero();
    b.var = cs()->new_witness_variable(&f);
    b.val = val;
    b.has_value = !cs()->is_in_setup_mode();
    return b;
  }
  // Boolean::new_witness -> AllocatedBool::new_variable: row <1-a | a | 0>
  Boolean new_witness(bool val) const {
    Bit b = alloc_unchecked(val);
    if (cs()->build_lc())
      cs()->enforce_constraint(not_v(b), v(b), lc_empty());
    else
      cs()->enforce_constraint(LCP(), LCP(), LCP());
    return Boolean::is(b);
  }
  void row(const LCP& a, const LCP& b, const LCP& cc) const { cs()->enforce_constraint(a, b, cc); }
  bool bl() const { return cs()->build_lc(); }
  // AllocatedBool::{and, or, and_not, nor}
  Bit a_and(const Bit& a, const Bit& b) const {
    Bit r = alloc_unchecked(a.val & b.val);
    bl() ? row(v(a), v(b), v(r)) : row(LCP(), LCP(), LCP());
    return r;
  }
  Bit a_or(const Bit& a, const Bit& b) const {
    Bit r = alloc_unchecked(a.val | b.val);
    bl() ? row(not_v(a), not_v(b), not_v(r)) : row(LCP(), LCP(), LCP());
    return r;
  }
  Bit a_and_not(const Bit& a, const Bit& b) const {
    Bit r = alloc_unchecked(a.val & !b.val);
    bl() ? row(v(a), not_v(b), v(r)) : row(LCP(), LCP(), LCP());
    return r;
  }
  Bit a_nor(const Bit& a, const Bit& b) const {
    Bit r = alloc_unchecked(!a.val & !b.val);
    bl() ? row(not_v(a), not_v(b), v(r)) : row(LCP(), LCP(), LCP());
    return r;
  }
  // Boolean::and
  Boolean and_(const Boolean& s, const Boolean& o) const {
    if (s.kind == Boolean::CONST) return s.cst ? o : Boolean::constant(false);
    if (o.kind == Boolean::CONST) return o.cst ? s : Boolean::constant(false);
    if (s.kind == Boolean::IS && o.kind == Boolean::NOT) return Boolean::is(a_and_not(s.bit, o.bit));
    if (s.kind == Boolean::NOT && o.kind == Boolean::IS) return Boolean::is(a_and_not(o.bit, s.bit));
    if (s.kind == Boolean::NOT && o.kind == Boolean::NOT) return Boolean::is(a_nor(s.bit, o.bit));
    return Boolean::is(a_and(s.bit, o.bit));
  }
  // Boolean::or: (Is,Is) -> AllocatedBool::or; otherwise NOT((NOT a) AND (NOT b))
  // with the match-arm bindings of ark-r1cs-std 0.3.1:
  //   (a@Is, b@Not) | (b@Not, a@Is) | (b@Not, a@Not)
  Boolean or_(const Boolean& s, const Boolean& o) const {
    if (s.kind == Boolean::CONST) return s.cst ? Boolean::constant(true) : o;
    if (o.kind == Boolean::CONST) return o.cst ? Boolean::constant(true) : s;
    if (s.kind == Boolean::IS && o.kind == Boolean::IS) return Boolean::is(a_or(s.bit, o.bit));
    const Boolean *a, *b;
    if (s.kind == Boolean::IS) {
      a = &s;
      b = &o;
    } else {
      b = &s;
      a = &o;
    }
    return and_(a->not_(), b->not_()).not_();
  }
  Boolean kary_or(const std::vector<Boolean>& bits, size_t lo, size_t hi) const {
    Boolean cur = bits[lo];
    for (size_t i = lo + 1; i < hi; i++) cur = or_(cur, bits[i]);
    return cur;
  }
  Boolean kary_and(const std::vector<Boolean>& bits, size_t lo, size_t hi) const {
    Boolean cur = bits[lo];
    for (size_t i = lo + 1; i < hi; i++) cur = and_(cur, bits[i]);
    return cur;
  }
  // Boolean::lc()
  LCP lc(const Boolean& x) const {
    if (x.kind == Boolean::CONST) return x.cst ? one() : lc_empty();
    return x.kind == Boolean::IS ? v(x.bit) : not_v(x.bit);
  }
  // x.enforce_equal(&Boolean::TRUE): Is(a) -> <1-a|1|0>, Not(a) -> <a|1|0>
  void enforce_true(const Boolean& x) const {
    if (x.kind == Boolean::CONST) {
      if (!x.cst) throw std::runtime_error("enforce_equal(FALSE, TRUE)");
      return;
    }
    if (!bl()) {
      row(LCP(), LCP(), LCP());
      return;
    }
    row(x.kind == Boolean::IS ? not_v(x.bit) : v(x.bit), one(), lc_empty());
  }
  // FpVar::from(Boolean)
  FpVar to_fp(const Boolean& x) const {
    if (x.kind == Boolean::CONST) return FpVar::constant(x.cst ? Fr::one() : Fr::zero());
    FpVar r;
    r.is_const = false;
    r.has_value = x.bit.has_value;
    if (r.has_value) r.cval = x.value() ? Fr::one() : Fr::zero();
    if (bl()) r.lc = lc(x);
    return r;
  }
  // FpVar::conditionally_select(cond, t, f): wit result; row <cond | t-f | result-f>
  FpVar select(const Boolean& cond, const FpVar& t, const FpVar& f) const {
    if (cond.kind == Boolean::CONST) return cond.cst ? t : f;
    Fr rv = Fr::zero();
    if (!cs()->is_in_setup_mode()) rv = cond.value() ? t.value() : f.value();
    FpVar r = c.new_witness(rv);
    if (bl())
      row(lc(cond), lc_axpy(c.as_lc(t), c.as_lc(f), true), lc_axpy(r.lc, c.as_lc(f), true));
    else
      row(LCP(), LCP(), LCP());
    return r;
  }
  // AllocatedFp::is_neq -> used via FpVar::is_eq (schoolbook circuit)
  Boolean is_eq(const FpVar& x, const FpVar& y) const {
    bool ne = false;
    Fr mult = Fr::one();
    if (!cs()->is_in_setup_mode()) {
      Fr d = x.value() - y.value();
      ne = !d.is_zero();
      mult = ne ? d.inverse() : Fr::one();
    }
    Boolean neb = new_witness(ne);  // is_not_equal bit (+ booleanity row)
    FpVar m = c.new_witness(mult);
    if (bl()) {
      LCP d = lc_axpy(c.as_lc(x), c.as_lc(y), true);
      row(d, m.lc, v(neb.bit));
      row(d, not_v(neb.bit), lc_empty());
    } else {
      row(LCP(), LCP(), LCP());
      row(LCP(), LCP(), LCP());
    }
    return neb.not_();
  }
};

}  // namespace orc
