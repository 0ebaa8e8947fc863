// Host-side circuit compiler: emits the R1CS matrices A, B, C (CSR, coefficients as
// canonical integers mod r) and the witness layout of the reference's
// FalconNTTVerificationCircuit (circuits/falcon_ntt.rs:26-123) in closed form, i.e.
// what `cs.to_matrices()` returns after `generate_constraints` + `finalize()`.
//
// Nothing here interprets gadgets symbolically: each gadget's rows are written down
// directly from its definition (row forms: SURVEY.md App. A/B), and the inlined
// linear combinations of ntt_circuit (gadgets/poly.rs:104-159) come from the product
// form of the unreduced butterfly network.  tests/ compares the result entry by
// entry with the oracle's generic arkworks-style synthesis.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <utility>
#include <vector>

namespace circuit {

static const uint32_t Q = 12289;  // falcon_rust::MODULUS

struct U256 {
  uint32_t v[8];
};
static const U256 R_MOD = {{0x00000001u, 0xffffffffu, 0xfffe5bfeu, 0x53bda402u, 0x09a1d805u, 0x3339d808u, 0x299d7d48u,
                            0x73eda753u}};

static inline U256 u256_small(uint64_t x) {
  U256 r;
  memset(&r, 0, sizeof r);
  r.v[0] = (uint32_t)x;
  r.v[1] = (uint32_t)(x >> 32);
  return r;
}
static inline U256 u256_pow2(int i) {
  U256 r;
  memset(&r, 0, sizeof r);
  r.v[i >> 5] = 1u << (i & 31);
  return r;
}
static inline U256 u256_sub(const U256& a, const U256& b) {
  U256 r;
  int64_t br = 0;
  for (int i = 0; i < 8; i++) {
    int64_t d = (int64_t)a.v[i] - b.v[i] - br;
    r.v[i] = (uint32_t)d;
    br = d < 0;
  }
  return r;
}
static inline U256 u256_add(const U256& a, const U256& b) {
  U256 r;
  uint64_t c = 0;
  for (int i = 0; i < 8; i++) {
    c += (uint64_t)a.v[i] + b.v[i];
    r.v[i] = (uint32_t)c;
    c >>= 32;
  }
  return r;
}
static inline U256 u256_mul_small(const U256& a, uint32_t s) {
  U256 r;
  uint64_t c = 0;
  for (int i = 0; i < 8; i++) {
    c += (uint64_t)a.v[i] * s;
    r.v[i] = (uint32_t)c;
    c >>= 32;
  }
  return r;
}
static inline bool u256_is_zero(const U256& a) {
  uint32_t o = 0;
  for (int i = 0; i < 8; i++) o |= a.v[i];
  return o == 0;
}
static inline U256 fr_neg(const U256& a) { return u256_is_zero(a) ? a : u256_sub(R_MOD, a); }  // a < r
static inline U256 fr_add(const U256& a, const U256& b) {
  U256 s = u256_add(a, b);  // < 2r < 2^256
  U256 t = u256_sub(s, R_MOD);
  // borrow <=> s < r
  bool lt = false;
  for (int i = 7; i >= 0; i--) {
    if (s.v[i] != R_MOD.v[i]) {
      lt = s.v[i] < R_MOD.v[i];
      break;
    }
  }
  return lt ? s : t;
}

static inline uint32_t powmod_q(uint32_t b, uint32_t e) {
  uint64_t r = 1, x = b % Q;
  while (e) {
    if (e & 1) r = r * x % Q;
    x = x * x % Q;
    e >>= 1;
  }
  return (uint32_t)r;
}
static inline uint32_t bitrev(uint32_t x, int bits) {
  uint32_t r = 0;
  for (int i = 0; i < bits; i++) r |= ((x >> i) & 1) << (bits - 1 - i);
  return r;
}
// falcon_rust::NTT_TABLE[i] = 7^bitrev10(i) mod q (script/ntt_param.sage:3-132)
static inline std::vector<uint32_t> ntt_table(int n) {
  std::vector<uint32_t> t(n);
  for (int i = 0; i < n; i++) t[i] = powmod_q(7, bitrev(i, 10));
  return t;
}

// ---- boolean-chain of enforce_less_than_norm_bound (range_proofs.rs:100-186, 192-272)
enum BoolOpKind { OP_AND = 0, OP_OR = 1, OP_AND_NOT = 2, OP_NOR = 3 };
struct BoolOp {
  uint8_t kind, a, b;  // operands: indices into the gadget-local witness list
};
struct NormProgram {
  int nbits;
  std::vector<BoolOp> ops;  // result k lives at local index nbits + k
};
static inline NormProgram norm_program(int logn) {
  NormProgram p;
  auto kary = [&](BoolOpKind k, int lo, int hi) {  // left fold over bits [lo, hi)
    int cur = lo;
    for (int i = lo + 1; i < hi; i++) {
      p.ops.push_back({(uint8_t)k, (uint8_t)cur, (uint8_t)i});
      cur = p.nbits + (int)p.ops.size() - 1;
    }
    return cur;
  };
  auto op = [&](BoolOpKind k, int a, int b) {
    p.ops.push_back({(uint8_t)k, (uint8_t)a, (uint8_t)b});
    return p.nbits + (int)p.ops.size() - 1;
  };
  if (logn == 9) {
    p.nbits = 26;
    int k19 = kary(OP_OR, 19, 25), k16 = kary(OP_AND, 16, 19), k6 = kary(OP_OR, 6, 10), k3 = kary(OP_OR, 3, 5),
        k1 = kary(OP_AND, 1, 3);
    int c = op(OP_NOR, k3, k1);
    c = op(OP_AND_NOT, 5, c);
    c = op(OP_NOR, k6, c);
    c = op(OP_AND_NOT, 10, c);
    c = op(OP_NOR, 11, c);
    c = op(OP_AND_NOT, 12, c);
    c = op(OP_NOR, 13, c);
    c = op(OP_AND_NOT, 14, c);
    c = op(OP_NOR, 15, c);
    c = op(OP_AND_NOT, k16, c);
    c = op(OP_NOR, k19, c);
    c = op(OP_AND_NOT, 25, c);
  } else {
    p.nbits = 27;
    int k22 = kary(OP_OR, 22, 26), k20 = kary(OP_AND, 20, 22), k14 = kary(OP_OR, 14, 20), k9 = kary(OP_OR, 9, 11),
        k7 = kary(OP_AND, 7, 9), k5 = kary(OP_OR, 5, 7), k3 = kary(OP_AND, 3, 5), k1 = kary(OP_OR, 1, 3);
    int c = op(OP_AND, k1, k3);
    c = op(OP_NOR, k5, c);
    c = op(OP_AND_NOT, k7, c);
    c = op(OP_NOR, k9, c);
    c = op(OP_AND_NOT, 11, c);
    c = op(OP_NOR, 12, c);
    c = op(OP_AND_NOT, 13, c);
    c = op(OP_NOR, k14, c);
    c = op(OP_AND_NOT, k20, c);
    c = op(OP_NOR, k22, c);
    c = op(OP_AND_NOT, 26, c);
  }
  return p;
}

// ---- witness / row layout of the NTT circuit -----------------
struct Layout {
  uint32_t logn, n, kind;
  uint32_t n_inst, n_wit, n_cons, n_z;
  // witness-index offsets (z column = n_inst + w)
  uint32_t w_sig, w_v, w_vrange, w_nttsig, w_nttv, w_pw, w_l2, w_norm;
  // row offsets
  uint32_t r_vrange, r_nttsig, r_nttv, r_pw, r_l2, r_norm;
  uint32_t norm_bits, norm_ops;
  uint32_t l2_bound;
  // schoolbook circuit (kind 1): v[i] and its 27 range witnesses are interleaved
  // (28 per coefficient, from w_v); column i of the product occupies sb_col witnesses from w_cols:
  // t, c, m_0..m_{N-1}, 27 range witnesses of c, ne1, mult1, ne2, mult2, w
  uint32_t w_cols, sb_col, r_cols, sb_col_rows;
  // dual-NTT circuit (kind 2, circuits/falcon_dual_ntt.rs:26-132): sig and v are (pos, neg) pairs; a pair takes
  // 3N + 2 witnesses (pos[N], neg[N], the N products pos_i * neg_i, then `ne` and `multiplier` of is_zero) and
  // N + 4 rows.  w_sig / w_v = first witness of the pair; w_ntt4 / r_ntt4 = the four ntt_circuit blocks (sig.pos,
  // sig.neg, v.pos, v.neg; 29N witnesses, 30N rows each); w_pw / r_pw = the pointwise section (60 witnesses,
  // 63 rows per index); w_l2 = the 4N squares (v.pos, v.neg, sig.pos, sig.neg).
  uint32_t w_ntt4, r_sig, r_v, r_ntt4;
};
static inline Layout make_layout_ntt(uint32_t logn) {
  Layout L;
  memset(&L, 0, sizeof L);
  uint32_t n = 1u << logn;
  NormProgram np = norm_program(logn);
  L.logn = logn;
  L.n = n;
  L.kind = 0;
  L.n_inst = 1 + 2 * n;
  L.w_sig = 0;
  L.w_v = n;
  L.w_vrange = 2 * n;
  L.w_nttsig = L.w_vrange + 27 * n;
  L.w_nttv = L.w_nttsig + 29 * n;
  L.w_pw = L.w_nttv + 29 * n;
  L.w_l2 = L.w_pw + 30 * n;
  L.w_norm = L.w_l2 + 18 * 2 * n;
  L.norm_bits = np.nbits;
  L.norm_ops = (uint32_t)np.ops.size();
  L.n_wit = L.w_norm + L.norm_bits + L.norm_ops;
  L.r_vrange = 0;
  L.r_nttsig = 29 * n;
  L.r_nttv = L.r_nttsig + 30 * n;
  L.r_pw = L.r_nttv + 30 * n;
  L.r_l2 = L.r_pw + 32 * n;
  L.r_norm = L.r_l2 + 19 * 2 * n;
  L.n_cons = L.r_norm + L.norm_bits + 1 + L.norm_ops + 1;
  L.n_z = L.n_inst + L.n_wit;
  L.l2_bound = logn == 9 ? 34034726u : 70265242u;  // range_proofs.rs:104,196
  return L;
}

// FalconSchoolBookVerificationCircuit (circuits/falcon_schoolbook.rs:26-132)
static inline Layout make_layout_sb(uint32_t logn) {
  Layout L;
  memset(&L, 0, sizeof L);
  const uint32_t n = 1u << logn;
  NormProgram np = norm_program(logn);
  L.logn = logn;
  L.n = n;
  L.kind = 1;
  L.n_inst = 1 + 2 * n;  // One, pk[N], hm[N]
  L.w_sig = 0;
  L.w_v = n;                // v[i] at w_v + 28 i, its range witnesses at +1
  L.w_vrange = n + 1;
  L.w_cols = n + 28 * n;
  L.sb_col = n + 34;
  L.w_l2 = L.w_cols + L.sb_col * n;
  L.w_norm = L.w_l2 + 18 * 2 * n;
  L.norm_bits = np.nbits;
  L.norm_ops = (uint32_t)np.ops.size();
  L.n_wit = L.w_norm + L.norm_bits + L.norm_ops;
  L.r_vrange = 0;
  L.r_cols = 29 * n;
  L.sb_col_rows = n + 38;
  L.r_l2 = L.r_cols + L.sb_col_rows * n;
  L.r_norm = L.r_l2 + 19 * 2 * n;
  L.n_cons = L.r_norm + L.norm_bits + 1 + L.norm_ops + 1;
  L.n_z = L.n_inst + L.n_wit;
  L.l2_bound = logn == 9 ? 34034726u : 70265242u;
  return L;
}

// FalconDualNTTVerificationCircuit (circuits/falcon_dual_ntt.rs:26-132, gadgets/dual_poly.rs:8-52)
static inline Layout make_layout_dual(uint32_t logn) {
  Layout L;
  memset(&L, 0, sizeof L);
  const uint32_t n = 1u << logn;
  NormProgram np = norm_program(logn);
  L.logn = logn;
  L.n = n;
  L.kind = 2;
  L.n_inst = 1 + 2 * n;
  L.w_sig = 0;
  L.w_v = 3 * n + 2;
  L.w_ntt4 = 6 * n + 4;
  L.w_pw = L.w_ntt4 + 4 * 29 * n;
  L.w_l2 = L.w_pw + 60 * n;
  L.w_norm = L.w_l2 + 4 * n;
  L.norm_bits = np.nbits;
  L.norm_ops = (uint32_t)np.ops.size();
  L.n_wit = L.w_norm + L.norm_bits + L.norm_ops;
  L.r_sig = 0;
  L.r_v = n + 4;
  L.r_ntt4 = 2 * n + 8;
  L.r_pw = L.r_ntt4 + 4 * 30 * n;
  L.r_l2 = L.r_pw + 63 * n;
  L.r_norm = L.r_l2 + 4 * n;
  L.n_cons = L.r_norm + L.norm_bits + 1 + L.norm_ops + 1;
  L.n_z = L.n_inst + L.n_wit;
  L.l2_bound = logn == 9 ? 34034726u : 70265242u;
  return L;
}

// ---- stand-alone gadget circuits (the reference's gadget entry points, gadgets/mod.rs:7-11) ----------------------
// z = [1 | operands | gadget witnesses | expected], all witnesses (the reference's gadget tests allocate the operands
// with FpVar::new_witness); `expected` (optional) is the extra witness + `out.enforce_equal(expected)` row of the
// test macros (arithmetics.rs:322-324, 461-463).
enum GadgetId {
  GADGET_MOD_Q = 0,           // arithmetics.rs:105-149
  GADGET_ADD_MOD = 1,         // arithmetics.rs:214-262
  GADGET_LESS_THAN_Q = 2,     // range_proofs.rs:42-94
  GADGET_LESS_THAN_6144 = 3,  // range_proofs.rs:289-333 (+ .enforce_equal(TRUE) when `expected` is requested)
  GADGET_NORM_BOUND = 4,      // range_proofs.rs:274-284
  GADGET_NTT_CIRCUIT = 5,     // poly.rs:104-159
  GADGET_COUNT = 6
};
struct GadgetShape {
  uint32_t n_operands, n_wit, n_rows;  // n_wit / n_rows without the `expected` witness / row
  int out;                             // gadget-local witness index of the output variable, -1 if none
};
static inline GadgetShape gadget_shape(int gadget, uint32_t logn) {
  const uint32_t n = 1u << logn;
  NormProgram np = norm_program(logn);
  switch (gadget) {
    case GADGET_MOD_Q: return {1, 29, 30, 1};
    case GADGET_ADD_MOD: return {2, 29, 30, 1};
    case GADGET_LESS_THAN_Q: return {1, 27, 29, -1};
    case GADGET_LESS_THAN_6144: return {1, 16, 17, 15};
    case GADGET_NORM_BOUND: return {1, (uint32_t)(np.nbits + np.ops.size()), (uint32_t)(np.nbits + 1 + np.ops.size() + 1), -1};
    case GADGET_NTT_CIRCUIT: return {n, 29 * n, 30 * n, -1};
  }
  return {0, 0, 0, -1};
}

struct HostCSR {
  std::vector<uint32_t> row_ptr, col;
  std::vector<U256> val;
};

// One ntt_circuit call (gadgets/poly.rs:104-159) inside a circuit: its N mod_q rows <L_k(x) - q t_k - b_k | 1 | 0> are
// the inlined outputs L_k of the unreduced butterfly network over the inputs z[in_col0 .. in_col0 + N) and the constant
// column.  Row k of the block is row0 + 30 k; t_k, b_k are the columns out_col0 + 29 k and + 1.  The R1CS evaluator
// may apply the network itself (N log N butterflies) instead of the N dense rows of ~N terms: same field elements.
struct NttBlock {
  uint32_t in_col0, out_col0, row0;
};

struct Matrices {
  Layout L;
  HostCSR a, b, c;
  std::vector<NttBlock> ntt_blocks;
};

class Builder {
 public:
  explicit Builder(uint32_t logn, uint32_t kind = 0)
      : L(kind == 1 ? make_layout_sb(logn) : kind == 2 ? make_layout_dual(logn) : make_layout_ntt(logn)) {
    one = u256_small(1);
    minus_one = fr_neg(one);
    minus_q = fr_neg(u256_small(Q));
    for (auto* m : {&M.a, &M.b, &M.c}) m->row_ptr.push_back(0);
  }
  // One gadget as a circuit of its own (see GadgetId); with_expected: the test macros' extra witness and row.
  static Matrices build_gadget(uint32_t logn, int gadget, bool with_expected) {
    Builder b(logn, 0);
    const GadgetShape gs = gadget_shape(gadget, logn);
    Layout& L = b.L;
    L.kind = 16 + (uint32_t)gadget;
    L.n_inst = 1;
    const bool expected_wit = with_expected && (gadget == GADGET_MOD_Q || gadget == GADGET_ADD_MOD);
    L.n_wit = gs.n_operands + gs.n_wit + (expected_wit ? 1 : 0);
    L.n_cons = gs.n_rows + (with_expected ? 1 : 0);
    L.n_z = L.n_inst + L.n_wit;
    const uint32_t k = gs.n_operands;  // first gadget witness
    const uint32_t expected = b.wcol(k + gs.n_wit);
    switch (gadget) {
      case GADGET_MOD_Q:
      case GADGET_ADD_MOD:
        for (uint32_t i = 0; i < k; i++) b.A(b.wcol(i), b.one);  // <a (+ b) - q t - c | 1 | 0>
        b.A(b.wcol(k), b.minus_q);
        b.A(b.wcol(k + 1), b.minus_one);
        b.B(0, b.one);
        b.end_row();
        b.less_than_q(b.wcol(k + 1), k + 2);
        if (with_expected) {  // out.enforce_equal(expected): <out - expected | 1 | 0>
          b.A(b.wcol(k + 1), b.one);
          b.A(expected, b.minus_one);
          b.B(0, b.one);
          b.end_row();
        }
        break;
      case GADGET_LESS_THAN_Q:
        b.less_than_q(b.wcol(0), 1);
        break;
      case GADGET_LESS_THAN_6144: {
        b.bits_and_decompose(b.wcol(0), 1, 14);
        const uint32_t b11 = b.wcol(1 + 11), b12 = b.wcol(1 + 12), b13 = b.wcol(1 + 13), y1 = b.wcol(15), y2 = b.wcol(16);
        b.A(b11, b.one);  // Not(b12).or(Not(b11)) -> b11.and(b12)
        b.B(b12, b.one);
        b.C(y1, b.one);
        b.end_row();
        b.A(0, b.one);  // Not(b13).and(Not(y1)) -> b13.nor(y1)
        b.A(b13, b.minus_one);
        b.B(0, b.one);
        b.B(y1, b.minus_one);
        b.C(y2, b.one);
        b.end_row();
        if (with_expected) {  // Is(y2).enforce_equal(TRUE): <1 - y2 | 1 | 0>  (no extra witness is used)
          b.A(0, b.one);
          b.A(y2, b.minus_one);
          b.B(0, b.one);
          b.end_row();
        }
        break;
      }
      case GADGET_NORM_BOUND:
        L.w_norm = 1;
        b.norm_rows({b.wcol(0)});
        break;
      case GADGET_NTT_CIRCUIT:
        b.ntt_rows(0, k);
        break;
    }
    b.M.L = L;
    return std::move(b.M);
  }

  Matrices build() {
    M.L = L;
    const uint32_t n = L.n;
    if (L.kind == 1) return build_schoolbook();
    if (L.kind == 2) return build_dual();
    // N x enforce_less_than_q(v[i])   (falcon_ntt.rs:73-77)
    for (uint32_t i = 0; i < n; i++) less_than_q(wcol(L.w_v + i), L.w_vrange + 27 * i);
    ntt_rows(L.w_sig, L.w_nttsig);  // ntt_circuit(sig)  (falcon_ntt.rs:88-89)
    ntt_rows(L.w_v, L.w_nttv);      // ntt_circuit(v)    (falcon_ntt.rs:90-91)
    // pointwise hm_ntt[i] == add_mod(v_ntt[i], sig_ntt[i]*pk_ntt[i])  (falcon_ntt.rs:94-111)
    for (uint32_t i = 0; i < n; i++) {
      uint32_t w = L.w_pw + 30 * i;
      uint32_t sig_ntt = wcol(L.w_nttsig + 29 * i + 1), v_ntt = wcol(L.w_nttv + 29 * i + 1);
      uint32_t p = wcol(w), t = wcol(w + 1), c = wcol(w + 2);
      A(sig_ntt, one);
      B(1 + i, one);  // pk_ntt[i] (instance)
      C(p, one);
      end_row();
      A(v_ntt, one);  // <v_ntt + p - q t - c | 1 | 0>   (arithmetics.rs:248-253)
      A(p, one);
      A(t, minus_q);
      A(c, minus_one);
      B(0, one);
      end_row();
      less_than_q(c, w + 3);
      A(1 + n + i, one);  // <hm_ntt[i] - c | 1 | 0>
      A(c, minus_one);
      B(0, one);
      end_row();
    }
    l2_and_norm_rows([&](uint32_t k) { return wcol(L.w_v + k); });
    return std::move(M);
  }

 private:
  Layout L;
  Matrices M;
  U256 one, minus_one, minus_q;
  std::vector<std::pair<uint32_t, U256>> ra, rb, rc;

  // l2_norm_var over v ++ sig (gadgets/misc.rs:30-51) followed by enforce_less_than_norm_bound
  template <class VCol>
  void l2_and_norm_rows(VCol v_col) {
    const uint32_t n = L.n;
    // l2_norm_var over v ++ sig  (gadgets/misc.rs:30-51, falcon_ntt.rs:116-120)
    std::vector<uint32_t> sq_cols;
    for (uint32_t k = 0; k < 2 * n; k++) {
      uint32_t e = k < n ? v_col(k) : wcol(L.w_sig + (k - n));
      uint32_t w = L.w_l2 + 18 * k;
      bits_and_decompose(e, w, 14);
      uint32_t b11 = wcol(w + 11), b12 = wcol(w + 12), b13 = wcol(w + 13);
      uint32_t y1 = wcol(w + 14), y2 = wcol(w + 15), s = wcol(w + 16), p = wcol(w + 17);
      A(b11, one);  // Not(b12).or(Not(b11)) -> b11.and(b12)
      B(b12, one);
      C(y1, one);
      end_row();
      A(0, one);  // Not(b13).and(Not(y1)) -> b13.nor(y1)
      A(b13, minus_one);
      B(0, one);
      B(y1, minus_one);
      C(y2, one);
      end_row();
      A(y2, one);  // conditionally_select: <y2 | 2e - q | s + e - q>
      B(0, minus_q);
      B(e, u256_small(2));
      C(0, minus_q);
      C(e, one);
      C(s, one);
      end_row();
      A(s, one);
      B(s, one);
      C(p, one);
      end_row();
      sq_cols.push_back(p);
    }
    norm_rows(sq_cols);
  }
  // enforce_less_than_norm_bound(l2)  (range_proofs.rs:100-186 / 192-272); the norm is the inlined sum of sq_cols
  void norm_rows(const std::vector<uint32_t>& sq_cols) {
    {
      NormProgram np = norm_program(L.logn);
      uint32_t w = L.w_norm;
      for (int i = 0; i < np.nbits; i++) booleanity(wcol(w + i));
      for (int i = 0; i < np.nbits; i++) A(wcol(w + i), u256_pow2(i));
      for (uint32_t p : sq_cols) A(p, minus_one);
      B(0, one);
      end_row();
      for (size_t k = 0; k < np.ops.size(); k++) {
        uint32_t a = wcol(w + np.ops[k].a), b = wcol(w + np.ops[k].b), c = wcol(w + np.nbits + (uint32_t)k);
        switch (np.ops[k].kind) {
          case OP_AND:
            A(a, one);
            B(b, one);
            C(c, one);
            break;
          case OP_OR:
            A(0, one);
            A(a, minus_one);
            B(0, one);
            B(b, minus_one);
            C(0, one);
            C(c, minus_one);
            break;
          case OP_AND_NOT:
            A(a, one);
            B(0, one);
            B(b, minus_one);
            C(c, one);
            break;
          case OP_NOR:
            A(0, one);
            A(a, minus_one);
            B(0, one);
            B(b, minus_one);
            C(c, one);
            break;
        }
        end_row();
      }
      A(wcol(w + np.nbits + (uint32_t)np.ops.size() - 1), one);  // Not(c_last).enforce_equal(TRUE)
      B(0, one);
      end_row();
    }
  }

  // DualPolyVar::alloc_vars (gadgets/dual_poly.rs:14-33): witnesses pos[N], neg[N] at w, then N x (product witness,
  // row <pos_i | neg_i | prod_i>), then acc.is_zero().enforce_equal(TRUE) with acc = sum prod_i:
  // is_zero = FpVar::is_eq(acc, 0) -> AllocatedFp(0).is_neq(acc): `ne` (+ booleanity row), `multiplier`,
  // rows <0 - acc | multiplier | ne>, <0 - acc | 1 - ne | 0>; then Not(ne).enforce_equal(TRUE): <ne | 1 | 0>
  void dual_alloc_rows(uint32_t w) {
    const uint32_t n = L.n;
    for (uint32_t i = 0; i < n; i++) {
      A(wcol(w + i), one);
      B(wcol(w + n + i), one);
      C(wcol(w + 2 * n + i), one);
      end_row();
    }
    const uint32_t ne = wcol(w + 3 * n), mult = wcol(w + 3 * n + 1);
    booleanity(ne);
    for (int rep = 0; rep < 2; rep++) {
      for (uint32_t i = 0; i < n; i++) A(wcol(w + 2 * n + i), minus_one);
      if (rep == 0) {
        B(mult, one);
        C(ne, one);
      } else {
        B(0, one);
        B(ne, minus_one);
      }
      end_row();
    }
    A(ne, one);
    B(0, one);
    end_row();
  }
  // circuits/falcon_dual_ntt.rs:26-132
  Matrices build_dual() {
    const uint32_t n = L.n;
    dual_alloc_rows(L.w_sig);  // sig_poly_vars  (:60-61)
    dual_alloc_rows(L.w_v);    // v_vars         (:73)
    // DualNTTPolyVar::ntt_circuit x 2 (dual_poly.rs:42-52): sig.pos, sig.neg, v.pos, v.neg
    const uint32_t in4[4] = {L.w_sig, L.w_sig + n, L.w_v, L.w_v + n};
    for (int k = 0; k < 4; k++) ntt_rows(in4[k], L.w_ntt4 + 29 * n * k);
    auto ntt_out = [&](int k, uint32_t i) { return wcol(L.w_ntt4 + 29 * n * k + 29 * i + 1); };
    // per index: left = mod_q(hm_ntt + v_neg_ntt + sig_neg_ntt * pk_ntt), right = mod_q(v_pos_ntt + sig_pos_ntt * pk_ntt),
    // left == right   (:96-116)
    for (uint32_t i = 0; i < n; i++) {
      uint32_t b_side[2];
      for (int side = 0; side < 2; side++) {
        const uint32_t w = L.w_pw + 60 * i + 30 * side;
        const uint32_t p = wcol(w), t = wcol(w + 1), b = wcol(w + 2);
        A(ntt_out(side == 0 ? 1 : 0, i), one);  // sig.neg (left) / sig.pos (right)
        B(1 + i, one);                          // pk_ntt[i]
        C(p, one);
        end_row();
        if (side == 0) A(1 + n + i, one);       // hm_ntt[i]
        A(ntt_out(side == 0 ? 3 : 2, i), one);  // v.neg (left) / v.pos (right)
        A(p, one);
        A(t, minus_q);
        A(b, minus_one);
        B(0, one);
        end_row();
        less_than_q(b, w + 3);
        b_side[side] = b;
      }
      A(b_side[0], one);
      A(b_side[1], minus_one);
      B(0, one);
      end_row();
    }
    // l2_norm_var_without_range_check over v.pos ++ v.neg ++ sig.pos ++ sig.neg (misc.rs:55-65; :121-129)
    std::vector<uint32_t> sq_cols;
    for (uint32_t k = 0; k < 4 * n; k++) {
      const uint32_t e = wcol(k < 2 * n ? L.w_v + k : L.w_sig + (k - 2 * n));
      const uint32_t p = wcol(L.w_l2 + k);
      A(e, one);
      B(e, one);
      C(p, one);
      end_row();
      sq_cols.push_back(p);
    }
    norm_rows(sq_cols);
    return std::move(M);
  }

  // circuits/falcon_schoolbook.rs:26-132
  Matrices build_schoolbook() {
    const uint32_t n = L.n;
    const U256 q = u256_small(Q);
    auto vcol = [&](uint32_t i) { return wcol(L.w_v + 28 * i); };
    // per i: v[i] then enforce_less_than_q(v[i])   (falcon_schoolbook.rs:86-92)
    for (uint32_t i = 0; i < n; i++) less_than_q(vcol(i), L.w_v + 28 * i + 1);
    // buf = reverse(neg_pk ++ pk); column i takes buf[n-1-i .. 2n-1-i]   (falcon_schoolbook.rs:102-121)
    for (uint32_t i = 0; i < n; i++) {
      const uint32_t w = L.w_cols + L.sb_col * i;
      const uint32_t t = wcol(w), c = wcol(w + 1), m0 = w + 2, rng = w + 2 + n;
      const uint32_t ne1 = wcol(rng + 27), mult1 = wcol(rng + 28), ne2 = wcol(rng + 29), mult2 = wcol(rng + 30),
                     wand = wcol(rng + 31);
      // inner_product_mod (arithmetics.rs:34-100): N products, then <sum m - q t - c | 1 | 0>, then c < q
      for (uint32_t k = 0; k < n; k++) {
        A(wcol(L.w_sig + k), one);
        if (k <= i) {
          B(1 + (i - k), one);  // pk[i-k]
        } else {
          B(0, q);  // neg_pk[n+i-k] = q - pk[n+i-k]
          B(1 + (n + i - k), minus_one);
        }
        C(wcol(m0 + k), one);
        end_row();
      }
      for (uint32_t k = 0; k < n; k++) A(wcol(m0 + k), one);
      A(t, minus_q);
      A(c, minus_one);
      B(0, one);
      end_row();
      less_than_q(c, rng);
      // rhs = hm[i] + q - c; rhs.is_eq(v[i]) and rhs.is_eq(v[i] + q): AllocatedFp::is_neq twice
      for (int which = 0; which < 2; which++) {
        const uint32_t ne = which ? ne2 : ne1, mult = which ? mult2 : mult1;
        booleanity(ne);
        for (int rep = 0; rep < 2; rep++) {
          A(1 + n + i, one);  // d = rhs - v[i] (- q)
          if (which == 0) A(0, q);
          A(c, minus_one);
          A(vcol(i), minus_one);
          if (rep == 0) {  // <d | mult | ne>
            B(mult, one);
            C(ne, one);
          } else {  // <d | 1 - ne | 0>
            B(0, one);
            B(ne, minus_one);
          }
          end_row();
        }
      }
      // Not(ne1).or(Not(ne2)) -> ne2.and(ne1) = w, enforced to be false
      A(ne2, one);
      B(ne1, one);
      C(wand, one);
      end_row();
      A(wand, one);
      B(0, one);
      end_row();
    }
    l2_and_norm_rows(vcol);
    return std::move(M);
  }

  uint32_t wcol(uint32_t w) const { return L.n_inst + w; }
  void A(uint32_t col, const U256& v) { ra.push_back({col, v}); }
  void B(uint32_t col, const U256& v) { rb.push_back({col, v}); }
  void C(uint32_t col, const U256& v) { rc.push_back({col, v}); }
  static void flush(std::vector<std::pair<uint32_t, U256>>& r, HostCSR& m) {
    std::stable_sort(r.begin(), r.end(), [](const auto& x, const auto& y) { return x.first < y.first; });
    for (size_t i = 0; i < r.size();) {
      U256 acc = r[i].second;
      size_t j = i + 1;
      while (j < r.size() && r[j].first == r[i].first) acc = fr_add(acc, r[j++].second);
      if (!u256_is_zero(acc)) {
        m.col.push_back(r[i].first);
        m.val.push_back(acc);
      }
      i = j;
    }
    m.row_ptr.push_back((uint32_t)m.col.size());
    r.clear();
  }
  void end_row() {
    flush(ra, M.a);
    flush(rb, M.b);
    flush(rc, M.c);
  }
  // Boolean::new_witness: <1 - b | b | 0>
  void booleanity(uint32_t b) {
    A(0, one);
    A(b, minus_one);
    B(b, one);
    end_row();
  }
  // k x Boolean::new_witness + enforce_decompose (misc.rs:9-24): <sum 2^i b_i - x | 1 | 0>
  void bits_and_decompose(uint32_t x, uint32_t w, int k) {
    for (int i = 0; i < k; i++) booleanity(wcol(w + i));
    for (int i = 0; i < k; i++) A(wcol(w + i), u256_pow2(i));
    A(x, minus_one);
    B(0, one);
    end_row();
  }
  // enforce_less_than_q (range_proofs.rs:42-94): 27 witnesses at w, 29 rows
  void less_than_q(uint32_t x, uint32_t w) {
    bits_and_decompose(x, w, 14);
    for (int k = 1; k <= 11; k++) {  // kary_or left fold: o_k = o_{k-1} | b_k
      uint32_t prev = k == 1 ? wcol(w) : wcol(w + 14 + (k - 2));
      A(0, one);
      A(prev, minus_one);
      B(0, one);
      B(wcol(w + k), minus_one);
      C(0, one);
      C(wcol(w + 14 + (k - 1)), minus_one);
      end_row();
    }
    uint32_t o11 = wcol(w + 24), x1 = wcol(w + 25), x2 = wcol(w + 26);
    A(o11, one);  // Not(b12).or(Not(o11)) -> o11.and(b12)
    B(wcol(w + 12), one);
    C(x1, one);
    end_row();
    A(x1, one);  // Not(b13).or(Not(x1)) -> x1.and(b13)
    B(wcol(w + 13), one);
    C(x2, one);
    end_row();
    A(x2, one);  // Not(x2).enforce_equal(TRUE)
    B(0, one);
    end_row();
  }
  // ntt_circuit (poly.rs:104-159): N x mod_q(out_k) with out_k the inlined LC
  //   out_k = sum_j M[k][j] x_j + K_k
  //   M[k][j] = prod_{b : bit b of j set} sgn * NTT_TABLE[2^l + (k >> (b+1))], l = logn-1-b,
  //             sgn = -1 if bit b of k is set (the u + (const - v) branch)
  //   K_k     = the network's output on the all-zero input (accumulated 2^(l+1) q^(l+2))
  void ntt_rows(uint32_t w_in, uint32_t w_out) {
    const uint32_t n = L.n, logn = L.logn;
    M.ntt_blocks.push_back({wcol(w_in), wcol(w_out), (uint32_t)M.a.row_ptr.size() - 1});
    std::vector<uint32_t> tab = ntt_table(n);
    // K: run the unreduced network on zeros
    std::vector<U256> kk(n, u256_small(0));
    {
      uint32_t t = n;
      for (uint32_t l = 0; l < logn; l++) {
        U256 cst = u256_pow2(l + 1);
        for (uint32_t e = 0; e < l + 2; e++) cst = u256_mul_small(cst, Q);
        uint32_t m = 1u << l, ht = t / 2, j1 = 0;
        for (uint32_t i = 0; i < m; i++) {
          uint32_t s = tab[m + i];
          for (uint32_t j = j1; j < j1 + ht; j++) {
            U256 u = kk[j], v = u256_mul_small(kk[j + ht], s);
            kk[j] = u256_add(u, v);
            kk[j + ht] = u256_add(u, u256_sub(cst, v));
          }
          j1 += t;
        }
        t = ht;
      }
    }
    std::vector<U256> mag(n);
    std::vector<uint8_t> neg(n);
    for (uint32_t k = 0; k < n; k++) {
      mag[0] = u256_small(1);
      neg[0] = 0;
      for (uint32_t b = 0; b < logn; b++) {
        uint32_t l = logn - 1 - b;
        uint32_t s = tab[(1u << l) + (k >> (b + 1))];
        uint8_t sg = (k >> b) & 1;
        for (uint32_t j = 0; j < (1u << b); j++) {
          mag[j | (1u << b)] = u256_mul_small(mag[j], s);
          neg[j | (1u << b)] = neg[j] ^ sg;
        }
      }
      uint32_t w = w_out + 29 * k;
      A(0, kk[k]);
      for (uint32_t j = 0; j < n; j++) A(wcol(w_in + j), neg[j] ? fr_neg(mag[j]) : mag[j]);
      A(wcol(w), minus_q);         // t
      A(wcol(w + 1), minus_one);   // b
      B(0, one);
      end_row();
      less_than_q(wcol(w + 1), w + 2);
    }
  }
};

}  // namespace circuit
