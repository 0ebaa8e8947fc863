// The reference's gadget entry points (falcon-r1cs/src/gadgets/mod.rs:7-11) as batched CUDA kernels behind the C ABI:
// mod_q, add_mod (gadgets/arithmetics.rs:105-149, 214-262), enforce_less_than_q, is_less_than_6144,
// enforce_less_than_norm_bound (gadgets/range_proofs.rs:42-94, 289-333, 274-284) and NTTPolyVar::ntt_circuit
// (gadgets/poly.rs:104-159).
//
// In the reference a gadget call allocates witnesses in `cs` and enforces rows; `cs.is_satisfied()` then tells whether
// the rows hold (the gadget tests run with the range panics compiled out, range_proofs.rs:55-60).  Here a call takes the
// operand values of n independent instances and returns, per instance, the gadget's witnesses in allocation order and
// the index of the first violated row of the gadget's own rows (-1: satisfied), evaluated on the device from the
// gadget's own A, B, C (circuit::Builder::build_gadget, the same row emitters the full circuits use).  `status` carries
// the code of the panic a non-test build of the reference would hit.
#include <vector>

#include "ctx.hpp"
#define FF_INLINE_MUL
#include "ff32.cuh"
#include "witness_dev.cuh"

using ff::Fr;
using namespace wdev;

namespace {

__device__ __forceinline__ Fr ld_fr(const uint32_t* p) {
  Fr r;
  uint4 a = *reinterpret_cast<const uint4*>(p), b = *reinterpret_cast<const uint4*>(p + 4);
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
// BigUint a / q, a % q on the canonical integer (arithmetics.rs:127-134, 239-246): short division of 8 limbs
__device__ __forceinline__ void divmod_q(const Fr& a, Fr& t, uint32_t& rem) {
  uint64_t r = 0;
#pragma unroll
  for (int k = 7; k >= 0; k--) {
    const uint64_t cur = (r << 32) | a.v[k];
    const uint64_t qk = cur / Q;
    r = cur - qk * Q;
    t.v[k] = (uint32_t)qk;
  }
  rem = (uint32_t)r;
}
__device__ __forceinline__ bool ge_small(const Fr& x, uint32_t bound) {  // canonical x >= bound
  uint32_t hi = 0;
#pragma unroll
  for (int k = 1; k < 8; k++) hi |= x.v[k];
  return hi != 0 || x.v[0] >= bound;
}

struct ScalarArgs {
  int gadget;
  uint32_t n_operands, n_wit, n_wit_total;  // n_wit_total = n_wit (+ 1 for the `expected` witness)
  uint32_t norm_bits, norm_ops, l2_bound;
  NormOpsDev ops;
};

// one thread per instance of a scalar gadget: witnesses in allocation order
__global__ void __launch_bounds__(128)
    gadget_scalar_kernel(ScalarArgs g, uint64_t n, const uint32_t* __restrict__ operands,
                         const uint32_t* __restrict__ expected, uint64_t* __restrict__ wit, int32_t* __restrict__ status) {
  const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t* w = wit + i * g.n_wit_total * 4;
  const Fr a = ld_fr(operands + i * g.n_operands * 8);
  int st = FRCS_OK;
  if (g.gadget == circuit::GADGET_MOD_Q || g.gadget == circuit::GADGET_ADD_MOD) {
    // t = x / q, c = x % q for x = a (mod_q) or the field sum a + b (add_mod); then enforce_less_than_q(c)
    Fr x = a;
    if (g.gadget == circuit::GADGET_ADD_MOD) x = x + ld_fr(operands + (i * g.n_operands + 1) * 8);
    x = x.from_mont();
    Fr t = Fr::zero();
    uint32_t c;
    divmod_q(x, t, c);
    store_fr(w, t.to_mont());
    store_fr(w + 4, Fr::from_u32(c));
    const uint32_t m = ltq_mask(c);
    for (uint32_t j = 0; j < 27; j++) store_bit(w + 4 * (2 + j), (m >> j) & 1u);
    if (expected) store_fr(w + 4 * 29, ld_fr(expected + i * 8));
  } else if (g.gadget == circuit::GADGET_LESS_THAN_Q || g.gadget == circuit::GADGET_LESS_THAN_6144) {
    // a.into_repr().to_bits_le() truncated to 14 bits (range_proofs.rs:62-69, 301-308), then the closed forms
    const Fr x = a.from_mont();
    const uint32_t low = x.v[0] & 0x3fffu;
    if (g.gadget == circuit::GADGET_LESS_THAN_Q) {
      if (ge_small(x, Q)) st = FRCS_E_COEFF_RANGE;  // range_proofs.rs:58-60
      const uint32_t m = ltq_mask(low);
      for (uint32_t j = 0; j < 27; j++) store_bit(w + 4 * j, (m >> j) & 1u);
    } else {
      const uint32_t m = l2_mask(low);
      for (uint32_t j = 0; j < 16; j++) store_bit(w + 4 * j, (m >> j) & 1u);
    }
  } else {  // enforce_less_than_norm_bound (range_proofs.rs:100-186 / 192-272): bits, then the and / or chain
    const Fr x = a.from_mont();
    if (ge_small(x, g.l2_bound)) st = FRCS_E_NORM_BOUND;  // range_proofs.rs:114-117, 205-208
    uint32_t b[64];
    for (uint32_t j = 0; j < g.norm_bits; j++) b[j] = (x.v[0] >> j) & 1u;
    for (uint32_t k = 0; k < g.norm_ops; k++) {
      const uint32_t p = b[g.ops.a[k]], q = b[g.ops.b[k]];
      uint32_t r;
      switch (g.ops.kind[k]) {
        case circuit::OP_AND: r = p & q; break;
        case circuit::OP_OR: r = p | q; break;
        case circuit::OP_AND_NOT: r = p & (q ^ 1u); break;
        default: r = (p ^ 1u) & (q ^ 1u); break;
      }
      b[g.norm_bits + k] = r;
    }
    for (uint32_t j = 0; j < g.norm_bits + g.norm_ops; j++) store_bit(w + 4 * j, b[j]);
  }
  status[i] = st;
}

// NTTPolyVar::ntt_circuit (poly.rs:104-159), one instance per block: LOG_N layers of unreduced integer butterflies
// (five 32-bit limbs: values stay below 2^160), then per output mod_q: t, b and the 27 range
// witnesses of b.  `fr_in` receives the N input coefficients as field elements (the operand columns of the row check).
struct NttArgs {
  uint32_t cst[11][5];  // 2^(l+1) q^(l+2)
};
template <int LOGN>
__global__ void __launch_bounds__(256)
    gadget_ntt_kernel(NttArgs P, const uint16_t* __restrict__ poly, const uint32_t* __restrict__ g_tab,
                      uint64_t* __restrict__ fr_in, uint64_t* __restrict__ wit, uint16_t* __restrict__ values,
                      int32_t* __restrict__ status) {
  constexpr int N = 1 << LOGN, NT = 256;
  extern __shared__ uint32_t sm[];
  uint32_t* s_tab = sm;           // [N]
  uint32_t* s_lazy = s_tab + N;   // [5][N]
  __shared__ int s_bad;
  const int tid = threadIdx.x;
  const uint64_t inst = blockIdx.x;
  if (tid == 0) s_bad = 0;
  __syncthreads();
  int bad = 0;
  for (int i = tid; i < N; i += NT) {
    s_tab[i] = g_tab[i];
    const uint32_t x = poly[inst * N + i];
    bad |= x >= Q;
    s_lazy[i] = modq(x);
#pragma unroll
    for (int k = 1; k < 5; k++) s_lazy[k * N + i] = 0;
    store_fr(fr_in + (inst * N + i) * 4, Fr::from_u32(modq(x)));
  }
  if (bad) s_bad = 1;
  __syncthreads();
  int t = N;
#pragma unroll 1
  for (int l = 0; l < LOGN; l++) {
    const int ht = t >> 1;
    for (int idx = tid; idx < N / 2; idx += NT) {
      const int i = idx / ht, j = idx - i * ht;
      const int p0 = i * t + j, p1 = p0 + ht;
      const uint32_t s = s_tab[(1 << l) + i];
      uint32_t u[5], sv[5];
      uint64_t c = 0;
#pragma unroll
      for (int k = 0; k < 5; k++) {
        u[k] = s_lazy[k * N + p0];
        c += (uint64_t)s_lazy[k * N + p1] * s;
        sv[k] = (uint32_t)c;
        c >>= 32;
      }
      uint64_t ca = 0, cb = 0;
      int64_t br = 0;
#pragma unroll
      for (int k = 0; k < 5; k++) {  // out[j] = u + v ; out[j + ht] = u + (const[l+1] - v)
        ca += (uint64_t)u[k] + sv[k];
        s_lazy[k * N + p0] = (uint32_t)ca;
        ca >>= 32;
        const int64_t d = (int64_t)P.cst[l][k] - sv[k] - br;
        br = d < 0;
        cb += (uint64_t)u[k] + (uint32_t)d;
        s_lazy[k * N + p1] = (uint32_t)cb;
        cb >>= 32;
      }
    }
    t = ht;
    __syncthreads();
  }
  for (int i = tid; i < N; i += NT) {
    uint64_t rem = 0;
    Fr q = Fr::zero();
#pragma unroll
    for (int k = 4; k >= 0; k--) {
      const uint64_t cur = (rem << 32) | s_lazy[k * N + i];
      const uint64_t qk = cur / Q;
      rem = cur - qk * Q;
      q.v[k] = (uint32_t)qk;
    }
    uint64_t* w = wit + (inst * N + i) * 29 * 4;
    store_fr(w, q.to_mont());
    store_fr(w + 4, Fr::from_u32((uint32_t)rem));
    const uint32_t m = ltq_mask((uint32_t)rem);
    for (uint32_t j = 0; j < 27; j++) store_bit(w + 4 * (2 + j), (m >> j) & 1u);
    if (values) values[inst * N + i] = (uint16_t)rem;
  }
  if (tid == 0) status[inst] = s_bad ? FRCS_E_COEFF_RANGE : FRCS_OK;
}

// cs.which_is_unsatisfied() over the gadget's rows: one block per instance, threads stride over the rows;
// z column 0 = One, 1 .. k = operands, then the witnesses
struct RowArgs {
  const uint32_t *a_ptr, *a_col, *a_val, *b_ptr, *b_col, *b_val, *c_ptr, *c_col, *c_val;
  uint32_t n_rows, n_operands, n_wit_total;
};
__device__ __forceinline__ Fr gadget_dot(const uint32_t* __restrict__ ptr, const uint32_t* __restrict__ col,
                                         const uint32_t* __restrict__ val, uint32_t row, const uint32_t* ops,
                                         const uint32_t* wit, uint32_t k) {
  Fr acc = Fr::zero();
  for (uint32_t e = ptr[row]; e < ptr[row + 1]; e++) {
    const uint32_t c = col[e];
    const Fr x = c == 0 ? Fr::one() : c <= k ? ld_fr(ops + (uint64_t)(c - 1) * 8) : ld_fr(wit + (uint64_t)(c - 1 - k) * 8);
    acc = acc + ld_fr(val + (uint64_t)e * 8) * x;
  }
  return acc;
}
__global__ void __launch_bounds__(256)
    gadget_rows_kernel(RowArgs g, const uint32_t* __restrict__ operands, const uint32_t* __restrict__ wit,
                       unsigned long long* __restrict__ first_unsat) {
  const uint64_t inst = blockIdx.x;
  const uint32_t* ops = operands + inst * g.n_operands * 8;
  const uint32_t* w = wit + inst * g.n_wit_total * 8;
  for (uint32_t row = threadIdx.x; row < g.n_rows; row += blockDim.x) {
    const Fr a = gadget_dot(g.a_ptr, g.a_col, g.a_val, row, ops, w, g.n_operands);
    const Fr b = gadget_dot(g.b_ptr, g.b_col, g.b_val, row, ops, w, g.n_operands);
    const Fr c = gadget_dot(g.c_ptr, g.c_col, g.c_val, row, ops, w, g.n_operands);
    if (a * b != c) atomicMin(first_unsat + inst, (unsigned long long)row);
  }
}
__global__ void gadget_init_unsat_kernel(unsigned long long* p, uint64_t n) {
  const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i < n) p[i] = ~0ull;
}

struct GBuf {  // RAII device allocation
  void* p = nullptr;
  ~GBuf() {
    if (p) cudaFree(p);
  }
  cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1); }
};

int32_t upload_gadget_csr(frcs_ctx* ctx, const circuit::HostCSR& h, DevCSR* d) {
  d->nnz = h.col.size();
  FRCS_CUDA_CHECK(cudaMalloc(&d->row_ptr, h.row_ptr.size() * 4));
  FRCS_CUDA_CHECK(cudaMalloc(&d->col, (h.col.size() + 1) * 4));
  FRCS_CUDA_CHECK(cudaMalloc(&d->val, (h.val.size() + 1) * 32));
  FRCS_CUDA_CHECK(cudaMemcpyAsync(d->row_ptr, h.row_ptr.data(), h.row_ptr.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
  FRCS_CUDA_CHECK(cudaMemcpyAsync(d->col, h.col.data(), h.col.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
  FRCS_CUDA_CHECK(cudaMemcpyAsync(d->val, h.val.data(), h.val.size() * 32, cudaMemcpyHostToDevice, ctx->stream));
  FRCS_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));  // the host vectors die with the caller's Matrices
  return launch_to_montgomery(ctx, d->val, d->nnz, ctx->stream);
}

// the gadget's A, B, C on the device (built once per context, gadget and variant)
int32_t ensure_gadget(frcs_ctx* ctx, int gadget, bool with_expected, DevGadget** out) {
  DevGadget& G = ctx->gadgets[gadget][with_expected ? 1 : 0];
  if (!G.ready) {
    circuit::Matrices m = circuit::Builder::build_gadget(ctx->L.logn, gadget, with_expected);
    int32_t rc;
    if ((rc = upload_gadget_csr(ctx, m.a, &G.m[0])) || (rc = upload_gadget_csr(ctx, m.b, &G.m[1])) ||
        (rc = upload_gadget_csr(ctx, m.c, &G.m[2])))
      return rc;
    G.n_rows = m.L.n_cons;
    G.n_wit_total = m.L.n_wit - circuit::gadget_shape(gadget, ctx->L.logn).n_operands;
    G.ready = true;
  }
  *out = &G;
  return FRCS_OK;
}

int32_t check_rows(frcs_ctx* ctx, const DevGadget& G, uint32_t n_operands, uint64_t n, const uint32_t* d_ops,
                   const uint32_t* d_wit, int64_t* d_fu, cudaStream_t st) {
  RowArgs ra{G.m[0].row_ptr, G.m[0].col, G.m[0].val, G.m[1].row_ptr, G.m[1].col, G.m[1].val,
             G.m[2].row_ptr, G.m[2].col, G.m[2].val, G.n_rows, n_operands, G.n_wit_total};
  gadget_init_unsat_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>((unsigned long long*)d_fu, n);
  gadget_rows_kernel<<<(unsigned)n, 256, 0, st>>>(ra, d_ops, d_wit, (unsigned long long*)d_fu);
  ctx->launches += 2;
  FRCS_CUDA_CHECK(cudaGetLastError());
  return FRCS_OK;
}

// scalar gadgets: operands n x k Fr (host), expected n Fr or NULL
int32_t run_scalar(frcs_ctx* ctx, int gadget, uint64_t n, const uint64_t* operands, const uint64_t* expected,
                   bool with_expected, uint64_t* wit_out, int64_t* first_unsat, int32_t* status) {
  if (!ctx || !operands || !wit_out || !first_unsat || !status) return FRCS_E_INVALID_ARG;
  if (n == 0) return FRCS_OK;
  if (n > (1u << 24)) {
    frcs_set_error("frcs_gadget_*: at most 2^24 instances per call");
    return FRCS_E_INVALID_ARG;
  }
  FRCS_CUDA_CHECK(cudaSetDevice(ctx->device));
  const circuit::GadgetShape gs = circuit::gadget_shape(gadget, ctx->L.logn);
  DevGadget* G;
  int32_t rc = ensure_gadget(ctx, gadget, with_expected, &G);
  if (rc) return rc;
  cudaStream_t st = ctx->stream;
  GBuf d_ops, d_exp, d_wit, d_fu, d_st;
  const size_t ob = n * gs.n_operands * 32, wb = n * (size_t)G->n_wit_total * 32;
  FRCS_CUDA_CHECK(d_ops.alloc(ob));
  FRCS_CUDA_CHECK(d_wit.alloc(wb));
  FRCS_CUDA_CHECK(d_fu.alloc(n * 8));
  FRCS_CUDA_CHECK(d_st.alloc(n * 4));
  FRCS_CUDA_CHECK(cudaMemcpyAsync(d_ops.p, operands, ob, cudaMemcpyHostToDevice, st));
  const bool exp_wit = expected && G->n_wit_total > gs.n_wit;
  if (exp_wit) {
    FRCS_CUDA_CHECK(d_exp.alloc(n * 32));
    FRCS_CUDA_CHECK(cudaMemcpyAsync(d_exp.p, expected, n * 32, cudaMemcpyHostToDevice, st));
  }
  ScalarArgs sa;
  sa.gadget = gadget;
  sa.n_operands = gs.n_operands;
  sa.n_wit = gs.n_wit;
  sa.n_wit_total = G->n_wit_total;
  sa.norm_bits = ctx->L.norm_bits;
  sa.norm_ops = ctx->L.norm_ops;
  sa.l2_bound = ctx->L.l2_bound;
  sa.ops = ctx->norm_ops;
  gadget_scalar_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(sa, n, (const uint32_t*)d_ops.p,
                                                                   exp_wit ? (const uint32_t*)d_exp.p : nullptr,
                                                                   (uint64_t*)d_wit.p, (int32_t*)d_st.p);
  ctx->launches++;
  rc = check_rows(ctx, *G, gs.n_operands, n, (const uint32_t*)d_ops.p, (const uint32_t*)d_wit.p, (int64_t*)d_fu.p, st);
  if (!rc) {
    FRCS_CUDA_CHECK(cudaMemcpyAsync(wit_out, d_wit.p, wb, cudaMemcpyDeviceToHost, st));
    FRCS_CUDA_CHECK(cudaMemcpyAsync(first_unsat, d_fu.p, n * 8, cudaMemcpyDeviceToHost, st));
    FRCS_CUDA_CHECK(cudaMemcpyAsync(status, d_st.p, n * 4, cudaMemcpyDeviceToHost, st));
  }
  FRCS_CUDA_CHECK(cudaStreamSynchronize(st));
  return rc;
}

}  // namespace

extern "C" {

int32_t frcs_gadget_shape(const frcs_ctx* ctx, int32_t gadget, uint32_t* n_operands, uint32_t* n_witness, uint32_t* n_rows) {
  if (!ctx || gadget < 0 || gadget >= circuit::GADGET_COUNT) return FRCS_E_INVALID_ARG;
  const circuit::GadgetShape gs = circuit::gadget_shape(gadget, ctx->L.logn);
  if (n_operands) *n_operands = gs.n_operands;
  if (n_witness) *n_witness = gs.n_wit;
  if (n_rows) *n_rows = gs.n_rows;
  return FRCS_OK;
}

int32_t frcs_gadget_mod_q(frcs_ctx* ctx, uint64_t n, const uint64_t* a, const uint64_t* expected, uint64_t* wit,
                          int64_t* first_unsat, int32_t* status) {
  return run_scalar(ctx, circuit::GADGET_MOD_Q, n, a, expected, expected != nullptr, wit, first_unsat, status);
}
int32_t frcs_gadget_add_mod(frcs_ctx* ctx, uint64_t n, const uint64_t* ab, const uint64_t* expected, uint64_t* wit,
                            int64_t* first_unsat, int32_t* status) {
  return run_scalar(ctx, circuit::GADGET_ADD_MOD, n, ab, expected, expected != nullptr, wit, first_unsat, status);
}
int32_t frcs_gadget_less_than_q(frcs_ctx* ctx, uint64_t n, const uint64_t* a, uint64_t* wit, int64_t* first_unsat,
                                int32_t* status) {
  return run_scalar(ctx, circuit::GADGET_LESS_THAN_Q, n, a, nullptr, false, wit, first_unsat, status);
}
int32_t frcs_gadget_less_than_6144(frcs_ctx* ctx, uint64_t n, const uint64_t* a, int32_t enforce_true, uint64_t* wit,
                                   int64_t* first_unsat, int32_t* status) {
  return run_scalar(ctx, circuit::GADGET_LESS_THAN_6144, n, a, nullptr, enforce_true != 0, wit, first_unsat, status);
}
int32_t frcs_gadget_norm_bound(frcs_ctx* ctx, uint64_t n, const uint64_t* a, uint64_t* wit, int64_t* first_unsat,
                               int32_t* status) {
  return run_scalar(ctx, circuit::GADGET_NORM_BOUND, n, a, nullptr, false, wit, first_unsat, status);
}

int32_t frcs_gadget_ntt_circuit(frcs_ctx* ctx, uint64_t n, const uint16_t* poly, uint16_t* values, uint64_t* wit,
                                int64_t* first_unsat, int32_t* status) {
  if (!ctx || !poly || !first_unsat || !status) return FRCS_E_INVALID_ARG;
  if (n == 0) return FRCS_OK;
  if (n > 65535) {
    frcs_set_error("frcs_gadget_ntt_circuit: at most 65535 polynomials per call");
    return FRCS_E_INVALID_ARG;
  }
  FRCS_CUDA_CHECK(cudaSetDevice(ctx->device));
  const uint32_t logn = ctx->L.logn, N = ctx->L.n;
  DevGadget* G;
  int32_t rc = ensure_gadget(ctx, circuit::GADGET_NTT_CIRCUIT, false, &G);
  if (rc) return rc;
  cudaStream_t st = ctx->stream;
  GBuf d_poly, d_fr, d_wit, d_val, d_fu, d_st;
  const size_t wb = n * (size_t)29 * N * 32;
  FRCS_CUDA_CHECK(d_poly.alloc(n * N * 2));
  FRCS_CUDA_CHECK(d_fr.alloc(n * N * 32));
  FRCS_CUDA_CHECK(d_wit.alloc(wb));
  FRCS_CUDA_CHECK(d_val.alloc(n * N * 2));
  FRCS_CUDA_CHECK(d_fu.alloc(n * 8));
  FRCS_CUDA_CHECK(d_st.alloc(n * 4));
  FRCS_CUDA_CHECK(cudaMemcpyAsync(d_poly.p, poly, n * N * 2, cudaMemcpyHostToDevice, st));
  NttArgs P;
  for (uint32_t l = 0; l < logn; l++) {  // 2^(l+1) q^(l+2)  (falcon_ntt.rs:31-39)
    circuit::U256 c = circuit::u256_pow2(l + 1);
    for (uint32_t e = 0; e < l + 2; e++) c = circuit::u256_mul_small(c, circuit::Q);
    for (int k = 0; k < 5; k++) P.cst[l][k] = c.v[k];
  }
  const size_t smem = (size_t)6 * N * 4;
  if (logn == 10)
    gadget_ntt_kernel<10><<<(unsigned)n, 256, smem, st>>>(P, (const uint16_t*)d_poly.p, ctx->ntt_tab, (uint64_t*)d_fr.p,
                                                         (uint64_t*)d_wit.p, (uint16_t*)d_val.p, (int32_t*)d_st.p);
  else
    gadget_ntt_kernel<9><<<(unsigned)n, 256, smem, st>>>(P, (const uint16_t*)d_poly.p, ctx->ntt_tab, (uint64_t*)d_fr.p,
                                                        (uint64_t*)d_wit.p, (uint16_t*)d_val.p, (int32_t*)d_st.p);
  ctx->launches++;
  rc = check_rows(ctx, *G, N, n, (const uint32_t*)d_fr.p, (const uint32_t*)d_wit.p, (int64_t*)d_fu.p, st);
  if (!rc) {
    if (wit) FRCS_CUDA_CHECK(cudaMemcpyAsync(wit, d_wit.p, wb, cudaMemcpyDeviceToHost, st));
    if (values) FRCS_CUDA_CHECK(cudaMemcpyAsync(values, d_val.p, n * N * 2, cudaMemcpyDeviceToHost, st));
    FRCS_CUDA_CHECK(cudaMemcpyAsync(first_unsat, d_fu.p, n * 8, cudaMemcpyDeviceToHost, st));
    FRCS_CUDA_CHECK(cudaMemcpyAsync(status, d_st.p, n * 4, cudaMemcpyDeviceToHost, st));
  }
  FRCS_CUDA_CHECK(cudaStreamSynchronize(st));
  return rc;
}

}  // extern "C"
