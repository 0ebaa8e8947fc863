// Subsystem (2): R1CS evaluation  a_i = <A_i, z>, b_i = <B_i, z>, c_i = <C_i, z>  and the
// satisfaction check a_i * b_i == c_i, over the circuit's CSR matrices.  Replaces
// `evaluate_constraint` in ark-groth16 0.3.0's R1CStoQAP::witness_map and
// ark-relations' `which_is_unsatisfied` ([EXT]; reached from examples/pok_sig.rs:32
// and the `cs.is_satisfied()` asserts, e.g. circuits/falcon_ntt.rs:159).
//
// Row classes: ~160k rows have <= 4 non-zeros (one thread per row,
// +-1 coefficients short-cut to add/sub), 2N+1 rows of A have > 600 non-zeros (one warp
// per row, lane-strided, shuffle tree reduction).
#include "ctx.hpp"
#define FF_INLINE_MUL
#include "ff32.cuh"

using ff::Fr;

namespace {

constexpr uint32_t LONG_ROW = 64;

__device__ __forceinline__ Fr load_fr(const uint32_t* p) {
  Fr r;
  uint64_t a, b, c, d;
  asm volatile("ld.global.nc.v4.b64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
  r.v[0] = (uint32_t)a; r.v[1] = (uint32_t)(a >> 32);
  r.v[2] = (uint32_t)b; r.v[3] = (uint32_t)(b >> 32);
  r.v[4] = (uint32_t)c; r.v[5] = (uint32_t)(c >> 32);
  r.v[6] = (uint32_t)d; r.v[7] = (uint32_t)(d >> 32);
  return r;
}
__device__ __forceinline__ void store_fr(uint32_t* p, const Fr& x) {
  uint64_t a = (uint64_t)x.v[0] | ((uint64_t)x.v[1] << 32), b = (uint64_t)x.v[2] | ((uint64_t)x.v[3] << 32);
  uint64_t c = (uint64_t)x.v[4] | ((uint64_t)x.v[5] << 32), d = (uint64_t)x.v[6] | ((uint64_t)x.v[7] << 32);
  asm volatile("st.global.v4.b64 [%0], {%1,%2,%3,%4};" ::"l"(p), "l"(a), "l"(b), "l"(c), "l"(d) : "memory");
}
__device__ __forceinline__ bool is_one(const Fr& c) {
  uint32_t o = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) o |= c.v[i] ^ FrParams::R1(i);
  return o == 0;
}

__global__ void to_montgomery_kernel(uint32_t* vals, uint64_t count) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= count) return;
  Fr x = load_fr(vals + 8 * i);
  store_fr(vals + 8 * i, x.to_mont());
}

// dot product of one CSR row with z, serial
__device__ __forceinline__ Fr row_dot(const uint32_t* __restrict__ row_ptr, const uint32_t* __restrict__ col,
                                      const uint32_t* __restrict__ val, const uint32_t* __restrict__ z, uint32_t row,
                                      const Fr& minus_one) {
  Fr acc = Fr::zero();
  uint32_t k0 = row_ptr[row], k1 = row_ptr[row + 1];
  for (uint32_t k = k0; k < k1; k++) {
    Fr c = load_fr(val + 8 * (uint64_t)k);
    Fr x = load_fr(z + 8 * (uint64_t)col[k]);
    if (is_one(c))
      acc = acc + x;
    else if (c == minus_one)
      acc = acc - x;
    else
      acc = acc + c * x;
  }
  return acc;
}

struct EvalArgs {
  const uint32_t *a_ptr, *a_col, *a_val;
  const uint32_t *b_ptr, *b_col, *b_val;
  const uint32_t *c_ptr, *c_col, *c_val;
  uint32_t n_cons, n_z;
  uint64_t out_stride;  // Fr elements between consecutive signatures in az / bz / cz
};

// one thread per (signature, row) for rows whose A-row is short
__global__ void __launch_bounds__(256)
    r1cs_short_kernel(EvalArgs g, const uint32_t* __restrict__ z_all, uint32_t* az, uint32_t* bz, uint32_t* cz,
                      unsigned long long* first_unsat) {
  uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t sid = blockIdx.y;
  if (row >= g.n_cons) return;
  if (g.a_ptr[row + 1] - g.a_ptr[row] > LONG_ROW || g.b_ptr[row + 1] - g.b_ptr[row] > LONG_ROW ||
      g.c_ptr[row + 1] - g.c_ptr[row] > LONG_ROW)
    return;
  const uint32_t* z = z_all + sid * (uint64_t)g.n_z * 8;
  Fr m1 = Fr::one().neg();
  Fr a = row_dot(g.a_ptr, g.a_col, g.a_val, z, row, m1);
  Fr b = row_dot(g.b_ptr, g.b_col, g.b_val, z, row, m1);
  Fr c = row_dot(g.c_ptr, g.c_col, g.c_val, z, row, m1);
  uint64_t o = (sid * g.out_stride + row) * 8;
  if (az) store_fr(az + o, a);
  if (bz) store_fr(bz + o, b);
  if (cz) store_fr(cz + o, c);
  if (first_unsat && a * b != c) atomicMin(first_unsat + sid, (unsigned long long)row);
}

// lane-strided dot product of one CSR row, reduced over the warp
__device__ __forceinline__ Fr warp_row_dot(const uint32_t* __restrict__ row_ptr, const uint32_t* __restrict__ col,
                                           const uint32_t* __restrict__ val, const uint32_t* __restrict__ z,
                                           uint32_t row, uint32_t lane) {
  Fr acc = Fr::zero();
  uint32_t k0 = row_ptr[row], k1 = row_ptr[row + 1];
  for (uint32_t k = k0 + lane; k < k1; k += 32) {
    Fr c = load_fr(val + 8 * (uint64_t)k);
    Fr x = load_fr(z + 8 * (uint64_t)col[k]);
    acc = acc + c * x;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    Fr other;
#pragma unroll
    for (int i = 0; i < 8; i++) other.v[i] = __shfl_xor_sync(0xffffffffu, acc.v[i], o);
    acc = acc + other;
  }
  return acc;
}

// one warp per (signature, long row): a row is "long" if it has > LONG_ROW non-zeros in any matrix
__global__ void __launch_bounds__(256)
    r1cs_long_kernel(EvalArgs g, const uint32_t* __restrict__ long_rows, uint32_t n_long,
                     const uint32_t* __restrict__ z_all, uint32_t* az, uint32_t* bz, uint32_t* cz,
                     unsigned long long* first_unsat) {
  uint32_t wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  uint64_t sid = blockIdx.y;
  if (wid >= n_long) return;
  uint32_t row = long_rows[wid];
  const uint32_t* z = z_all + sid * (uint64_t)g.n_z * 8;
  Fr a = warp_row_dot(g.a_ptr, g.a_col, g.a_val, z, row, lane);
  Fr b = warp_row_dot(g.b_ptr, g.b_col, g.b_val, z, row, lane);
  Fr c = warp_row_dot(g.c_ptr, g.c_col, g.c_val, z, row, lane);
  if (lane == 0) {
    uint64_t o = (sid * g.out_stride + row) * 8;
    if (az) store_fr(az + o, a);
    if (bz) store_fr(bz + o, b);
    if (cz) store_fr(cz + o, c);
    if (first_unsat && a * b != c) atomicMin(first_unsat + sid, (unsigned long long)row);
  }
}

__global__ void init_unsat_kernel(unsigned long long* p, uint64_t n) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i < n) p[i] = ~0ull;  // == (int64_t)-1 when no row is violated
}

}  // namespace

int32_t launch_to_montgomery(frcs_ctx* ctx, uint32_t* d_vals, uint64_t count, cudaStream_t st) {
  if (count == 0) return FRCS_OK;
  to_montgomery_kernel<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(d_vals, count);
  ctx->launches++;
  FRCS_CUDA_CHECK(cudaGetLastError());
  return FRCS_OK;
}

int32_t launch_r1cs_eval(frcs_ctx* ctx, uint64_t n, const uint64_t* d_z, uint64_t* d_az, uint64_t* d_bz,
                         uint64_t* d_cz, int64_t* d_first_unsat, cudaStream_t st, uint64_t out_stride) {
  if (n == 0) return FRCS_OK;
  if (out_stride == 0) out_stride = ctx->L.n_cons;
  EvalArgs g{ctx->A.row_ptr, ctx->A.col, ctx->A.val, ctx->B.row_ptr, ctx->B.col, ctx->B.val,
             ctx->C.row_ptr, ctx->C.col, ctx->C.val, ctx->L.n_cons,  ctx->L.n_z,     out_stride};
  unsigned long long* fu = (unsigned long long*)d_first_unsat;
  int ph = prof_begin(ctx, PROF_R1CS, st);
  if (fu) {
    init_unsat_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(fu, n);
    ctx->launches++;
  }
  // gridDim.y is limited to 65535: chunk the batch
  for (uint64_t s0 = 0; s0 < n; s0 += 32768) {
    unsigned ny = (unsigned)(n - s0 < 32768 ? n - s0 : 32768);
    const uint32_t* z = (const uint32_t*)d_z + s0 * ctx->L.n_z * 8;
    uint64_t oo = s0 * out_stride * 8;
    uint32_t* az = d_az ? (uint32_t*)d_az + oo : nullptr;
    uint32_t* bz = d_bz ? (uint32_t*)d_bz + oo : nullptr;
    uint32_t* cz = d_cz ? (uint32_t*)d_cz + oo : nullptr;
    dim3 g1((ctx->L.n_cons + 255) / 256, ny);
    r1cs_short_kernel<<<g1, 256, 0, st>>>(g, z, az, bz, cz, fu ? fu + s0 : nullptr);
    ctx->launches++;
    if (ctx->n_long_rows) {
      dim3 g2((ctx->n_long_rows * 32 + 255) / 256, ny);
      r1cs_long_kernel<<<g2, 256, 0, st>>>(g, ctx->long_rows, ctx->n_long_rows, z, az, bz, cz, fu ? fu + s0 : nullptr);
      ctx->launches++;
    }
  }
  prof_end(ctx, ph, st);
  FRCS_CUDA_CHECK(cudaGetLastError());
  return FRCS_OK;
}

// generic y = M x for three CSR matrices sharing the row space (used by the set-up on the
// transposed circuit matrices); long_rows: rows with > 64 non-zeros in any of the three
int32_t launch_matvec3(frcs_ctx* ctx, const DevCSR* m, uint32_t n_rows, const uint32_t* d_long, uint32_t n_long,
                       const uint32_t* d_x, uint32_t* ya, uint32_t* yb, uint32_t* yc, cudaStream_t st) {
  EvalArgs g{m[0].row_ptr, m[0].col, m[0].val, m[1].row_ptr, m[1].col, m[1].val,
             m[2].row_ptr, m[2].col, m[2].val, n_rows,       0,        n_rows};
  r1cs_short_kernel<<<dim3((n_rows + 255) / 256, 1), 256, 0, st>>>(g, d_x, ya, yb, yc, nullptr);
  ctx->launches++;
  if (n_long) {
    r1cs_long_kernel<<<dim3((n_long * 32 + 255) / 256, 1), 256, 0, st>>>(g, d_long, n_long, d_x, ya, yb, yc, nullptr);
    ctx->launches++;
  }
  FRCS_CUDA_CHECK(cudaGetLastError());
  return FRCS_OK;
}
