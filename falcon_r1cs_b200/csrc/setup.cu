// Groth16 parameter generation on the GPU for the context's circuit: what
// Groth16::circuit_specific_setup (examples/pok_sig.rs:30-31) -> ark-groth16 0.3.0
// generate_parameters computes, from explicit toxic waste.
//
//   u_i   = L_i(tau)                       Lagrange coefficients of the radix-2 domain
//   a_j   = sum_i u_i A[i][j]  (+ u_{m+j} for the instance-consistency rows), b_j, c_j likewise
//           (a transposed sparse product: the circuit matrices are transposed once on the host)
//   l_j   = (beta a_j + alpha b_j + c_j) / delta   (witness columns)
//   ic_j  = (beta a_j + alpha b_j + c_j) / gamma   (instance columns -> gamma_abc_g1)
//   h_i   = tau^i Z(tau) / delta
//   a_query = [a_j] G1, b_g1_query = [b_j] G1, b_g2_query = [b_j] G2, h_query, l_query: fixed-base
//           multiplications with an 8-bit window table of the generator (32 mixed additions each),
//           one thread per point, affine output.
// The proving key never leaves the device: the queries go straight into the pre-processed MSM
// tables of the prover.
#include <vector>

#include "ctx.hpp"
#include "nvtx.hpp"
#include "ec.cuh"
#include "msm.hpp"

using namespace ff;

namespace {

template <class F>
struct W {
  static constexpr int N = sizeof(F) / 4;
};
__device__ __forceinline__ Fr ldfr(const uint32_t* p) {
  Fr r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = p[i];
  return r;
}
__device__ __forceinline__ void stfr(uint32_t* p, const Fr& x) {
#pragma unroll
  for (int i = 0; i < 8; i++) p[i] = x.v[i];
}
template <class F>
__device__ __forceinline__ void st_aff(uint32_t* p, const ec::Affine<F>& a) {
  const uint32_t* w = reinterpret_cast<const uint32_t*>(&a);
  for (int i = 0; i < 2 * W<F>::N; i++) p[i] = w[i];
}
template <class F>
__device__ __forceinline__ ec::Affine<F> ld_aff(const uint32_t* p) {
  ec::Affine<F> a;
  uint32_t* w = reinterpret_cast<uint32_t*>(&a);
  for (int i = 0; i < 2 * W<F>::N; i++) w[i] = p[i];
  return a;
}
template <class F> __device__ ec::Affine<F> generator();
template <> __device__ ec::Affine<Fq> generator<Fq>() {
  ec::Affine<Fq> g;
  for (int i = 0; i < 12; i++) {
    g.x.v[i] = FqParams::G1X(i);
    g.y.v[i] = FqParams::G1Y(i);
  }
  return g;
}
template <> __device__ ec::Affine<Fq2> generator<Fq2>() {
  ec::Affine<Fq2> g;
  for (int i = 0; i < 12; i++) {
    g.x.c0.v[i] = FqParams::G2X0(i);
    g.x.c1.v[i] = FqParams::G2X1(i);
    g.y.c0.v[i] = FqParams::G2Y0(i);
    g.y.c1.v[i] = FqParams::G2Y1(i);
  }
  return g;
}

// consts (Fr, Montgomery): [0] alpha [1] beta [2] gamma [3] delta [4] tau [5] g1_scalar [6] g2_scalar
//                          [7] Z(tau) = tau^n - 1   [8] Z(tau)/n   [9] 1/gamma   [10] 1/delta   [11] Z(tau)/delta
__global__ void setup_consts_kernel(uint32_t* c, uint32_t L) {
  Fr tau = ldfr(c + 8 * 4), t = tau;
  for (uint32_t i = 0; i < L; i++) t = t.sqr();
  Fr zt = t - Fr::one();
  Fr ninv = Fr::from_u32(1u << L).inverse();
  Fr gi = ldfr(c + 8 * 2).inverse(), di = ldfr(c + 8 * 3).inverse();
  stfr(c + 8 * 7, zt);
  stfr(c + 8 * 8, zt * ninv);
  stfr(c + 8 * 9, gi);
  stfr(c + 8 * 10, di);
  stfr(c + 8 * 11, zt * di);
}

// u_i = (Z(tau)/n) w^i / (tau - w^i)   (Radix2EvaluationDomain::evaluate_all_lagrange_coefficients, tau outside the domain)
__global__ void lagrange_kernel(const uint32_t* __restrict__ c, const uint32_t* __restrict__ tw, uint32_t n, uint32_t* u) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fr w = i < n / 2 ? ldfr(tw + 8 * (uint64_t)i) : ldfr(tw + 8 * (uint64_t)(i - n / 2)).neg();  // w^(n/2) = -1
  Fr d = ldfr(c + 8 * 4) - w;
  stfr(u + 8 * (uint64_t)i, ldfr(c + 8 * 8) * w * d.inverse());
}

// a_j += u_{m+j} (j < n_inst); then l_j, ic_j; b is left as is
__global__ void combine_kernel(const uint32_t* __restrict__ c, const uint32_t* __restrict__ u, uint32_t* a,
                               const uint32_t* __restrict__ b, const uint32_t* __restrict__ cc, uint32_t n_inst,
                               uint32_t n_wit, uint32_t n_cons, uint32_t* l, uint32_t* ic) {
  uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_inst + n_wit) return;
  Fr aj = ldfr(a + 8 * (uint64_t)j);
  if (j < n_inst) {
    aj = aj + ldfr(u + 8 * (uint64_t)(n_cons + j));
    stfr(a + 8 * (uint64_t)j, aj);
  }
  Fr v = ldfr(c + 8 * 1) * aj + ldfr(c + 8 * 0) * ldfr(b + 8 * (uint64_t)j) + ldfr(cc + 8 * (uint64_t)j);
  if (j < n_inst)
    stfr(ic + 8 * (uint64_t)j, v * ldfr(c + 8 * 9));
  else
    stfr(l + 8 * (uint64_t)(j - n_inst), v * ldfr(c + 8 * 10));
}

// table[w][d] = d * 2^(8w) * (k G), d < 256, w < 32, affine; one thread per entry
template <class F>
__global__ void fixed_table_kernel(const uint32_t* __restrict__ k_mont, uint32_t* table) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 32 * 256) return;
  const uint32_t w = t >> 8, d = t & 255;
  // scalar = d * 2^(8w) * k mod r, canonical
  Fr e = Fr::zero();
  e.v[w >> 2] = d << ((w & 3) * 8);
  Fr sc = (e.to_mont() * ldfr(k_mont)).from_mont();
  ec::XYZZ<F> p = ec::XYZZ<F>::from_affine(generator<F>()).mul(sc.v, 255);
  st_aff<F>(table + (uint64_t)t * 2 * W<F>::N, p.to_affine());
}

// out_i = s_i * (k G): 32 mixed additions from the window table, affine output (zero scalar -> infinity)
template <class F>
__global__ void __launch_bounds__(128) fixed_mul_kernel(const uint32_t* __restrict__ table, const uint32_t* __restrict__ s,
                                                        uint64_t n, uint32_t* out) {
  const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fr k = ldfr(s + 8 * i).from_mont();
  ec::XYZZ<F> acc = ec::XYZZ<F>::infinity();
  for (int w = 0; w < 32; w++) {
    uint32_t d = (k.v[w >> 2] >> ((w & 3) * 8)) & 255;
    if (d) acc.add_mixed(ld_aff<F>(table + ((uint64_t)w * 256 + d) * 2 * W<F>::N));
  }
  st_aff<F>(out + i * 2 * W<F>::N, acc.to_affine());
}

// h_i = tau^i * Z(tau)/delta
__global__ void h_scalars_kernel(const uint32_t* __restrict__ c, uint32_t count, uint32_t* out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  Fr b = ldfr(c + 8 * 4), r = ldfr(c + 8 * 11);
  for (uint32_t e = i; e; e >>= 1) {
    if (e & 1) r = r * b;
    b = b.sqr();
  }
  stfr(out + 8 * (uint64_t)i, r);
}

struct DevBuf {
  void* p = nullptr;
  ~DevBuf() { cudaFree(p); }
  cudaError_t alloc(size_t b) { return cudaMalloc(&p, b ? b : 1); }
  uint32_t* u32() { return (uint32_t*)p; }
};

// transposed CSR of one matrix (columns become rows), canonical values
void transpose(const circuit::HostCSR& h, uint32_t n_cols, circuit::HostCSR* t) {
  const size_t nnz = h.col.size(), n_rows = h.row_ptr.size() - 1;
  t->row_ptr.assign(n_cols + 1, 0);
  for (size_t k = 0; k < nnz; k++) t->row_ptr[h.col[k] + 1]++;
  for (uint32_t c = 0; c < n_cols; c++) t->row_ptr[c + 1] += t->row_ptr[c];
  t->col.resize(nnz);
  t->val.resize(nnz);
  std::vector<uint32_t> cur(t->row_ptr.begin(), t->row_ptr.end() - 1);
  for (size_t r = 0; r < n_rows; r++)
    for (uint32_t k = h.row_ptr[r]; k < h.row_ptr[r + 1]; k++) {
      uint32_t pos = cur[h.col[k]]++;
      t->col[pos] = (uint32_t)r;
      t->val[pos] = h.val[k];
    }
}

int32_t upload_csr_t(frcs_ctx* ctx, const circuit::HostCSR& h, DevCSR* d) {
  d->nnz = h.col.size();
  FRCS_CUDA_CHECK(cudaMalloc(&d->row_ptr, h.row_ptr.size() * 4));
  FRCS_CUDA_CHECK(cudaMalloc(&d->col, (h.col.size() + 1) * 4));
  FRCS_CUDA_CHECK(cudaMalloc(&d->val, (h.val.size() + 1) * 32));
  FRCS_CUDA_CHECK(cudaMemcpy(d->row_ptr, h.row_ptr.data(), h.row_ptr.size() * 4, cudaMemcpyHostToDevice));
  FRCS_CUDA_CHECK(cudaMemcpy(d->col, h.col.data(), h.col.size() * 4, cudaMemcpyHostToDevice));
  FRCS_CUDA_CHECK(cudaMemcpy(d->val, h.val.data(), h.val.size() * 32, cudaMemcpyHostToDevice));
  return launch_to_montgomery(ctx, d->val, d->nnz, ctx->stream);
}

}  // namespace

int32_t get_ntt_plan(frcs_ctx* ctx, uint32_t L, cudaStream_t st, NttPlan** out);
extern "C" int32_t install_pk_from_device(frcs_ctx* ctx, const uint32_t* d_a, const uint32_t* d_b1, const uint32_t* d_b2,
                               const uint32_t* d_h, const uint32_t* d_l, const uint32_t* d_consts_g1 /* alpha, beta, delta */,
                               const uint32_t* d_consts_g2 /* beta, delta */, uint32_t shard, uint32_t n_shards);

extern "C" int32_t frcs_setup(frcs_ctx* ctx, const uint64_t* trapdoor, uint64_t* vk_alpha_g1, uint64_t* vk_g2,
                              uint64_t* gamma_abc_g1) {
  return frcs_setup_shard(ctx, trapdoor, 0, 1, vk_alpha_g1, vk_g2, gamma_abc_g1);
}

extern "C" int32_t frcs_setup_shard(frcs_ctx* ctx, const uint64_t* trapdoor, uint32_t shard, uint32_t n_shards,
                                    uint64_t* vk_alpha_g1, uint64_t* vk_g2, uint64_t* gamma_abc_g1) {
  if (!ctx || !trapdoor || n_shards == 0 || shard >= n_shards) return FRCS_E_INVALID_ARG;
  NvtxRange nvtx("frcs:setup");
  // gamma and delta are inverted (the queries are divided by them): zero is not a valid trapdoor; alpha, beta, tau
  // and the generator scalars must be non-zero for a sound key as well
  for (int k = 0; k < 7; k++)
    if (!(trapdoor[4 * k] | trapdoor[4 * k + 1] | trapdoor[4 * k + 2] | trapdoor[4 * k + 3])) {
      frcs_set_error("frcs_setup: zero trapdoor element");
      return FRCS_E_INVALID_ARG;
    }
  FRCS_CUDA_CHECK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const uint32_t L = ctx->domain_log2, n = 1u << L, ni = ctx->L.n_inst, nw = ctx->L.n_wit, nc = ctx->L.n_cons;
  const uint32_t nv = ni + nw;
  NttPlan* plan;
  int32_t rc = get_ntt_plan(ctx, L, st, &plan);
  if (rc) return rc;
  // transposed circuit matrices (columns -> rows)
  circuit::Builder bld(ctx->L.logn, ctx->L.kind);
  circuit::Matrices m = bld.build();
  struct TGuard {  // the transposed matrices are freed on every path out of this function
    DevCSR T[3];
    ~TGuard() {
      for (int k = 0; k < 3; k++) {
        cudaFree(T[k].row_ptr);
        cudaFree(T[k].col);
        cudaFree(T[k].val);
      }
    }
  } tg;
  DevCSR* T = tg.T;
  std::vector<uint32_t> long_cols;
  {
    circuit::HostCSR t[3];
    transpose(m.a, nv, &t[0]);
    transpose(m.b, nv, &t[1]);
    transpose(m.c, nv, &t[2]);
    for (uint32_t c = 0; c < nv; c++)
      if (t[0].row_ptr[c + 1] - t[0].row_ptr[c] > 64 || t[1].row_ptr[c + 1] - t[1].row_ptr[c] > 64 ||
          t[2].row_ptr[c + 1] - t[2].row_ptr[c] > 64)
        long_cols.push_back(c);
    for (int k = 0; k < 3; k++)
      if ((rc = upload_csr_t(ctx, t[k], &T[k]))) return rc;
  }
  DevBuf consts, u, a, b, c, l, ic, hs, d_long, tab1, tab2, q_a, q_b1, q_b2, q_h, q_l, q_ic, q_c1, q_c2, sc_c;
  FRCS_CUDA_CHECK(consts.alloc(12 * 32));
  FRCS_CUDA_CHECK(u.alloc((size_t)n * 32));
  FRCS_CUDA_CHECK(a.alloc((size_t)nv * 32));
  FRCS_CUDA_CHECK(b.alloc((size_t)nv * 32));
  FRCS_CUDA_CHECK(c.alloc((size_t)nv * 32));
  FRCS_CUDA_CHECK(l.alloc((size_t)nw * 32));
  FRCS_CUDA_CHECK(ic.alloc((size_t)ni * 32));
  FRCS_CUDA_CHECK(hs.alloc((size_t)n * 32));
  FRCS_CUDA_CHECK(d_long.alloc((long_cols.size() + 1) * 4));
  FRCS_CUDA_CHECK(tab1.alloc((size_t)32 * 256 * 96));
  FRCS_CUDA_CHECK(tab2.alloc((size_t)32 * 256 * 192));
  FRCS_CUDA_CHECK(q_a.alloc((size_t)nv * 96));
  FRCS_CUDA_CHECK(q_b1.alloc((size_t)nv * 96));
  FRCS_CUDA_CHECK(q_b2.alloc((size_t)nv * 192));
  FRCS_CUDA_CHECK(q_h.alloc((size_t)n * 96));
  FRCS_CUDA_CHECK(q_l.alloc((size_t)nw * 96));
  FRCS_CUDA_CHECK(q_ic.alloc((size_t)ni * 96));
  FRCS_CUDA_CHECK(q_c1.alloc(3 * 96));
  FRCS_CUDA_CHECK(q_c2.alloc(3 * 192));
  FRCS_CUDA_CHECK(sc_c.alloc(4 * 32));
  FRCS_CUDA_CHECK(cudaMemcpyAsync(consts.p, trapdoor, 7 * 32, cudaMemcpyHostToDevice, st));
  FRCS_CUDA_CHECK(cudaMemcpyAsync(d_long.p, long_cols.data(), long_cols.size() * 4, cudaMemcpyHostToDevice, st));
  FRCS_CUDA_CHECK(cudaDeviceSynchronize());  // legacy-stream uploads of the transposed matrices have landed
  setup_consts_kernel<<<1, 1, 0, st>>>(consts.u32(), L);
  lagrange_kernel<<<(n + 127) / 128, 128, 0, st>>>(consts.u32(), plan->tw_fwd, n, u.u32());
  ctx->launches += 2;
  if ((rc = launch_matvec3(ctx, T, nv, (const uint32_t*)d_long.p, (uint32_t)long_cols.size(), u.u32(), a.u32(), b.u32(),
                           c.u32(), st)))
    return rc;
  combine_kernel<<<(nv + 127) / 128, 128, 0, st>>>(consts.u32(), u.u32(), a.u32(), b.u32(), c.u32(), ni, nw, nc, l.u32(),
                                                   ic.u32());
  h_scalars_kernel<<<(n - 1 + 127) / 128, 128, 0, st>>>(consts.u32(), n - 1, hs.u32());
  fixed_table_kernel<Fq><<<32, 256, 0, st>>>(consts.u32() + 8 * 5, tab1.u32());
  fixed_table_kernel<Fq2><<<32, 256, 0, st>>>(consts.u32() + 8 * 6, tab2.u32());
  auto g1 = [&](const uint32_t* s, uint64_t cnt, uint32_t* out) {
    fixed_mul_kernel<Fq><<<(unsigned)((cnt + 127) / 128), 128, 0, st>>>(tab1.u32(), s, cnt, out);
    ctx->launches++;
  };
  g1(a.u32(), nv, q_a.u32());
  g1(b.u32(), nv, q_b1.u32());
  fixed_mul_kernel<Fq2><<<(nv + 127) / 128, 128, 0, st>>>(tab2.u32(), b.u32(), nv, q_b2.u32());
  g1(hs.u32(), n - 1, q_h.u32());
  g1(l.u32(), nw, q_l.u32());
  g1(ic.u32(), ni, q_ic.u32());
  // constants: G1 alpha, beta, delta (consts 0,1,3); G2 beta, delta, gamma (1,3,2)
  FRCS_CUDA_CHECK(cudaMemcpyAsync(sc_c.u32(), consts.u32(), 64, cudaMemcpyDeviceToDevice, st));                    // alpha, beta
  FRCS_CUDA_CHECK(cudaMemcpyAsync(sc_c.u32() + 16, consts.u32() + 8 * 3, 32, cudaMemcpyDeviceToDevice, st));       // delta
  g1(sc_c.u32(), 3, q_c1.u32());
  FRCS_CUDA_CHECK(cudaMemcpyAsync(sc_c.u32(), consts.u32() + 8, 32, cudaMemcpyDeviceToDevice, st));                // beta
  FRCS_CUDA_CHECK(cudaMemcpyAsync(sc_c.u32() + 8, consts.u32() + 8 * 3, 32, cudaMemcpyDeviceToDevice, st));        // delta
  FRCS_CUDA_CHECK(cudaMemcpyAsync(sc_c.u32() + 16, consts.u32() + 8 * 2, 32, cudaMemcpyDeviceToDevice, st));       // gamma
  fixed_mul_kernel<Fq2><<<1, 128, 0, st>>>(tab2.u32(), sc_c.u32(), 3, q_c2.u32());
  ctx->launches += 6;
  FRCS_CUDA_CHECK(cudaGetLastError());
  // scrub the toxic waste and everything derived from it as scalars (tau powers, Lagrange values, query scalars)
  FRCS_CUDA_CHECK(cudaMemsetAsync(consts.p, 0, 12 * 32, st));
  FRCS_CUDA_CHECK(cudaMemsetAsync(u.p, 0, (size_t)n * 32, st));
  FRCS_CUDA_CHECK(cudaMemsetAsync(a.p, 0, (size_t)nv * 32, st));
  FRCS_CUDA_CHECK(cudaMemsetAsync(b.p, 0, (size_t)nv * 32, st));
  FRCS_CUDA_CHECK(cudaMemsetAsync(c.p, 0, (size_t)nv * 32, st));
  FRCS_CUDA_CHECK(cudaMemsetAsync(l.p, 0, (size_t)nw * 32, st));
  FRCS_CUDA_CHECK(cudaMemsetAsync(ic.p, 0, (size_t)ni * 32, st));
  FRCS_CUDA_CHECK(cudaMemsetAsync(hs.p, 0, (size_t)n * 32, st));
  FRCS_CUDA_CHECK(cudaMemsetAsync(sc_c.p, 0, 4 * 32, st));
  FRCS_CUDA_CHECK(cudaMemsetAsync(tab1.p, 0, (size_t)32 * 256 * 96, st));
  FRCS_CUDA_CHECK(cudaMemsetAsync(tab2.p, 0, (size_t)32 * 256 * 192, st));
  FRCS_CUDA_CHECK(cudaStreamSynchronize(st));
  if (vk_alpha_g1) FRCS_CUDA_CHECK(cudaMemcpy(vk_alpha_g1, q_c1.p, 96, cudaMemcpyDeviceToHost));
  if (vk_g2) {  // beta_g2, gamma_g2, delta_g2
    FRCS_CUDA_CHECK(cudaMemcpy(vk_g2, q_c2.p, 192, cudaMemcpyDeviceToHost));
    FRCS_CUDA_CHECK(cudaMemcpy(vk_g2 + 24, (uint8_t*)q_c2.p + 384, 192, cudaMemcpyDeviceToHost));
    FRCS_CUDA_CHECK(cudaMemcpy(vk_g2 + 48, (uint8_t*)q_c2.p + 192, 192, cudaMemcpyDeviceToHost));
  }
  if (gamma_abc_g1) FRCS_CUDA_CHECK(cudaMemcpy(gamma_abc_g1, q_ic.p, (size_t)ni * 96, cudaMemcpyDeviceToHost));
  return install_pk_from_device(ctx, q_a.u32(), q_b1.u32(), q_b2.u32(), q_h.u32(), q_l.u32(), q_c1.u32(), q_c2.u32(), shard,
                                n_shards);
}
