// The whole path: ark_groth16::create_proof(circuit, pk, r, s) (ark-groth16 0.3.0
// prover.rs, called by create_random_proof at examples/pok_sig.rs:32), batched.
//
//   witness generation (witness.cu)  ->  z
//   R1CS evaluation + witness map (spmv.cu, ntt.cu)  ->  h
//   five MSMs on five streams (msm_impl.cuh):
//       A  = sum z_i a_query[i] + alpha + r delta_1          (calculate_coeff)
//       B1 = sum z_i b_g1_query[i] + beta_1 + s delta_1
//       B2 = sum z_i b_g2_query[i] + beta_2 + s delta_2
//       L  = sum w_j l_query[j] - (r s) delta_1
//       H  = sum h_i h_query[i]
//     The constant points and the r/s multiples of delta are folded into the MSMs as extra
//     bases (alpha, beta with scalar 1; delta with scalar r, s or -rs): no separate
//     scalar-multiplication kernels.
//   host tail (finalize.cpp): C = s A + r B1 + L + H, three points to affine.
#include <chrono>
#include <cstdlib>

#include "ctx.hpp"
#define FF_INLINE_MUL
#include "ff32.cuh"
#include "finalize.hpp"
#include "msm.hpp"

using ff::Fq;
using ff::Fq2;
using ff::Fr;

namespace {

void timed_finalize(frcs_ctx* ctx, const uint64_t* msm, const uint64_t* r, const uint64_t* s, uint64_t* proof) {
  auto t0 = std::chrono::steady_clock::now();
  host_finalize_proof(msm, r, s, proof);
  if (ctx->prof.on) {
    ctx->prof.ms[PROF_HOST_TAIL] += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    ctx->prof.count[PROF_HOST_TAIL]++;
  }
}

__device__ __forceinline__ Fr ld_fr(const uint32_t* p) {
  Fr r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = p[i];
  return r;
}
__device__ __forceinline__ void st_fr(uint32_t* p, const Fr& x) {
#pragma unroll
  for (int i = 0; i < 8; i++) p[i] = x.v[i];
}
// extra scalars (Montgomery): ex[0..1] = {1, r}; ex[2..3] = {1, s}; ex[4] = -(r s)
__global__ void extras_kernel(const uint32_t* r, const uint32_t* s, uint32_t* ex) {
  Fr rr = ld_fr(r), ss = ld_fr(s);
  st_fr(ex, Fr::one());
  st_fr(ex + 8, rr);
  st_fr(ex + 16, Fr::one());
  st_fr(ex + 24, ss);
  st_fr(ex + 32, (rr * ss).neg());
}

// IMAD.WIDE peak: 8 independent 64-bit accumulators per thread, mad.wide.u32 in a tight loop
__global__ void __launch_bounds__(256) imad_peak_kernel(uint64_t* out, uint32_t iters, uint32_t seed) {
  uint32_t a = seed + threadIdx.x, b = seed * 2654435761u + blockIdx.x;
  uint64_t acc[8];
#pragma unroll
  for (int k = 0; k < 8; k++) acc[k] = k;
  for (uint32_t i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 8; k++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[k]) : "r"(a), "r"(b));
    a += 1;
  }
  uint64_t s = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) s ^= acc[k];
  if (s == 0x1234567) out[0] = s;  // keep the chain alive
}

struct PkDev {
  uint64_t n_a, n_b1, n_b2, n_l, n_h;  // base counts incl. the appended constants
};

int32_t upload_and_precompute_g1(frcs_ctx* ctx, const uint64_t* q, uint64_t len, const uint64_t* const* extra,
                                 int n_extra, DevBases* out) {
  const uint64_t n = len + n_extra;
  uint32_t* d_in = nullptr;
  FRCS_CUDA_CHECK(cudaMalloc(&d_in, n * 96));
  FRCS_CUDA_CHECK(cudaMemcpy(d_in, q, len * 96, cudaMemcpyHostToDevice));
  for (int e = 0; e < n_extra; e++)
    FRCS_CUDA_CHECK(cudaMemcpy((uint8_t*)d_in + (len + e) * 96, extra[e], 96, cudaMemcpyHostToDevice));
  FRCS_CUDA_CHECK(cudaMalloc(&out->pts, n * 96 * MSM_WINDOWS));
  out->n = n;
  out->windows = MSM_WINDOWS;
  out->g2 = false;
  int32_t rc = msm_precompute<Fq>(ctx, d_in, n, (uint32_t*)out->pts, ctx->stream);
  FRCS_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  cudaFree(d_in);
  return rc;
}
int32_t upload_and_precompute_g2(frcs_ctx* ctx, const uint64_t* q, uint64_t len, const uint64_t* const* extra,
                                 int n_extra, DevBases* out) {
  const uint64_t n = len + n_extra;
  uint32_t* d_in = nullptr;
  FRCS_CUDA_CHECK(cudaMalloc(&d_in, n * 192));
  FRCS_CUDA_CHECK(cudaMemcpy(d_in, q, len * 192, cudaMemcpyHostToDevice));
  for (int e = 0; e < n_extra; e++)
    FRCS_CUDA_CHECK(cudaMemcpy((uint8_t*)d_in + (len + e) * 192, extra[e], 192, cudaMemcpyHostToDevice));
  FRCS_CUDA_CHECK(cudaMalloc(&out->pts, n * 192 * MSM_WINDOWS));
  out->n = n;
  out->windows = MSM_WINDOWS;
  out->g2 = true;
  int32_t rc = msm_precompute<Fq2>(ctx, d_in, n, (uint32_t*)out->pts, ctx->stream);
  FRCS_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  cudaFree(d_in);
  return rc;
}

// device buffers of the proving pipeline, allocated once per context
int32_t ensure_prover(frcs_ctx* ctx) {
  if (ctx->prover_ready) return FRCS_OK;
  ProverState& P = ctx->prover;
  const uint64_t n = 1ull << ctx->domain_log2;
  for (int i = 0; i < 5; i++) FRCS_CUDA_CHECK(cudaStreamCreateWithFlags(&P.streams[i], cudaStreamNonBlocking));
  for (int i = 0; i < 5; i++) FRCS_CUDA_CHECK(cudaEventCreateWithFlags(&P.done[i], cudaEventDisableTiming));
  FRCS_CUDA_CHECK(cudaEventCreateWithFlags(&P.fork, cudaEventDisableTiming));
  for (int i = 0; i < 2; i++) FRCS_CUDA_CHECK(cudaEventCreateWithFlags(&P.copied[i], cudaEventDisableTiming));
  FRCS_CUDA_CHECK(cudaMalloc(&P.ntt_work, 3 * n * 32));
  FRCS_CUDA_CHECK(cudaMalloc(&P.h, n * 32));
  FRCS_CUDA_CHECK(cudaMalloc(&P.extras, 2 * 5 * 32));
  FRCS_CUDA_CHECK(cudaMalloc(&P.results, 2 * PROOF_MSM_WORDS * 8));
  FRCS_CUDA_CHECK(cudaMallocHost(&P.h_results, 2 * PROOF_MSM_WORDS * 8));
  const uint64_t sizes[5] = {ctx->pk_a.n, ctx->pk_b1.n, ctx->pk_l.n, ctx->pk_h.n, ctx->pk_b2.n};
  for (int i = 0; i < 5; i++) {
    size_t b = i == 4 ? msm_work_bytes<Fq2>(sizes[i]) : msm_work_bytes<Fq>(sizes[i]);
    FRCS_CUDA_CHECK(cudaMalloc(&P.msm_work[i], b));
  }
  ctx->prover_ready = true;
  return FRCS_OK;
}

// Launches everything for one proof whose assignment z is on the device.  Results land in
// pinned host slot `slot` (event copied[slot]).
int32_t launch_proof(frcs_ctx* ctx, const uint64_t* d_z, const uint32_t* d_r, const uint32_t* d_s, int slot,
                     cudaStream_t st) {
  ProverState& P = ctx->prover;
  const uint64_t n_inst = ctx->L.n_inst, n_wit = ctx->L.n_wit, nv = n_inst + n_wit;
  const uint64_t n = 1ull << ctx->domain_log2;
  int32_t rc = launch_witness_map(ctx, d_z, (uint64_t*)P.h, (uint32_t*)P.ntt_work, st);
  if (rc) return rc;
  uint32_t* ex = (uint32_t*)P.extras + slot * 40;
  extras_kernel<<<1, 1, 0, st>>>(d_r, d_s, ex);
  ctx->launches++;
  FRCS_CUDA_CHECK(cudaEventRecord(P.fork, st));
  uint32_t* res = (uint32_t*)P.results + (size_t)slot * PROOF_MSM_WORDS * 2;
  const uint32_t* z32 = (const uint32_t*)d_z;
  for (int i = 0; i < 5; i++) FRCS_CUDA_CHECK(cudaStreamWaitEvent(P.streams[i], P.fork, 0));
  // order: H first (largest), then B2 (G2), then the three small G1 MSMs
  if ((rc = msm_run<Fq>(ctx, (uint32_t*)ctx->pk_h.pts, ctx->pk_h.n, (uint32_t*)P.h, ctx->pk_h.n, nullptr, 1,
                        P.msm_work[3], res + 3 * 48, P.streams[3], PROF_MSM_H, PROF_MSM_H_ACCUM)))
    return rc;
  if ((rc = msm_run<Fq2>(ctx, (uint32_t*)ctx->pk_b2.pts, ctx->pk_b2.n, z32, nv, ex + 16, 1, P.msm_work[4],
                         res + 4 * 48, P.streams[4], PROF_MSM_B2)))
    return rc;
  if ((rc = msm_run<Fq>(ctx, (uint32_t*)ctx->pk_a.pts, ctx->pk_a.n, z32, nv, ex, 1, P.msm_work[0], res,
                        P.streams[0], PROF_MSM_A)))
    return rc;
  if ((rc = msm_run<Fq>(ctx, (uint32_t*)ctx->pk_b1.pts, ctx->pk_b1.n, z32, nv, ex + 16, 1, P.msm_work[1],
                        res + 48, P.streams[1], PROF_MSM_B1)))
    return rc;
  if ((rc = msm_run<Fq>(ctx, (uint32_t*)ctx->pk_l.pts, ctx->pk_l.n, z32 + 8 * n_inst, n_wit, ex + 32, 1,
                        P.msm_work[2], res + 2 * 48, P.streams[2], PROF_MSM_L)))
    return rc;
  (void)n;
  for (int i = 0; i < 5; i++) {
    FRCS_CUDA_CHECK(cudaEventRecord(P.done[i], P.streams[i]));
    FRCS_CUDA_CHECK(cudaStreamWaitEvent(st, P.done[i], 0));
  }
  FRCS_CUDA_CHECK(cudaMemcpyAsync(P.h_results + (size_t)slot * PROOF_MSM_WORDS, res, PROOF_MSM_WORDS * 8,
                                  cudaMemcpyDeviceToHost, st));
  FRCS_CUDA_CHECK(cudaEventRecord(P.copied[slot], st));
  return FRCS_OK;
}

// proves n assignments already on the device; r, s on the device (Montgomery); proofs to host memory
int32_t prove_device_z(frcs_ctx* ctx, uint64_t n, const uint64_t* d_z, const uint64_t* d_r, const uint64_t* d_s,
                       const uint64_t* h_r, const uint64_t* h_s, uint64_t* proofs_host, cudaStream_t st) {
  if (!ctx->has_pk) {
    frcs_set_error("no proving key loaded (frcs_load_pk)");
    return FRCS_E_NO_PK;
  }
  int32_t rc = ensure_prover(ctx);
  if (rc) return rc;
  ProverState& P = ctx->prover;
  for (uint64_t i = 0; i < n; i++) {
    int slot = (int)(i & 1);
    rc = launch_proof(ctx, d_z + i * ctx->L.n_z * 4, (const uint32_t*)(d_r + 4 * i), (const uint32_t*)(d_s + 4 * i),
                      slot, st);
    if (rc) return rc;
    if (i > 0) {  // finish the previous proof on the host while this one runs
      FRCS_CUDA_CHECK(cudaEventSynchronize(P.copied[slot ^ 1]));
      timed_finalize(ctx, P.h_results + (size_t)(slot ^ 1) * PROOF_MSM_WORDS, h_r + 4 * (i - 1), h_s + 4 * (i - 1),
                          proofs_host + 48 * (i - 1));
    }
  }
  if (n > 0) {
    int slot = (int)((n - 1) & 1);
    FRCS_CUDA_CHECK(cudaEventSynchronize(P.copied[slot]));
    timed_finalize(ctx, P.h_results + (size_t)slot * PROOF_MSM_WORDS, h_r + 4 * (n - 1), h_s + 4 * (n - 1),
                        proofs_host + 48 * (n - 1));
  }
  return FRCS_OK;
}

}  // namespace

extern "C" {

int32_t frcs_load_pk(frcs_ctx* ctx, const frcs_pk_view* pk) {
  if (!ctx || !pk || !pk->a_query || !pk->b_g1_query || !pk->b_g2_query || !pk->h_query || !pk->l_query)
    return FRCS_E_INVALID_ARG;
  const uint64_t nv = (uint64_t)ctx->L.n_inst + ctx->L.n_wit, n = 1ull << ctx->domain_log2;
  if (pk->a_len != nv || pk->b_g1_len != nv || pk->b_g2_len != nv || pk->l_len != ctx->L.n_wit || pk->h_len != n - 1) {
    frcs_set_error("frcs_load_pk: query lengths do not match the circuit");
    return FRCS_E_INVALID_ARG;
  }
  FRCS_CUDA_CHECK(cudaSetDevice(ctx->device));
  for (DevBases* b : {&ctx->pk_a, &ctx->pk_b1, &ctx->pk_b2, &ctx->pk_h, &ctx->pk_l}) {
    cudaFree(b->pts);
    b->pts = nullptr;
  }
  ctx->has_pk = false;
  int32_t rc;
  const uint64_t* ea[2] = {pk->alpha_g1, pk->delta_g1};
  const uint64_t* eb1[2] = {pk->beta_g1, pk->delta_g1};
  const uint64_t* eb2[2] = {pk->beta_g2, pk->delta_g2};
  const uint64_t* el[1] = {pk->delta_g1};
  if ((rc = upload_and_precompute_g1(ctx, pk->a_query, nv, ea, 2, &ctx->pk_a))) return rc;
  if ((rc = upload_and_precompute_g1(ctx, pk->b_g1_query, nv, eb1, 2, &ctx->pk_b1))) return rc;
  if ((rc = upload_and_precompute_g2(ctx, pk->b_g2_query, nv, eb2, 2, &ctx->pk_b2))) return rc;
  if ((rc = upload_and_precompute_g1(ctx, pk->l_query, ctx->L.n_wit, el, 1, &ctx->pk_l))) return rc;
  if ((rc = upload_and_precompute_g1(ctx, pk->h_query, n - 1, nullptr, 0, &ctx->pk_h))) return rc;
  ctx->has_pk = true;
  return FRCS_OK;
}

int32_t frcs_prove_from_z(frcs_ctx* ctx, uint64_t n, const uint64_t* z, const uint64_t* r, const uint64_t* s,
                          uint64_t* proofs_out) {
  if (!ctx || !z || !r || !s || !proofs_out) return FRCS_E_INVALID_ARG;
  FRCS_CUDA_CHECK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  uint64_t *d_z = nullptr, *d_rs = nullptr;
  const size_t zb = (size_t)ctx->L.n_z * 32;
  FRCS_CUDA_CHECK(cudaMalloc(&d_z, n * zb));
  FRCS_CUDA_CHECK(cudaMalloc(&d_rs, 2 * n * 32 + 32));
  FRCS_CUDA_CHECK(cudaMemcpyAsync(d_z, z, n * zb, cudaMemcpyHostToDevice, st));
  FRCS_CUDA_CHECK(cudaMemcpyAsync(d_rs, r, n * 32, cudaMemcpyHostToDevice, st));
  FRCS_CUDA_CHECK(cudaMemcpyAsync(d_rs + 4 * n, s, n * 32, cudaMemcpyHostToDevice, st));
  int32_t rc = prove_device_z(ctx, n, d_z, d_rs, d_rs + 4 * n, r, s, proofs_out, st);
  cudaStreamSynchronize(st);
  cudaFree(d_z);
  cudaFree(d_rs);
  return rc;
}

int32_t frcs_prove_batch(frcs_ctx* ctx, uint64_t n, const uint16_t* sig, const uint16_t* pk, const uint16_t* hm,
                         const uint64_t* r, const uint64_t* s, uint64_t* proofs_out, int32_t* status) {
  if (!ctx || !sig || !pk || !hm || !r || !s || !proofs_out || !status) return FRCS_E_INVALID_ARG;
  FRCS_CUDA_CHECK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const size_t in_b = (size_t)ctx->L.n * 2, zb = (size_t)ctx->L.n_z * 32;
  const uint64_t CH = 64;  // assignments resident at a time (64 x 5 MB)
  uint16_t *d_in = nullptr;
  uint64_t *d_z = nullptr, *d_rs = nullptr;
  int32_t* d_st = nullptr;
  const uint64_t ch = n < CH ? n : CH;
  FRCS_CUDA_CHECK(cudaMalloc(&d_in, 3 * ch * in_b + 16));
  FRCS_CUDA_CHECK(cudaMalloc(&d_z, ch * zb + 32));
  FRCS_CUDA_CHECK(cudaMalloc(&d_rs, 2 * ch * 32 + 32));
  FRCS_CUDA_CHECK(cudaMalloc(&d_st, ch * 4 + 4));
  int32_t rc = FRCS_OK;
  for (uint64_t i0 = 0; i0 < n && rc == FRCS_OK; i0 += ch) {
    const uint64_t m = n - i0 < ch ? n - i0 : ch;
    uint16_t *ds = d_in, *dp = d_in + ch * ctx->L.n, *dh = d_in + 2 * ch * ctx->L.n;
    FRCS_CUDA_CHECK(cudaMemcpyAsync(ds, sig + i0 * ctx->L.n, m * in_b, cudaMemcpyHostToDevice, st));
    FRCS_CUDA_CHECK(cudaMemcpyAsync(dp, pk + i0 * ctx->L.n, m * in_b, cudaMemcpyHostToDevice, st));
    FRCS_CUDA_CHECK(cudaMemcpyAsync(dh, hm + i0 * ctx->L.n, m * in_b, cudaMemcpyHostToDevice, st));
    FRCS_CUDA_CHECK(cudaMemcpyAsync(d_rs, r + 4 * i0, m * 32, cudaMemcpyHostToDevice, st));
    FRCS_CUDA_CHECK(cudaMemcpyAsync(d_rs + 4 * ch, s + 4 * i0, m * 32, cudaMemcpyHostToDevice, st));
    rc = launch_witness(ctx, m, ds, dp, dh, d_z, d_st, st);
    if (rc) break;
    FRCS_CUDA_CHECK(cudaMemcpyAsync(status + i0, d_st, m * 4, cudaMemcpyDeviceToHost, st));
    rc = prove_device_z(ctx, m, d_z, d_rs, d_rs + 4 * ch, r + 4 * i0, s + 4 * i0, proofs_out + 48 * i0, st);
  }
  cudaStreamSynchronize(st);
  cudaFree(d_in);
  cudaFree(d_z);
  cudaFree(d_rs);
  cudaFree(d_st);
  return rc;
}

// inputs already in HBM; proofs written to HBM (through the pinned host tail) on `stream`
int32_t frcs_prove_batch_dev(frcs_ctx* ctx, uint64_t n, const uint16_t* d_sig, const uint16_t* d_pk,
                             const uint16_t* d_hm, const uint64_t* d_r, const uint64_t* d_s, uint64_t* d_proofs,
                             int32_t* d_status, void* stream) {
  if (!ctx || !d_sig || !d_pk || !d_hm || !d_r || !d_s || !d_proofs || !d_status) return FRCS_E_INVALID_ARG;
  FRCS_CUDA_CHECK(cudaSetDevice(ctx->device));
  cudaStream_t st = (cudaStream_t)stream;
  const size_t zb = (size_t)ctx->L.n_z * 32;
  const uint64_t CH = 64;
  const uint64_t ch = n < CH ? n : CH;
  uint64_t* d_z = nullptr;
  FRCS_CUDA_CHECK(cudaMalloc(&d_z, ch * zb + 32));
  std::vector<uint64_t> h_rs(8 * n + 8), h_proofs(48 * n + 48);
  FRCS_CUDA_CHECK(cudaMemcpyAsync(h_rs.data(), d_r, n * 32, cudaMemcpyDeviceToHost, st));
  FRCS_CUDA_CHECK(cudaMemcpyAsync(h_rs.data() + 4 * n, d_s, n * 32, cudaMemcpyDeviceToHost, st));
  FRCS_CUDA_CHECK(cudaStreamSynchronize(st));
  int32_t rc = FRCS_OK;
  for (uint64_t i0 = 0; i0 < n && rc == FRCS_OK; i0 += ch) {
    const uint64_t m = n - i0 < ch ? n - i0 : ch;
    rc = launch_witness(ctx, m, d_sig + i0 * ctx->L.n, d_pk + i0 * ctx->L.n, d_hm + i0 * ctx->L.n, d_z,
                        d_status + i0, st);
    if (rc) break;
    rc = prove_device_z(ctx, m, d_z, d_r + 4 * i0, d_s + 4 * i0, h_rs.data() + 4 * i0, h_rs.data() + 4 * n + 4 * i0,
                        h_proofs.data() + 48 * i0, st);
  }
  if (rc == FRCS_OK) {
    FRCS_CUDA_CHECK(cudaMemcpyAsync(d_proofs, h_proofs.data(), n * 384, cudaMemcpyHostToDevice, st));
    FRCS_CUDA_CHECK(cudaStreamSynchronize(st));
  }
  cudaFree(d_z);
  return rc;
}

int32_t frcs_proof_compress(const uint64_t* proof_affine, uint8_t* out192) {
  if (!proof_affine || !out192) return FRCS_E_INVALID_ARG;
  host_compress_proof(proof_affine, out192);
  return FRCS_OK;
}

int32_t frcs_imad_peak(frcs_ctx* ctx, double* lp_per_s) {
  if (!ctx || !lp_per_s) return FRCS_E_INVALID_ARG;
  FRCS_CUDA_CHECK(cudaSetDevice(ctx->device));
  int sms = 0;
  FRCS_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device));
  uint64_t* d_out = nullptr;
  FRCS_CUDA_CHECK(cudaMalloc(&d_out, 8));
  cudaEvent_t e0, e1;
  FRCS_CUDA_CHECK(cudaEventCreate(&e0));
  FRCS_CUDA_CHECK(cudaEventCreate(&e1));
  const uint32_t iters = 1 << 15, blocks = sms * 8;
  cudaStream_t st = ctx->stream;
  double best = 0;
  for (int rep = 0; rep < 4; rep++) {
    FRCS_CUDA_CHECK(cudaEventRecord(e0, st));
    imad_peak_kernel<<<blocks, 256, 0, st>>>(d_out, iters, 12345u + rep);
    FRCS_CUDA_CHECK(cudaEventRecord(e1, st));
    FRCS_CUDA_CHECK(cudaEventSynchronize(e1));
    float ms = 0;
    FRCS_CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
    double lp = (double)blocks * 256 * iters * 8 / (ms * 1e-3);
    if (rep > 0 && lp > best) best = lp;
    ctx->launches++;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d_out);
  *lp_per_s = best;
  return FRCS_OK;
}

}  // extern "C"
