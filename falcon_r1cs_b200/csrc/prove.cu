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
#include <thread>
#include <vector>

#include "ctx.hpp"
#include "nvtx.hpp"
#define FF_INLINE_MUL
#include "ff32.cuh"
#include "finalize.hpp"
#include "msm.hpp"

using ff::Fq;
using ff::Fq2;
using ff::Fr;

namespace {

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
// extra scalars of proof p (Montgomery): ex[p] = {1, r, s, -(r s)}, 32 words
#define EX_WORDS 32
__global__ void extras_kernel(const uint32_t* r, const uint32_t* s, uint32_t* ex) {
  r += 8 * blockIdx.x;
  s += 8 * blockIdx.x;
  ex += EX_WORDS * blockIdx.x;
  Fr rr = ld_fr(r), ss = ld_fr(s);
  st_fr(ex, Fr::one());
  st_fr(ex + 8, rr);
  st_fr(ex + 16, ss);
  st_fr(ex + 24, (rr * ss).neg());
}

// IMAD.WIDE.U32 peak (the 32x32->64 limb product every Montgomery multiplication is made of): 8
// independent accumulate chains per thread, acc_i += hi(acc_{i+1}) * b, written as mad.lo.cc / madc.hi
// pairs that ptxas fuses into one IMAD.WIDE.U32 each.  The multiplicand depends on another chain's
// previous value, so nothing can be hoisted or folded (an earlier version with loop-invariant operands
// was folded into additions and over-stated the peak).  Measured on B200: 9.2e12 products/s = 32 per
// clock per SM, i.e. one warp instruction per 4 cycles per SM sub-partition.
__global__ void __launch_bounds__(256) imad_peak_kernel(uint64_t* out, uint32_t iters, uint32_t seed) {
  uint32_t b = seed * 2654435761u + blockIdx.x * 977u + threadIdx.x;
  uint32_t lo[8], hi[8];
#pragma unroll
  for (int k = 0; k < 8; k++) {
    lo[k] = b + k * 7;
    hi[k] = b ^ (k * 13);
  }
  for (uint32_t i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 8; k++)
      asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;"
                   : "+r"(lo[k]), "+r"(hi[k])
                   : "r"(hi[(k + 1) & 7]), "r"(b));
  }
  uint64_t s = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) s ^= lo[k] ^ ((uint64_t)hi[k] << 32);
  if (s == 0x1234567) out[0] = s;  // keep the chains alive
}

// One segment of a base table: `len` affine points at host pointer `p` (nullptr = points at infinity).
struct BaseSeg {
  const uint64_t* p;
  uint64_t len;
  bool on_device = false;
};
// Concatenates the segments on the device and pre-processes them into the 16-window table.
template <class F>
int32_t upload_and_precompute(frcs_ctx* ctx, const std::vector<BaseSeg>& segs, DevBases* out, int cb) {
  constexpr size_t AB = 2 * sizeof(F);
  uint64_t n = 0;
  for (auto& sg : segs) n += sg.len;
  uint8_t* d_in = nullptr;
  FRCS_CUDA_CHECK(cudaMalloc(&d_in, n * AB));
  uint64_t at = 0;
  for (auto& sg : segs) {
    // everything on ctx->stream (non-blocking): a device-to-device cudaMemcpy on the legacy stream would not be
    // ordered with the pre-processing kernel below
    if (sg.p)
      FRCS_CUDA_CHECK(cudaMemcpyAsync(d_in + at * AB, sg.p, sg.len * AB,
                                      sg.on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, ctx->stream));
    else
      FRCS_CUDA_CHECK(cudaMemsetAsync(d_in + at * AB, 0, sg.len * AB, ctx->stream));
    at += sg.len;
  }
  FRCS_CUDA_CHECK(cudaMalloc(&out->pts, n * AB * msm_windows(cb)));
  out->n = n;
  out->windows = (int)msm_windows(cb);
  out->cb = cb;
  out->g2 = sizeof(F) == sizeof(Fq2);
  int32_t rc = msm_precompute<F>(ctx, (const uint32_t*)d_in, n, (uint32_t*)out->pts, ctx->stream, cb);
  FRCS_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  cudaFree(d_in);
  return rc;
}

// Window geometry of the a / b_g1 / b_g2 tables (msm.hpp): 8-bit digits where the assignment is dominated by bits
// and 14-bit values (the NTT circuits), 16-bit digits for the schoolbook circuit, whose N^2 product witnesses are
// 28-bit values (two 16-bit digits instead of four 8-bit ones).  FRCS_Z_WINDOW_BITS overrides (8 or 16).
int z_window_bits(const frcs_ctx* ctx) {
  if (const char* e = getenv("FRCS_Z_WINDOW_BITS")) {
    const int v = atoi(e);
    if (v == MSM_CB_NARROW || v == MSM_CB_WIDE) return v;
  }
  // a proving key split over several GPUs serves single proofs: there the latency of the 32768-bucket reduction chain
  // (five launches of a few blocks each) outweighs the extra additions of the narrow geometry
  if (ctx->shard.n > 1) return MSM_CB_NARROW;
  return ctx->L.kind == FRCS_KIND_SCHOOLBOOK ? MSM_CB_WIDE : MSM_CB_NARROW;
}

// geometry of the L + H tables: wide.  (The narrow geometry was measured for key shards, FRCS_LH_WINDOW_BITS=8: twice the
// additions and a 10 ms digit sort of the dense h scalars, 40 ms instead of 20 ms per split proof on 2 GPUs.)
int lh_window_bits(const frcs_ctx* ctx) {
  if (const char* e = getenv("FRCS_LH_WINDOW_BITS")) {
    const int v = atoi(e);
    if (v == MSM_CB_NARROW || v == MSM_CB_WIDE) return v;
  }
  (void)ctx;
  return MSM_CB_WIDE;
}

// proofs per group: every kernel of the pipeline is launched once per group with the proof index
// as a grid dimension, so the latency-bound steps (sort plan, bucket reduction, slice trees) are
// amortised over the group.  FRCS_GROUP overrides the default.
uint32_t group_capacity() {
  const char* e = getenv("FRCS_GROUP");
  int g = e ? atoi(e) : 16;
  return (uint32_t)(g < 1 ? 1 : g > 256 ? 256 : g);
}

void free_prover_buffers(ProverState& P) {
  cudaFree(P.ntt_work);
  cudaFree(P.h);
  cudaFree(P.extras);
  cudaFree(P.results);
  if (P.h_results) cudaFreeHost(P.h_results);
  for (int i = 0; i < 6; i++) cudaFree(P.msm_work[i]);
  P.ntt_work = P.h = P.extras = P.results = nullptr;
  P.h_results = nullptr;
  for (int i = 0; i < 6; i++) P.msm_work[i] = nullptr;
  P.cap = 0;
}

// device buffers of the proving pipeline, sized for groups of up to `want` proofs
int32_t ensure_prover(frcs_ctx* ctx, uint32_t want) {
  ProverState& P = ctx->prover;
  const uint64_t n = 1ull << ctx->domain_log2;
  if (!ctx->prover_ready) {
    // Streams: [0] sort of z, then the a / b_g1 accumulation; [1] the b_g2 accumulation; [3] sort of (w, -rs, h): high
    // priority, short or latency-bound launches.  [2] the l+h accumulation: low priority, large grids that fill whatever
    // the SMs have left.  (The l+h sort must not sit on the low-priority stream: it would starve until the whole z
    // chain is through, 8.4 ms instead of 1.2 ms per group, with the big accumulation waiting behind it.)
    int prio_lo = 0, prio_hi = 0;
    FRCS_CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    for (int i = 0; i < 5; i++)
      FRCS_CUDA_CHECK(cudaStreamCreateWithPriority(&P.streams[i], cudaStreamNonBlocking, i == 2 ? prio_lo : prio_hi));
    FRCS_CUDA_CHECK(cudaEventCreateWithFlags(&P.sorted_z, cudaEventDisableTiming));
    FRCS_CUDA_CHECK(cudaEventCreateWithFlags(&P.sorted_zb, cudaEventDisableTiming));
    FRCS_CUDA_CHECK(cudaEventCreateWithFlags(&P.sorted_lh, cudaEventDisableTiming));
    FRCS_CUDA_CHECK(cudaEventCreateWithFlags(&P.z_ready, cudaEventDisableTiming));
    for (int k = 0; k < 2; k++)
      for (int i = 0; i < 3; i++) FRCS_CUDA_CHECK(cudaEventCreateWithFlags(&P.done[k][i], cudaEventDisableTiming));
    FRCS_CUDA_CHECK(cudaEventCreateWithFlags(&P.fork, cudaEventDisableTiming));
    for (int i = 0; i < 2; i++) FRCS_CUDA_CHECK(cudaEventCreateWithFlags(&P.copied[i], cudaEventDisableTiming));
    ctx->prover_ready = true;
  }
  uint32_t cap = group_capacity();
  if (want < cap) cap = want;
  if (cap <= P.cap) return FRCS_OK;
  FRCS_CUDA_CHECK(cudaDeviceSynchronize());
  free_prover_buffers(P);
  FRCS_CUDA_CHECK(cudaMalloc(&P.ntt_work, (size_t)cap * 3 * n * 32));
  FRCS_CUDA_CHECK(cudaMalloc(&P.h, (size_t)cap * n * 32));
  FRCS_CUDA_CHECK(cudaMalloc(&P.extras, (size_t)2 * cap * EX_WORDS * 4));
  FRCS_CUDA_CHECK(cudaMalloc(&P.results, (size_t)2 * cap * PROOF_MSM_WORDS * 8));
  FRCS_CUDA_CHECK(cudaMallocHost(&P.h_results, (size_t)2 * cap * PROOF_MSM_WORDS * 8));
  const uint64_t nz = ctx->pk_a.n, nlh = ctx->pk_lh.n;
  const int cz = ctx->pk_a.cb, clh = ctx->pk_lh.cb;
  const size_t wb[5] = {msm_sort_bytes(nz, cz), 2 * msm_acc_bytes<Fq>(nz, cz), msm_acc_bytes<Fq2>(nz, cz),
                        msm_sort_bytes(nlh, clh), msm_acc_bytes<Fq>(nlh, clh)};
  // the two sort buffers are double-buffered (slot = group parity): the sorts of group k+1 run while the accumulations
  // of group k still read group k's sorted entries
  P.sort_stride[0] = wb[0] * cap;
  P.sort_stride[1] = wb[3] * cap;
  for (int i = 0; i < 5; i++) FRCS_CUDA_CHECK(cudaMalloc(&P.msm_work[i], wb[i] * cap * (i == 0 || i == 3 ? 2 : 1)));
  if (ctx->b_skip) FRCS_CUDA_CHECK(cudaMalloc(&P.msm_work[5], wb[0] * cap * 2));
  FRCS_CUDA_CHECK(cudaMemset(P.results, 0, (size_t)2 * cap * PROOF_MSM_WORDS * 8));  // the unused H slot stays infinity
  FRCS_CUDA_CHECK(cudaDeviceSynchronize());  // the legacy-stream memset is not ordered with the non-blocking streams
  P.cap = cap;
  return FRCS_OK;
}

// Enqueues the compute of a group of g <= cap proofs whose assignments are consecutive on the device: witness map on
// `st`, then the MSM chains on the prover's streams.  On return `st` has waited for both digit sorts, i.e. for every
// reader of z, h and the extra scalars, so the caller may enqueue the next group's witness map (it overlaps this
// group's accumulations and fills the low-occupancy tail of their bucket reductions).  join_group() enqueues the rest.
// launch_group = group_head (extra scalars, the z chain) + witness map + group_tail (the L + H chain); the split-key
// entry points call the halves around the exchange of the coset vectors.
int32_t group_head(frcs_ctx* ctx, uint32_t g, const uint64_t* d_z, const uint32_t* d_r, const uint32_t* d_s, int slot,
                   cudaStream_t st) {
  ProverState& P = ctx->prover;
  const uint64_t zs = 8ull * ctx->L.n_z;  // u32 words between assignments
  int32_t rc;
  uint32_t* ex = (uint32_t*)P.extras + (size_t)slot * P.cap * EX_WORDS;
  extras_kernel<<<g, 1, 0, st>>>(d_r, d_s, ex);
  ctx->launches++;
  // the A / B1 / B2 MSMs need only z and (r, s): their chain starts now and runs under the witness map
  FRCS_CUDA_CHECK(cudaEventRecord(P.z_ready, st));
  FRCS_CUDA_CHECK(cudaStreamWaitEvent(P.streams[0], P.z_ready, 0));
  const uint64_t RS = PROOF_MSM_WORDS * 2;  // u32 words per proof in the result buffer
  uint32_t* res = (uint32_t*)P.results + (size_t)slot * P.cap * RS;
  const uint32_t* z32 = (const uint32_t*)d_z;
  void* sort_z = (uint8_t*)P.msm_work[0] + (size_t)slot * P.sort_stride[0];
  // this slot's sort buffers were last read by the accumulations of the group before the previous one
  FRCS_CUDA_CHECK(cudaStreamWaitEvent(P.streams[3], P.done[slot][2], 0));
  FRCS_CUDA_CHECK(cudaStreamWaitEvent(P.streams[0], P.done[slot][1], 0));
  // (1) A, B1 (G1) and B2 (G2) share the scalars z ++ (1, r, s): one sort, two accumulation chains
  {
    MsmScalars sc{{z32 + 8 * ctx->shard.z_lo, ex, nullptr}, {zs, EX_WORDS, 0}, {ctx->shard.z_n, 3, 0}};
    const int ps = prof_begin(ctx, PROF_SORT_Z, P.streams[0]);
    if ((rc = msm_sort(ctx, ctx->pk_a.n, sc, 1, g, sort_z, P.streams[0], ctx->pk_a.cb))) return rc;
    prof_end(ctx, ps, P.streams[0]);
    FRCS_CUDA_CHECK(cudaEventRecord(P.sorted_z, P.streams[0]));
    const uint32_t* tabs2[1] = {(const uint32_t*)ctx->pk_b2.pts};
    uint32_t* outs2[1] = {res + 4 * 48};
    if (ctx->b_skip) {
      // B1 and B2 over their own sort of z (without the scalars of the infinity bases), made on the G2 stream; A alone
      // over the full sort.  This slot's B sort was last read by the b_g1 accumulation (stream 0) two groups ago.
      void* sort_zb = (uint8_t*)P.msm_work[5] + (size_t)slot * P.sort_stride[0];
      FRCS_CUDA_CHECK(cudaStreamWaitEvent(P.streams[1], P.z_ready, 0));
      FRCS_CUDA_CHECK(cudaStreamWaitEvent(P.streams[1], P.done[slot][0], 0));
      if ((rc = msm_sort(ctx, ctx->pk_b1.n, sc, 1, g, sort_zb, P.streams[1], ctx->pk_a.cb, ctx->b_skip))) return rc;
      FRCS_CUDA_CHECK(cudaEventRecord(P.sorted_zb, P.streams[1]));
      if ((rc = msm_accumulate<Fq2>(ctx, 1, tabs2, ctx->pk_b2.n, g, sort_zb, P.msm_work[2], outs2, RS, P.streams[1],
                                    ctx->pk_a.cb, PROF_MSM_B2, -1)))
        return rc;
      const uint32_t* tab_a[1] = {(const uint32_t*)ctx->pk_a.pts};
      uint32_t* out_a[1] = {res};
      if ((rc = msm_accumulate<Fq>(ctx, 1, tab_a, ctx->pk_a.n, g, sort_z, P.msm_work[1], out_a, RS, P.streams[0],
                                   ctx->pk_a.cb, PROF_MSM_A, -1)))
        return rc;
      FRCS_CUDA_CHECK(cudaStreamWaitEvent(P.streams[0], P.sorted_zb, 0));
      const uint32_t* tab_b[1] = {(const uint32_t*)ctx->pk_b1.pts};
      uint32_t* out_b[1] = {res + 48};
      void* acc_b = (uint8_t*)P.msm_work[1] + (size_t)P.cap * msm_acc_bytes<Fq>(ctx->pk_a.n, ctx->pk_a.cb);
      if ((rc = msm_accumulate<Fq>(ctx, 1, tab_b, ctx->pk_b1.n, g, sort_zb, acc_b, out_b, RS, P.streams[0], ctx->pk_a.cb,
                                   PROF_MSM_B1, -1)))
        return rc;
      return FRCS_OK;
    }
    FRCS_CUDA_CHECK(cudaStreamWaitEvent(P.streams[1], P.sorted_z, 0));
    if ((rc = msm_accumulate<Fq2>(ctx, 1, tabs2, ctx->pk_b2.n, g, sort_z, P.msm_work[2], outs2, RS, P.streams[1],
                                  ctx->pk_a.cb, PROF_MSM_B2, -1)))
      return rc;
    const uint32_t* tabs[2] = {(const uint32_t*)ctx->pk_a.pts, (const uint32_t*)ctx->pk_b1.pts};
    uint32_t* outs[2] = {res, res + 48};
    if ((rc = msm_accumulate<Fq>(ctx, 2, tabs, ctx->pk_a.n, g, sort_z, P.msm_work[1], outs, RS, P.streams[0],
                                 ctx->pk_a.cb, PROF_MSM_A, -1)))
      return rc;
  }
  return FRCS_OK;
}

int32_t group_tail(frcs_ctx* ctx, uint32_t g, const uint64_t* d_z, int slot, cudaStream_t st) {
  ProverState& P = ctx->prover;
  const uint64_t n_inst = ctx->L.n_inst;
  const uint64_t n = 1ull << ctx->domain_log2;
  const uint64_t zs = 8ull * ctx->L.n_z;
  const uint64_t RS = PROOF_MSM_WORDS * 2;
  uint32_t* res = (uint32_t*)P.results + (size_t)slot * P.cap * RS;
  const uint32_t* z32 = (const uint32_t*)d_z;
  const uint32_t* ex = (const uint32_t*)P.extras + (size_t)slot * P.cap * EX_WORDS;
  void* sort_lh = (uint8_t*)P.msm_work[3] + (size_t)slot * P.sort_stride[1];
  int32_t rc;
  FRCS_CUDA_CHECK(cudaEventRecord(P.fork, st));
  FRCS_CUDA_CHECK(cudaStreamWaitEvent(P.streams[3], P.fork, 0));
  // (2) L + H: one MSM over l_query ++ delta_1 ++ h_query with scalars w ++ (-rs) ++ h
  {
    const frcs_ctx::Shard& sh = ctx->shard;
    MsmScalars sc{{z32 + 8 * (n_inst + sh.l_lo), ex + 24, (const uint32_t*)P.h + 8 * sh.h_lo},
                  {zs, EX_WORDS, 8 * n},
                  {sh.l_n, 1, sh.h_n}};
    const int ps = prof_begin(ctx, PROF_SORT_LH, P.streams[3]);
    if ((rc = msm_sort(ctx, ctx->pk_lh.n, sc, 1, g, sort_lh, P.streams[3], ctx->pk_lh.cb))) return rc;
    prof_end(ctx, ps, P.streams[3]);
    FRCS_CUDA_CHECK(cudaEventRecord(P.sorted_lh, P.streams[3]));
    FRCS_CUDA_CHECK(cudaStreamWaitEvent(P.streams[2], P.sorted_lh, 0));
    const uint32_t* tabs[1] = {(const uint32_t*)ctx->pk_lh.pts};
    uint32_t* outs[1] = {res + 2 * 48};
    if ((rc = msm_accumulate<Fq>(ctx, 1, tabs, ctx->pk_lh.n, g, sort_lh, P.msm_work[4], outs, RS, P.streams[2],
                                 ctx->pk_lh.cb, PROF_MSM_H, PROF_MSM_H_ACCUM)))
      return rc;
  }
  for (int i = 0; i < 3; i++) FRCS_CUDA_CHECK(cudaEventRecord(P.done[slot][i], P.streams[i]));
  FRCS_CUDA_CHECK(cudaStreamWaitEvent(st, P.sorted_lh, 0));
  FRCS_CUDA_CHECK(cudaStreamWaitEvent(st, P.sorted_z, 0));
  if (ctx->b_skip) FRCS_CUDA_CHECK(cudaStreamWaitEvent(st, P.sorted_zb, 0));
  return FRCS_OK;
}

int32_t launch_group(frcs_ctx* ctx, uint32_t g, const uint64_t* d_z, const uint32_t* d_r, const uint32_t* d_s, int slot,
                     cudaStream_t st) {
  NvtxRange nvtx("frcs:proof_group");
  ProverState& P = ctx->prover;
  int32_t rc;
  if ((rc = group_head(ctx, g, d_z, d_r, d_s, slot, st))) return rc;
  if ((rc = launch_witness_map(ctx, g, d_z, (uint64_t*)P.h, (uint32_t*)P.ntt_work, st))) return rc;
  return group_tail(ctx, g, d_z, slot, st);
}

// Enqueues the join of a launched group on `st` and the copy of its g x PROOF_MSM_WORDS MSM sums into pinned host slot
// `slot` (event copied[slot]).
int32_t join_group(frcs_ctx* ctx, uint32_t g, int slot, cudaStream_t st) {
  ProverState& P = ctx->prover;
  for (int i = 0; i < 3; i++) FRCS_CUDA_CHECK(cudaStreamWaitEvent(st, P.done[slot][i], 0));
  const uint64_t* res = (const uint64_t*)P.results + (size_t)slot * P.cap * PROOF_MSM_WORDS;
  FRCS_CUDA_CHECK(cudaMemcpyAsync(P.h_results + (size_t)slot * P.cap * PROOF_MSM_WORDS, res,
                                  (size_t)g * PROOF_MSM_WORDS * 8, cudaMemcpyDeviceToHost, st));
  FRCS_CUDA_CHECK(cudaEventRecord(P.copied[slot], st));
  return FRCS_OK;
}

// host tail of a finished group, spread over a few threads (each proof is ~0.5 ms of one core)
void finalize_group(frcs_ctx* ctx, uint32_t g, const uint64_t* msm, const uint64_t* h_r, const uint64_t* h_s,
                    uint64_t* proofs) {
  NvtxRange nvtx("frcs:host_tail");
  auto t0 = std::chrono::steady_clock::now();
  unsigned hw = std::thread::hardware_concurrency();
  uint32_t nt = hw ? hw : 4;
  if (nt > 16) nt = 16;
  if (nt > g) nt = g;
  auto work = [&](uint32_t t) {
    for (uint32_t i = t; i < g; i += nt)
      host_finalize_proof(msm + (size_t)i * PROOF_MSM_WORDS, h_r + 4 * i, h_s + 4 * i, proofs + 48 * i);
  };
  if (nt <= 1) {
    work(0);
  } else {
    std::vector<std::thread> th;
    for (uint32_t t = 1; t < nt; t++) th.emplace_back(work, t);
    work(0);
    for (auto& x : th) x.join();
  }
  if (ctx->prof.on) {
    ctx->prof.ms[PROF_HOST_TAIL] += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    ctx->prof.count[PROF_HOST_TAIL]++;
  }
}

// proves n assignments already on the device; r, s on the device (Montgomery); proofs to host memory
// partials_dev != nullptr: instead of finishing the proofs, the raw MSM sums (PROOF_MSM_WORDS u64 per proof) are
// copied there (device memory): the caller combines the shards of a split proving key.
// Pipeline over the groups: launch(k) is enqueued before join(k-1), so the witness map of group k runs while the
// accumulations of group k-1 are still in flight; the host tail of group k-2 runs meanwhile on the CPU.
int32_t prove_device_z(frcs_ctx* ctx, uint64_t n, const uint64_t* d_z, const uint64_t* d_r, const uint64_t* d_s,
                       const uint64_t* h_r, const uint64_t* h_s, uint64_t* proofs_host, cudaStream_t st,
                       uint64_t* partials_dev = nullptr) {
  if (!ctx->has_pk) {
    frcs_set_error("no proving key loaded (frcs_load_pk)");
    return FRCS_E_NO_PK;
  }
  if (n == 0) return FRCS_OK;
  int32_t rc = ensure_prover(ctx, (uint32_t)(n > 256 ? 256 : n));
  if (rc) return rc;
  ProverState& P = ctx->prover;
  const uint64_t cap = P.cap;
  auto size_of = [&](uint64_t i0) { return (uint32_t)(n - i0 < cap ? n - i0 : cap); };
  auto finalize = [&](uint64_t i0, int slot) -> int32_t {
    FRCS_CUDA_CHECK(cudaEventSynchronize(P.copied[slot]));
    finalize_group(ctx, size_of(i0), P.h_results + (size_t)slot * cap * PROOF_MSM_WORDS, h_r + 4 * i0, h_s + 4 * i0,
                   proofs_host + 48 * i0);
    return FRCS_OK;
  };
  int k = 0;
  for (uint64_t i0 = 0; i0 < n; i0 += cap, k++) {
    const uint32_t g = size_of(i0);
    const int slot = k & 1;
    const int pg = prof_begin(ctx, PROF_GROUP, st);
    rc = launch_group(ctx, g, d_z + i0 * ctx->L.n_z * 4, (const uint32_t*)(d_r + 4 * i0), (const uint32_t*)(d_s + 4 * i0),
                      slot, st);
    prof_end(ctx, pg, st);
    if (rc) return rc;
    if (partials_dev) {
      for (int i = 0; i < 3; i++) FRCS_CUDA_CHECK(cudaStreamWaitEvent(st, P.done[slot][i], 0));
      FRCS_CUDA_CHECK(cudaMemcpyAsync(partials_dev + i0 * PROOF_MSM_WORDS,
                                      (const uint64_t*)P.results + (size_t)slot * cap * PROOF_MSM_WORDS,
                                      (size_t)g * PROOF_MSM_WORDS * 8, cudaMemcpyDeviceToDevice, st));
      FRCS_CUDA_CHECK(cudaStreamSynchronize(st));  // the slot buffers are reused by the next group
      continue;
    }
    if (k >= 1 && (rc = join_group(ctx, size_of(i0 - cap), slot ^ 1, st))) return rc;
    if (k >= 2 && (rc = finalize(i0 - 2 * cap, slot))) return rc;  // group k-2 used this slot; its copy is long done
  }
  if (partials_dev) return FRCS_OK;
  // drain: join the last group, finish the last two on the host
  const uint64_t last0 = (uint64_t)(k - 1) * cap;
  if ((rc = join_group(ctx, size_of(last0), (k - 1) & 1, st))) return rc;
  if (k >= 2 && (rc = finalize(last0 - cap, (k - 2) & 1))) return rc;
  return finalize(last0, (k - 1) & 1);
}

// bit i of mask: base i is the point at infinity in BOTH tables (affine (0, 0); the first n entries of a table are the
// bases themselves); *count = number of bits set
__global__ void __launch_bounds__(256)
    b_skip_kernel(const uint32_t* __restrict__ b1, const uint32_t* __restrict__ b2, uint64_t n, uint32_t* __restrict__ mask,
                  unsigned long long* count) {
  const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  bool inf = false;
  if (i < n) {
    uint32_t any = 0;
    for (int k = 0; k < 24; k++) any |= b1[i * 24 + k];
    for (int k = 0; k < 48; k++) any |= b2[i * 48 + k];
    inf = any == 0;
  }
  const unsigned m = __ballot_sync(0xffffffffu, inf);
  if ((threadIdx.x & 31) == 0 && i < n) {
    mask[i >> 5] = m;
    if (m) atomicAdd(count, (unsigned long long)__popc(m));
  }
}

// after the tables of a key are installed: the B MSMs get their own digit sort when enough of their bases are infinity
int32_t build_b_skip(frcs_ctx* ctx) {
  cudaFree(ctx->b_skip);
  ctx->b_skip = nullptr;
  ctx->b_skip_count = 0;
  static const bool off = getenv("FRCS_NO_BSKIP") != nullptr;
  const uint64_t n = ctx->pk_b1.n;
  if (off || n == 0 || ctx->pk_b2.n != n || ctx->pk_a.n != n) return FRCS_OK;
  uint32_t* mask = nullptr;
  unsigned long long* d_count = nullptr;
  const size_t words = (n + 31) / 32;
  FRCS_CUDA_CHECK(cudaMalloc(&mask, words * 4));
  if (cudaMalloc(&d_count, 8) != cudaSuccess) {
    cudaFree(mask);
    frcs_set_error("build_b_skip: out of device memory");
    return FRCS_E_CUDA;
  }
  cudaMemsetAsync(d_count, 0, 8, ctx->stream);
  b_skip_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>((const uint32_t*)ctx->pk_b1.pts,
                                                                     (const uint32_t*)ctx->pk_b2.pts, n, mask, d_count);
  ctx->launches++;
  unsigned long long h_count = 0;
  cudaError_t e = cudaMemcpyAsync(&h_count, d_count, 8, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(d_count);
  if (e != cudaSuccess) {
    cudaFree(mask);
    frcs_set_error(std::string("build_b_skip: ") + cudaGetErrorString(e));
    return FRCS_E_CUDA;
  }
  ctx->b_skip_count = h_count;
  if (h_count * 8 >= n)  // at least one base in eight
    ctx->b_skip = mask;
  else
    cudaFree(mask);
  return FRCS_OK;
}

}  // namespace

extern "C" {

int32_t frcs_load_pk_shard(frcs_ctx* ctx, const frcs_pk_view* pk, uint32_t shard, uint32_t n_shards) {
  if (!ctx || !pk || !pk->a_query || !pk->b_g1_query || !pk->b_g2_query || !pk->h_query || !pk->l_query)
    return FRCS_E_INVALID_ARG;
  if (n_shards == 0 || shard >= n_shards) {
    frcs_set_error("frcs_load_pk_shard: shard index out of range");
    return FRCS_E_INVALID_ARG;
  }
  const uint64_t nv = (uint64_t)ctx->L.n_inst + ctx->L.n_wit, n = 1ull << ctx->domain_log2, nw = ctx->L.n_wit;
  if (pk->a_len != nv || pk->b_g1_len != nv || pk->b_g2_len != nv || pk->l_len != nw || pk->h_len != n - 1) {
    frcs_set_error("frcs_load_pk: query lengths do not match the circuit");
    return FRCS_E_INVALID_ARG;
  }
  FRCS_CUDA_CHECK(cudaSetDevice(ctx->device));
  FRCS_CUDA_CHECK(cudaDeviceSynchronize());
  for (DevBases* b : {&ctx->pk_a, &ctx->pk_b1, &ctx->pk_b2, &ctx->pk_lh}) {
    cudaFree(b->pts);
    b->pts = nullptr;
  }
  free_prover_buffers(ctx->prover);  // sized for the previous key
  ctx->has_pk = false;
  // contiguous base ranges of this shard
  auto lo = [&](uint64_t len) { return len * shard / n_shards; };
  auto hi = [&](uint64_t len) { return len * (shard + 1) / n_shards; };
  frcs_ctx::Shard& sh = ctx->shard;
  sh.idx = shard;
  sh.n = n_shards;
  sh.z_lo = lo(nv);
  sh.z_n = hi(nv) - sh.z_lo;
  sh.l_lo = lo(nw);
  sh.l_n = hi(nw) - sh.l_lo;
  sh.h_lo = lo(n - 1);
  sh.h_n = hi(n - 1) - sh.h_lo;
  const bool first = shard == 0;  // the constant bases (alpha, beta, delta) live on shard 0; elsewhere infinity
  const int zcb = z_window_bits(ctx);
  int32_t rc;
  // z-tables: query ++ (base of scalar 1, base of scalar r, base of scalar s)   (calculate_coeff, prover.rs)
  if ((rc = upload_and_precompute<Fq>(ctx, {{pk->a_query + 12 * sh.z_lo, sh.z_n}, {first ? pk->alpha_g1 : nullptr, 1},
                                            {first ? pk->delta_g1 : nullptr, 1}, {nullptr, 1}}, &ctx->pk_a, zcb)))
    return rc;
  if ((rc = upload_and_precompute<Fq>(ctx, {{pk->b_g1_query + 12 * sh.z_lo, sh.z_n}, {first ? pk->beta_g1 : nullptr, 1},
                                            {nullptr, 1}, {first ? pk->delta_g1 : nullptr, 1}}, &ctx->pk_b1, zcb)))
    return rc;
  if ((rc = upload_and_precompute<Fq2>(ctx, {{pk->b_g2_query + 24 * sh.z_lo, sh.z_n}, {first ? pk->beta_g2 : nullptr, 1},
                                             {nullptr, 1}, {first ? pk->delta_g2 : nullptr, 1}}, &ctx->pk_b2, zcb)))
    return rc;
  // l_query ++ delta_1 (scalar -rs) ++ h_query: L and H only ever appear as L + H in C
  if ((rc = upload_and_precompute<Fq>(ctx, {{pk->l_query + 12 * sh.l_lo, sh.l_n}, {first ? pk->delta_g1 : nullptr, 1},
                                            {pk->h_query + 12 * sh.h_lo, sh.h_n}}, &ctx->pk_lh, lh_window_bits(ctx))))
    return rc;
  if ((rc = build_b_skip(ctx))) return rc;
  ctx->has_pk = true;
  return FRCS_OK;
}

// the queries are already on the device (frcs_setup): affine G1 (24 words) / G2 (48 words) arrays; the context keeps
// base-range shard `shard` of `n_shards` (see frcs_load_pk_shard)
int32_t install_pk_from_device(frcs_ctx* ctx, const uint32_t* d_a, const uint32_t* d_b1, const uint32_t* d_b2,
                               const uint32_t* d_h, const uint32_t* d_l, const uint32_t* d_c1, const uint32_t* d_c2,
                               uint32_t shard, uint32_t n_shards) {
  const uint64_t nv = (uint64_t)ctx->L.n_inst + ctx->L.n_wit, n = 1ull << ctx->domain_log2, nw = ctx->L.n_wit;
  if (n_shards == 0 || shard >= n_shards) return FRCS_E_INVALID_ARG;
  FRCS_CUDA_CHECK(cudaDeviceSynchronize());
  for (DevBases* b : {&ctx->pk_a, &ctx->pk_b1, &ctx->pk_b2, &ctx->pk_lh}) {
    cudaFree(b->pts);
    b->pts = nullptr;
  }
  free_prover_buffers(ctx->prover);
  ctx->has_pk = false;
  auto lo = [&](uint64_t len) { return len * shard / n_shards; };
  auto hi = [&](uint64_t len) { return len * (shard + 1) / n_shards; };
  frcs_ctx::Shard& sh = ctx->shard;
  sh = frcs_ctx::Shard();
  sh.idx = shard;
  sh.n = n_shards;
  sh.z_lo = lo(nv);
  sh.z_n = hi(nv) - sh.z_lo;
  sh.l_lo = lo(nw);
  sh.l_n = hi(nw) - sh.l_lo;
  sh.h_lo = lo(n - 1);
  sh.h_n = hi(n - 1) - sh.h_lo;
  const bool first = shard == 0;  // the constant bases live on shard 0; elsewhere infinity
  auto dv = [](const uint32_t* p, uint64_t len) { return BaseSeg{(const uint64_t*)p, len, true}; };
  auto cst = [&](const uint32_t* p) { return first ? BaseSeg{(const uint64_t*)p, 1, true} : BaseSeg{nullptr, 1}; };
  const uint32_t *alpha = d_c1, *beta1 = d_c1 + 24, *delta1 = d_c1 + 48, *beta2 = d_c2, *delta2 = d_c2 + 48;
  int32_t rc;
  const int zcb = z_window_bits(ctx);
  if ((rc = upload_and_precompute<Fq>(ctx, {dv(d_a + 24 * sh.z_lo, sh.z_n), cst(alpha), cst(delta1), {nullptr, 1}}, &ctx->pk_a, zcb))) return rc;
  if ((rc = upload_and_precompute<Fq>(ctx, {dv(d_b1 + 24 * sh.z_lo, sh.z_n), cst(beta1), {nullptr, 1}, cst(delta1)}, &ctx->pk_b1, zcb))) return rc;
  if ((rc = upload_and_precompute<Fq2>(ctx, {dv(d_b2 + 48 * sh.z_lo, sh.z_n), cst(beta2), {nullptr, 1}, cst(delta2)}, &ctx->pk_b2, zcb))) return rc;
  if ((rc = upload_and_precompute<Fq>(ctx, {dv(d_l + 24 * sh.l_lo, sh.l_n), cst(delta1), dv(d_h + 24 * sh.h_lo, sh.h_n)}, &ctx->pk_lh, lh_window_bits(ctx)))) return rc;
  if ((rc = build_b_skip(ctx))) return rc;
  ctx->has_pk = true;
  return FRCS_OK;
}

// the proving-key queries held by the context, back to the host (parity tests of frcs_setup):
// which: 0 a_query, 1 b_g1_query, 2 b_g2_query, 3 h_query, 4 l_query
extern "C" int32_t frcs_export_pk(frcs_ctx* ctx, int32_t which, uint64_t* out) {
  if (!ctx || !out || which < 0 || (which > 4 && (which < 10 || which > 12))) return FRCS_E_INVALID_ARG;
  if (!ctx->has_pk || ctx->shard.n != 1) {
    frcs_set_error("frcs_export_pk: no complete proving key in this context");
    return FRCS_E_NO_PK;
  }
  FRCS_CUDA_CHECK(cudaSetDevice(ctx->device));
  const uint64_t nv = (uint64_t)ctx->L.n_inst + ctx->L.n_wit, n = 1ull << ctx->domain_log2, nw = ctx->L.n_wit;
  // window 0 of every table holds the bases themselves
  switch (which) {
    case 0: FRCS_CUDA_CHECK(cudaMemcpy(out, ctx->pk_a.pts, nv * 96, cudaMemcpyDeviceToHost)); break;
    case 1: FRCS_CUDA_CHECK(cudaMemcpy(out, ctx->pk_b1.pts, nv * 96, cudaMemcpyDeviceToHost)); break;
    case 2: FRCS_CUDA_CHECK(cudaMemcpy(out, ctx->pk_b2.pts, nv * 192, cudaMemcpyDeviceToHost)); break;
    case 3: FRCS_CUDA_CHECK(cudaMemcpy(out, (uint8_t*)ctx->pk_lh.pts + (nw + 1) * 96, (n - 1) * 96, cudaMemcpyDeviceToHost)); break;
    case 4: FRCS_CUDA_CHECK(cudaMemcpy(out, ctx->pk_lh.pts, nw * 96, cudaMemcpyDeviceToHost)); break;
    // debug: the three constant bases appended to the a / b_g1 / b_g2 tables
    case 10: FRCS_CUDA_CHECK(cudaMemcpy(out, (uint8_t*)ctx->pk_a.pts + nv * 96, 3 * 96, cudaMemcpyDeviceToHost)); break;
    case 11: FRCS_CUDA_CHECK(cudaMemcpy(out, (uint8_t*)ctx->pk_b1.pts + nv * 96, 3 * 96, cudaMemcpyDeviceToHost)); break;
    case 12: FRCS_CUDA_CHECK(cudaMemcpy(out, (uint8_t*)ctx->pk_b2.pts + nv * 192, 3 * 192, cudaMemcpyDeviceToHost)); break;
  }
  return FRCS_OK;
}

int32_t frcs_load_pk(frcs_ctx* ctx, const frcs_pk_view* pk) { return frcs_load_pk_shard(ctx, pk, 0, 1); }

int32_t frcs_combine_partials(uint32_t n_shards, uint64_t n, const uint64_t* partials, const uint64_t* r,
                              const uint64_t* s, uint64_t* proofs_out) {
  if (!n_shards || !partials || !r || !s || !proofs_out) return FRCS_E_INVALID_ARG;
  std::vector<const uint64_t*> ps(n_shards);
  for (uint64_t i = 0; i < n; i++) {
    for (uint32_t sh = 0; sh < n_shards; sh++) ps[sh] = partials + ((size_t)sh * n + i) * PROOF_MSM_WORDS;
    uint64_t sum[PROOF_MSM_WORDS];
    host_sum_partials(n_shards, ps.data(), sum);
    host_finalize_proof(sum, r + 4 * i, s + 4 * i, proofs_out + 48 * i);
  }
  return FRCS_OK;
}

// staging buffer of the host entry points (persistent: no cudaMalloc/cudaFree per call)
static int32_t ensure_io(frcs_ctx* ctx, size_t bytes, uint8_t** out) {
  ProverState& P = ctx->prover;
  if (P.io_bytes < bytes) {
    FRCS_CUDA_CHECK(cudaDeviceSynchronize());
    cudaFree(P.io);
    P.io = nullptr;
    P.io_bytes = 0;
    FRCS_CUDA_CHECK(cudaMalloc(&P.io, bytes));
    P.io_bytes = bytes;
  }
  *out = (uint8_t*)P.io;
  return FRCS_OK;
}
static size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

int32_t frcs_prove_from_z(frcs_ctx* ctx, uint64_t n, const uint64_t* z, const uint64_t* r, const uint64_t* s,
                          uint64_t* proofs_out) {
  if (!ctx || !z || !r || !s || !proofs_out) return FRCS_E_INVALID_ARG;
  FRCS_CUDA_CHECK(cudaSetDevice(ctx->device));
  if (!ctx->has_pk) {
    frcs_set_error("no proving key loaded (frcs_load_pk)");
    return FRCS_E_NO_PK;
  }
  cudaStream_t st = ctx->stream;
  const size_t zb = (size_t)ctx->L.n_z * 32;
  const uint64_t CH = 64;
  const uint64_t ch = n < CH ? n : CH;
  uint8_t* io;
  int32_t rc = ensure_io(ctx, al256(ch * zb) + 2 * al256(ch * 32), &io);
  if (rc) return rc;
  uint64_t* d_z = (uint64_t*)io;
  uint64_t* d_r = (uint64_t*)(io + al256(ch * zb));
  uint64_t* d_s = (uint64_t*)(io + al256(ch * zb) + al256(ch * 32));
  for (uint64_t i0 = 0; i0 < n && rc == FRCS_OK; i0 += ch) {
    const uint64_t m = n - i0 < ch ? n - i0 : ch;
    FRCS_CUDA_CHECK(cudaMemcpyAsync(d_z, z + i0 * ctx->L.n_z * 4, m * zb, cudaMemcpyHostToDevice, st));
    FRCS_CUDA_CHECK(cudaMemcpyAsync(d_r, r + 4 * i0, m * 32, cudaMemcpyHostToDevice, st));
    FRCS_CUDA_CHECK(cudaMemcpyAsync(d_s, s + 4 * i0, m * 32, cudaMemcpyHostToDevice, st));
    rc = prove_device_z(ctx, m, d_z, d_r, d_s, r + 4 * i0, s + 4 * i0, proofs_out + 48 * i0, st);
  }
  cudaStreamSynchronize(st);
  return rc;
}

int32_t frcs_prove_batch(frcs_ctx* ctx, uint64_t n, const uint16_t* sig, const uint16_t* pk, const uint16_t* hm,
                         const uint64_t* r, const uint64_t* s, uint64_t* proofs_out, int32_t* status) {
  if (!ctx || !sig || !pk || !hm || !r || !s || !proofs_out || !status) return FRCS_E_INVALID_ARG;
  FRCS_CUDA_CHECK(cudaSetDevice(ctx->device));
  if (!ctx->has_pk) {
    frcs_set_error("no proving key loaded (frcs_load_pk)");
    return FRCS_E_NO_PK;
  }
  cudaStream_t st = ctx->stream;
  const size_t in_b = (size_t)ctx->L.n * 2, zb = (size_t)ctx->L.n_z * 32;
  const uint64_t CH = 64;  // assignments resident at a time (64 x 5 MB)
  const uint64_t ch = n < CH ? n : CH;
  uint8_t* io;
  int32_t rc = ensure_io(ctx, al256(ch * zb) + 3 * al256(ch * in_b) + 2 * al256(ch * 32) + al256(ch * 4), &io);
  if (rc) return rc;
  uint64_t* d_z = (uint64_t*)io;
  io += al256(ch * zb);
  uint16_t *ds = (uint16_t*)io, *dp = (uint16_t*)(io + al256(ch * in_b)), *dh = (uint16_t*)(io + 2 * al256(ch * in_b));
  io += 3 * al256(ch * in_b);
  uint64_t *d_r = (uint64_t*)io, *d_s = (uint64_t*)(io + al256(ch * 32));
  int32_t* d_st = (int32_t*)(io + 2 * al256(ch * 32));
  for (uint64_t i0 = 0; i0 < n && rc == FRCS_OK; i0 += ch) {
    const uint64_t m = n - i0 < ch ? n - i0 : ch;
    FRCS_CUDA_CHECK(cudaMemcpyAsync(ds, sig + i0 * ctx->L.n, m * in_b, cudaMemcpyHostToDevice, st));
    FRCS_CUDA_CHECK(cudaMemcpyAsync(dp, pk + i0 * ctx->L.n, m * in_b, cudaMemcpyHostToDevice, st));
    FRCS_CUDA_CHECK(cudaMemcpyAsync(dh, hm + i0 * ctx->L.n, m * in_b, cudaMemcpyHostToDevice, st));
    FRCS_CUDA_CHECK(cudaMemcpyAsync(d_r, r + 4 * i0, m * 32, cudaMemcpyHostToDevice, st));
    FRCS_CUDA_CHECK(cudaMemcpyAsync(d_s, s + 4 * i0, m * 32, cudaMemcpyHostToDevice, st));
    rc = launch_witness(ctx, m, ds, dp, dh, d_z, d_st, st);
    if (rc) break;
    FRCS_CUDA_CHECK(cudaMemcpyAsync(status + i0, d_st, m * 4, cudaMemcpyDeviceToHost, st));
    rc = prove_device_z(ctx, m, d_z, d_r, d_s, r + 4 * i0, s + 4 * i0, proofs_out + 48 * i0, st);
  }
  cudaStreamSynchronize(st);
  return rc;
}

// inputs already in HBM; proofs written to HBM (through the pinned host tail) on `stream`
int32_t frcs_prove_batch_dev(frcs_ctx* ctx, uint64_t n, const uint16_t* d_sig, const uint16_t* d_pk,
                             const uint16_t* d_hm, const uint64_t* d_r, const uint64_t* d_s, uint64_t* d_proofs,
                             int32_t* d_status, void* stream) {
  if (!ctx || !d_sig || !d_pk || !d_hm || !d_r || !d_s || !d_proofs || !d_status) return FRCS_E_INVALID_ARG;
  FRCS_CUDA_CHECK(cudaSetDevice(ctx->device));
  if (!ctx->has_pk) {
    frcs_set_error("no proving key loaded (frcs_load_pk)");
    return FRCS_E_NO_PK;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const size_t zb = (size_t)ctx->L.n_z * 32;
  const uint64_t CH = 64;
  const uint64_t ch = n < CH ? n : CH;
  uint8_t* io;
  int32_t rc = ensure_io(ctx, al256(ch * zb), &io);
  if (rc) return rc;
  uint64_t* d_z = (uint64_t*)io;
  std::vector<uint64_t> h_rs(8 * n + 8), h_proofs(48 * n + 48);
  FRCS_CUDA_CHECK(cudaMemcpyAsync(h_rs.data(), d_r, n * 32, cudaMemcpyDeviceToHost, st));
  FRCS_CUDA_CHECK(cudaMemcpyAsync(h_rs.data() + 4 * n, d_s, n * 32, cudaMemcpyDeviceToHost, st));
  FRCS_CUDA_CHECK(cudaStreamSynchronize(st));
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
  return rc;
}

// Sharded proving key (frcs_load_pk_shard): the raw MSM sums of this shard's base ranges, PROOF_MSM_WORDS u64 per
// proof, to device memory; the shards' sums are gathered by the caller (NCCL) and finished by frcs_combine_partials.
int32_t frcs_prove_partial_dev(frcs_ctx* ctx, uint64_t n, const uint16_t* d_sig, const uint16_t* d_pk,
                               const uint16_t* d_hm, const uint64_t* d_r, const uint64_t* d_s, uint64_t* d_partials,
                               int32_t* d_status, void* stream) {
  if (!ctx || !d_sig || !d_pk || !d_hm || !d_r || !d_s || !d_partials || !d_status) return FRCS_E_INVALID_ARG;
  FRCS_CUDA_CHECK(cudaSetDevice(ctx->device));
  if (!ctx->has_pk) {
    frcs_set_error("no proving key loaded (frcs_load_pk)");
    return FRCS_E_NO_PK;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const size_t zb = (size_t)ctx->L.n_z * 32;
  const uint64_t CH = 64;
  const uint64_t ch = n < CH ? n : CH;
  uint8_t* io;
  int32_t rc = ensure_io(ctx, al256(ch * zb), &io);
  if (rc) return rc;
  uint64_t* d_z = (uint64_t*)io;
  for (uint64_t i0 = 0; i0 < n && rc == FRCS_OK; i0 += ch) {
    const uint64_t m = n - i0 < ch ? n - i0 : ch;
    rc = launch_witness(ctx, m, d_sig + i0 * ctx->L.n, d_pk + i0 * ctx->L.n, d_hm + i0 * ctx->L.n, d_z,
                        d_status + i0, st);
    if (rc) break;
    rc = prove_device_z(ctx, m, d_z, d_r + 4 * i0, d_s + 4 * i0, nullptr, nullptr, nullptr, st,
                        d_partials + i0 * PROOF_MSM_WORDS);
  }
  return rc;
}

// One proof, sharded proving key, witness map shared between the shards.  begin: witness generation, the A / B chains of
// this shard, and the ifft + coset fft of the vectors v of (a, b, c) with v % n_shards == shard, left in d_abc[v].  The
// caller then makes every vector known to every shard (ncclBroadcast of d_abc[v] from shard v % n_shards, on `stream`
// or ordered after it).  finish: pointwise product, coset ifft, the L + H chain of this shard, the MSM sums to d_partials.
int32_t frcs_prove_split_begin_dev(frcs_ctx* ctx, const uint16_t* d_sig, const uint16_t* d_pk, const uint16_t* d_hm,
                                   const uint64_t* d_r, const uint64_t* d_s, uint64_t* d_abc, int32_t* d_status,
                                   void* stream) {
  if (!ctx || !d_sig || !d_pk || !d_hm || !d_r || !d_s || !d_abc || !d_status) return FRCS_E_INVALID_ARG;
  FRCS_CUDA_CHECK(cudaSetDevice(ctx->device));
  if (!ctx->has_pk) {
    frcs_set_error("no proving key loaded (frcs_load_pk)");
    return FRCS_E_NO_PK;
  }
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* io;
  int32_t rc = ensure_io(ctx, al256((size_t)ctx->L.n_z * 32), &io);
  if (rc) return rc;
  if ((rc = ensure_prover(ctx, 1))) return rc;
  uint64_t* d_z = (uint64_t*)io;
  if ((rc = launch_witness(ctx, 1, d_sig, d_pk, d_hm, d_z, d_status, st))) return rc;
  if ((rc = group_head(ctx, 1, d_z, (const uint32_t*)d_r, (const uint32_t*)d_s, 0, st))) return rc;
  uint32_t mask = 0;
  for (uint32_t v = 0; v < 3; v++)
    if (v % ctx->shard.n == ctx->shard.idx) mask |= 1u << v;
  return launch_witness_map_head(ctx, d_z, (uint32_t*)d_abc, mask, st);
}

int32_t frcs_prove_split_finish_dev(frcs_ctx* ctx, uint64_t* d_abc, uint64_t* d_partials, void* stream) {
  if (!ctx || !d_abc || !d_partials) return FRCS_E_INVALID_ARG;
  FRCS_CUDA_CHECK(cudaSetDevice(ctx->device));
  if (!ctx->has_pk || !ctx->prover_ready || !ctx->prover.io) {
    frcs_set_error("frcs_prove_split_finish_dev without frcs_prove_split_begin_dev");
    return FRCS_E_INVALID_ARG;
  }
  cudaStream_t st = (cudaStream_t)stream;
  ProverState& P = ctx->prover;
  int32_t rc;
  if ((rc = launch_witness_map_tail(ctx, (uint32_t*)d_abc, (uint64_t*)P.h, st))) return rc;
  if ((rc = group_tail(ctx, 1, (const uint64_t*)P.io, 0, st))) return rc;
  for (int i = 0; i < 3; i++) FRCS_CUDA_CHECK(cudaStreamWaitEvent(st, P.done[0][i], 0));
  FRCS_CUDA_CHECK(cudaMemcpyAsync(d_partials, P.results, PROOF_MSM_WORDS * 8, cudaMemcpyDeviceToDevice, st));
  return FRCS_OK;
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
