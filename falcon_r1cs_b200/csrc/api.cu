// C ABI entry points (include/falcon_r1cs_b200.h): context, matrices, and the host /
// device-pointer wrappers of each subsystem.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "ctx.hpp"

static thread_local std::string g_last_error;
void frcs_set_error(const std::string& msg) { g_last_error = msg; }

int prof_begin(frcs_ctx* ctx, int id, cudaStream_t st) {
  Profiler& P = ctx->prof;
  if (!P.on) return -1;
  cudaEvent_t ev[2];
  for (int k = 0; k < 2; k++) {
    if (P.pool.empty()) {
      cudaEventCreate(&ev[k]);
    } else {
      ev[k] = P.pool.back();
      P.pool.pop_back();
    }
  }
  cudaEventRecord(ev[0], st);
  P.spans.push_back({id, ev[0], ev[1]});
  return (int)P.spans.size() - 1;
}
void prof_end(frcs_ctx* ctx, int handle, cudaStream_t st) {
  if (handle < 0) return;
  cudaEventRecord(ctx->prof.spans[handle].b, st);
}

namespace {

int32_t upload_csr(frcs_ctx* ctx, const circuit::HostCSR& h, DevCSR* d) {
  d->nnz = h.col.size();
  FRCS_CUDA_CHECK(cudaMalloc(&d->row_ptr, h.row_ptr.size() * 4));
  FRCS_CUDA_CHECK(cudaMalloc(&d->col, (h.col.size() + 1) * 4));
  FRCS_CUDA_CHECK(cudaMalloc(&d->val, (h.val.size() + 1) * 32));
  FRCS_CUDA_CHECK(cudaMemcpy(d->row_ptr, h.row_ptr.data(), h.row_ptr.size() * 4, cudaMemcpyHostToDevice));
  FRCS_CUDA_CHECK(cudaMemcpy(d->col, h.col.data(), h.col.size() * 4, cudaMemcpyHostToDevice));
  FRCS_CUDA_CHECK(cudaMemcpy(d->val, h.val.data(), h.val.size() * 32, cudaMemcpyHostToDevice));
  return launch_to_montgomery(ctx, d->val, d->nnz, ctx->stream);
}
void free_csr(DevCSR* d) {
  cudaFree(d->row_ptr);
  cudaFree(d->col);
  cudaFree(d->val);
}

struct DevBuf {  // RAII device allocation for the host entry points
  void* p = nullptr;
  ~DevBuf() {
    if (p) cudaFree(p);
  }
  cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1); }
  template <class T>
  T* as() {
    return (T*)p;
  }
};

}  // namespace

extern "C" {

const char* frcs_last_error(void) { return g_last_error.c_str(); }

// Test hook, host only (no GPU): the circuit compiler's matrices (csrc/circuit.hpp) for differential tests against
// cs.to_matrices().  counts (optional): n_instance, n_witness, n_constraints, nnz of matrix `which`; the other
// buffers may be NULL to query the counts first.  val: canonical integers (not Montgomery), 4 x u64 per entry.
int32_t frcs_debug_host_matrix(uint32_t logn, uint32_t kind, int32_t which, uint32_t* row_ptr, uint32_t* col,
                               uint64_t* val, uint64_t* counts) {
  // kind 16 + g (+ 8): the stand-alone circuit of gadget g (with the test macros' expected-output row)
  const bool is_gadget = kind >= 16 && kind < 32 && ((kind - 16) & 7) < (uint32_t)circuit::GADGET_COUNT;
  if ((logn != 9 && logn != 10) || (kind > FRCS_KIND_DUAL_NTT && !is_gadget) || which < 0 || which > 2)
    return FRCS_E_INVALID_ARG;
  circuit::Matrices m = is_gadget ? circuit::Builder::build_gadget(logn, (int)((kind - 16) & 7), ((kind - 16) & 8) != 0)
                                  : circuit::Builder(logn, kind).build();
  const circuit::HostCSR& h = which == 0 ? m.a : which == 1 ? m.b : m.c;
  if (counts) {
    counts[0] = m.L.n_inst;
    counts[1] = m.L.n_wit;
    counts[2] = m.L.n_cons;
    counts[3] = h.col.size();
  }
  if (row_ptr) memcpy(row_ptr, h.row_ptr.data(), h.row_ptr.size() * 4);
  if (col) memcpy(col, h.col.data(), h.col.size() * 4);
  if (val) memcpy(val, h.val.data(), h.val.size() * 32);
  return FRCS_OK;
}

int32_t frcs_ctx_create(uint32_t logn, uint32_t kind, int32_t device, frcs_ctx** out) {
  if (!out || (logn != 9 && logn != 10)) {
    frcs_set_error("frcs_ctx_create: logn must be 9 (Falcon-512) or 10 (Falcon-1024)");
    return FRCS_E_INVALID_ARG;
  }
  if (kind != FRCS_KIND_NTT && kind != FRCS_KIND_SCHOOLBOOK && kind != FRCS_KIND_DUAL_NTT) {
    frcs_set_error("frcs_ctx_create: kind must be FRCS_KIND_NTT, FRCS_KIND_SCHOOLBOOK or FRCS_KIND_DUAL_NTT");
    return FRCS_E_INVALID_ARG;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    frcs_set_error("frcs_ctx_create: no CUDA device (this library has no CPU fallback)");
    return FRCS_E_CUDA;
  }
  FRCS_CUDA_CHECK(cudaSetDevice(device));
  if (getenv("FRCS_STACK")) FRCS_CUDA_CHECK(cudaDeviceSetLimit(cudaLimitStackSize, atoi(getenv("FRCS_STACK"))));
  frcs_ctx* ctx = new (std::nothrow) frcs_ctx;
  if (!ctx) return FRCS_E_ALLOC;
  struct Guard {  // a failure below (every FRCS_CUDA_CHECK returns) must not leak the half-built context
    frcs_ctx* c;
    ~Guard() {
      if (c) frcs_ctx_destroy(c);
    }
  } guard{ctx};
  ctx->device = device;
  FRCS_CUDA_CHECK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  circuit::Builder b(logn, kind);
  circuit::Matrices m = b.build();
  ctx->L = m.L;
  ctx->domain_log2 = 0;
  while ((1ull << ctx->domain_log2) < (uint64_t)m.L.n_cons + m.L.n_inst) ctx->domain_log2++;
  circuit::NormProgram np = circuit::norm_program(logn);
  for (size_t k = 0; k < np.ops.size(); k++) {
    ctx->norm_ops.kind[k] = np.ops[k].kind;
    ctx->norm_ops.a[k] = np.ops[k].a;
    ctx->norm_ops.b[k] = np.ops[k].b;
  }
  int32_t rc;
  if ((rc = upload_csr(ctx, m.a, &ctx->A)) || (rc = upload_csr(ctx, m.b, &ctx->B)) ||
      (rc = upload_csr(ctx, m.c, &ctx->C)))
    return rc;
  // long rows, those of the ntt_circuit blocks last (the R1CS evaluator can take them through the butterfly network
  // and then runs its generic long-row kernels on the leading part of every list only)
  ctx->ntt_blocks = m.ntt_blocks;
  auto in_ntt_block = [&](uint32_t r) {
    for (const circuit::NttBlock& b : m.ntt_blocks)
      if (r >= b.row0 && (r - b.row0) % 30 == 0 && (r - b.row0) / 30 < m.L.n) return true;
    return false;
  };
  for (int pass = 0; pass < 2; pass++)
    for (uint32_t r = 0; r < m.L.n_cons; r++)
      if ((m.a.row_ptr[r + 1] - m.a.row_ptr[r] > 64 || m.b.row_ptr[r + 1] - m.b.row_ptr[r] > 64 ||
           m.c.row_ptr[r + 1] - m.c.row_ptr[r] > 64) &&
          in_ntt_block(r) == (pass == 1))
        ctx->long_rows_host.push_back(r);
  ctx->n_long_rows = (uint32_t)ctx->long_rows_host.size();
  if ((rc = build_fast_r1cs(ctx, m))) return rc;
  FRCS_CUDA_CHECK(cudaMalloc(&ctx->long_rows, (ctx->n_long_rows + 1) * 4));
  FRCS_CUDA_CHECK(cudaMemcpy(ctx->long_rows, ctx->long_rows_host.data(), ctx->n_long_rows * 4, cudaMemcpyHostToDevice));
  // Falcon NTT twiddles mod q: forward table and its element-wise inverse
  {
    std::vector<uint32_t> tab = circuit::ntt_table(m.L.n), both(2 * m.L.n);
    for (uint32_t i = 0; i < m.L.n; i++) {
      both[i] = tab[i];
      both[m.L.n + i] = circuit::powmod_q(tab[i], circuit::Q - 2);
    }
    FRCS_CUDA_CHECK(cudaMalloc(&ctx->ntt_tab, both.size() * 4));
    FRCS_CUDA_CHECK(cudaMemcpy(ctx->ntt_tab, both.data(), both.size() * 4, cudaMemcpyHostToDevice));
  }
  // the uploads above are legacy-stream copies from pageable memory (staged, then DMA): make sure every one has
  // landed before kernels on the context's non-blocking streams can read the tables
  FRCS_CUDA_CHECK(cudaDeviceSynchronize());
  guard.c = nullptr;
  *out = ctx;
  return FRCS_OK;
}

void frcs_ctx_destroy(frcs_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  free_csr(&ctx->A);
  free_csr(&ctx->B);
  free_csr(&ctx->C);
  cudaFree(ctx->long_rows);
  free_fast_r1cs(ctx);
  cudaFree(ctx->ntt_tab);
  cudaFree(ctx->ntt_tw_mont);
  cudaFree(ctx->ntt_cst_mont);
  cudaFree(ctx->mont_tab);
  cudaFree(ctx->wit_scratch);
  cudaFree(ctx->check_z);
  for (auto& p : ctx->plans) {
    cudaFree(p.consts);
    cudaFree(p.tw_fwd);
    cudaFree(p.tw_inv);
    cudaFree(p.cp);
    cudaFree(p.cpi);
    cudaFree(p.cpz);
  }
  for (auto& gv : ctx->gadgets)
    for (DevGadget& G : gv)
      for (DevCSR& m : G.m) free_csr(&m);
  cudaFree(ctx->pk_a.pts);
  cudaFree(ctx->pk_b1.pts);
  cudaFree(ctx->pk_b2.pts);
  cudaFree(ctx->pk_lh.pts);
  cudaFree(ctx->b_skip);
  if (ctx->prover_ready) {
    ProverState& P = ctx->prover;
    for (int i = 0; i < 5; i++) cudaStreamDestroy(P.streams[i]);
    for (int i = 0; i < 6; i++) cudaFree(P.msm_work[i]);
    for (int k = 0; k < 2; k++)
      for (int i = 0; i < 3; i++) cudaEventDestroy(P.done[k][i]);
    cudaEventDestroy(P.fork);
    cudaEventDestroy(P.sorted_z);
    cudaEventDestroy(P.sorted_zb);
    cudaEventDestroy(P.sorted_lh);
    cudaEventDestroy(P.z_ready);
    cudaEventDestroy(P.copied[0]);
    cudaEventDestroy(P.copied[1]);
    cudaFree(P.ntt_work);
    cudaFree(P.h);
    cudaFree(P.extras);
    cudaFree(P.results);
    cudaFreeHost(P.h_results);
  }
  cudaFree(ctx->prover.io);
  cudaFree(ctx->hostio);
  if (ctx->io_stream) {
    cudaStreamDestroy(ctx->io_stream);
    for (int i = 0; i < 2; i++) {
      cudaEventDestroy(ctx->io_in[i]);
      cudaEventDestroy(ctx->io_done[i]);
    }
  }
  cudaFree(ctx->red_corr[0]);
  cudaFree(ctx->red_corr[1]);
  cudaFree(ctx->scratch);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

int32_t frcs_shape_get(const frcs_ctx* ctx, frcs_shape* out) {
  if (!ctx || !out) return FRCS_E_INVALID_ARG;
  out->logn = ctx->L.logn;
  out->kind = ctx->L.kind;
  out->n_instance = ctx->L.n_inst;
  out->n_witness = ctx->L.n_wit;
  out->n_constraints = ctx->L.n_cons;
  out->domain_log2 = ctx->domain_log2;
  out->nnz_a = ctx->A.nnz;
  out->nnz_b = ctx->B.nnz;
  out->nnz_c = ctx->C.nnz;
  return FRCS_OK;
}

int32_t frcs_get_matrix(frcs_ctx* ctx, int32_t which, uint32_t* row_ptr, uint32_t* col, uint64_t* val) {
  if (!ctx || which < 0 || which > 2) return FRCS_E_INVALID_ARG;
  FRCS_CUDA_CHECK(cudaSetDevice(ctx->device));
  const DevCSR& m = which == 0 ? ctx->A : which == 1 ? ctx->B : ctx->C;
  if (row_ptr) FRCS_CUDA_CHECK(cudaMemcpy(row_ptr, m.row_ptr, (ctx->L.n_cons + 1) * 4, cudaMemcpyDeviceToHost));
  if (col) FRCS_CUDA_CHECK(cudaMemcpy(col, m.col, m.nnz * 4, cudaMemcpyDeviceToHost));
  if (val) FRCS_CUDA_CHECK(cudaMemcpy(val, m.val, m.nnz * 32, cudaMemcpyDeviceToHost));
  return FRCS_OK;
}

int32_t frcs_witness_batch_dev(frcs_ctx* ctx, uint64_t n, const uint16_t* d_sig, const uint16_t* d_pk,
                               const uint16_t* d_hm, uint64_t* d_z, int32_t* d_status, void* stream) {
  if (!ctx || !d_sig || !d_pk || !d_hm || !d_z || !d_status) return FRCS_E_INVALID_ARG;
  FRCS_CUDA_CHECK(cudaSetDevice(ctx->device));
  return launch_witness(ctx, n, d_sig, d_pk, d_hm, d_z, d_status, (cudaStream_t)stream);
}

// signatures per pass of the host entry points: bounds the device buffers (1184 x 5.08 MB = 6 GB for Falcon-1024)
static uint64_t host_chunk(const frcs_ctx* ctx) {
  const uint64_t zb = (uint64_t)ctx->L.n_z * 32;
  uint64_t ch = (6ull << 30) / zb;
  return ch < 1 ? 1 : ch;
}

int32_t frcs_witness_batch(frcs_ctx* ctx, uint64_t n, const uint16_t* sig, const uint16_t* pk, const uint16_t* hm,
                           uint64_t* z_out, int32_t* status) {
  if (!ctx || !sig || !pk || !hm || !z_out || !status) return FRCS_E_INVALID_ARG;
  FRCS_CUDA_CHECK(cudaSetDevice(ctx->device));
  const uint64_t ch = std::min<uint64_t>(n ? n : 1, host_chunk(ctx));
  const size_t in_b = ch * ctx->L.n * 2, z_b = ch * (size_t)ctx->L.n_z * 32;
  DevBuf d_sig, d_pk, d_hm, d_z, d_st;
  FRCS_CUDA_CHECK(d_sig.alloc(in_b));
  FRCS_CUDA_CHECK(d_pk.alloc(in_b));
  FRCS_CUDA_CHECK(d_hm.alloc(in_b));
  FRCS_CUDA_CHECK(d_z.alloc(z_b));
  FRCS_CUDA_CHECK(d_st.alloc(ch * 4));
  cudaStream_t st = ctx->stream;
  for (uint64_t i0 = 0; i0 < n; i0 += ch) {
    const uint64_t m = std::min(ch, n - i0);
    const size_t ib = m * ctx->L.n * 2;
    FRCS_CUDA_CHECK(cudaMemcpyAsync(d_sig.p, sig + i0 * ctx->L.n, ib, cudaMemcpyHostToDevice, st));
    FRCS_CUDA_CHECK(cudaMemcpyAsync(d_pk.p, pk + i0 * ctx->L.n, ib, cudaMemcpyHostToDevice, st));
    FRCS_CUDA_CHECK(cudaMemcpyAsync(d_hm.p, hm + i0 * ctx->L.n, ib, cudaMemcpyHostToDevice, st));
    int32_t rc = launch_witness(ctx, m, d_sig.as<uint16_t>(), d_pk.as<uint16_t>(), d_hm.as<uint16_t>(),
                                d_z.as<uint64_t>(), d_st.as<int32_t>(), st);
    if (rc) return rc;
    FRCS_CUDA_CHECK(cudaMemcpyAsync(z_out + i0 * ctx->L.n_z * 4, d_z.p, m * (size_t)ctx->L.n_z * 32,
                                    cudaMemcpyDeviceToHost, st));
    FRCS_CUDA_CHECK(cudaMemcpyAsync(status + i0, d_st.p, m * 4, cudaMemcpyDeviceToHost, st));
    FRCS_CUDA_CHECK(cudaStreamSynchronize(st));
  }
  return FRCS_OK;
}

// BASELINE configs[2]: batched witness generation + R1CS satisfaction, no assignment leaves the device.
// Inputs on the device; z lives in a per-call buffer of at most host_chunk() signatures.
// (Measured and dropped: sub-chunks with two assignment buffers, the witness kernel of sub-chunk k + 1 on the caller's
// stream next to the satisfaction check of sub-chunk k on a second stream: 249 k instead of 282 k witnesses/s at 592
// signatures, 267 k instead of 280 k/s over 65,536 -- the two kernels take each other's SM slots and HBM queue.)
int32_t frcs_witness_check_batch_dev(frcs_ctx* ctx, uint64_t n, const uint16_t* d_sig, const uint16_t* d_pk,
                                     const uint16_t* d_hm, int64_t* d_first_unsat, int32_t* d_status, void* stream) {
  if (!ctx || !d_sig || !d_pk || !d_hm || !d_first_unsat || !d_status) return FRCS_E_INVALID_ARG;
  FRCS_CUDA_CHECK(cudaSetDevice(ctx->device));
  cudaStream_t st = (cudaStream_t)stream;
  const uint64_t ch = std::min<uint64_t>(n ? n : 1, host_chunk(ctx));
  const size_t z_b = ch * (size_t)ctx->L.n_z * 32;
  if (ctx->check_z_bytes < z_b) {
    FRCS_CUDA_CHECK(cudaDeviceSynchronize());
    cudaFree(ctx->check_z);
    ctx->check_z = nullptr;
    ctx->check_z_bytes = 0;
    FRCS_CUDA_CHECK(cudaMalloc(&ctx->check_z, z_b));
    ctx->check_z_bytes = z_b;
  }
  for (uint64_t i0 = 0; i0 < n; i0 += ch) {
    const uint64_t m = std::min(ch, n - i0);
    int32_t rc = launch_witness(ctx, m, d_sig + i0 * ctx->L.n, d_pk + i0 * ctx->L.n, d_hm + i0 * ctx->L.n,
                                (uint64_t*)ctx->check_z, d_status + i0, st);
    if (rc) return rc;
    rc = launch_r1cs_eval(ctx, m, (const uint64_t*)ctx->check_z, nullptr, nullptr, nullptr, d_first_unsat + i0, st);
    if (rc) return rc;
  }
  return FRCS_OK;
}

int32_t frcs_witness_check_batch(frcs_ctx* ctx, uint64_t n, const uint16_t* sig, const uint16_t* pk, const uint16_t* hm,
                                 int64_t* first_unsat, int32_t* status) {
  if (!ctx || !sig || !pk || !hm || !first_unsat || !status) return FRCS_E_INVALID_ARG;
  FRCS_CUDA_CHECK(cudaSetDevice(ctx->device));
  const uint64_t ch = std::min<uint64_t>(n ? n : 1, 8192);
  const size_t in_b = (ch * ctx->L.n * 2 + 255) & ~(size_t)255;
  // persistent staging (grown on demand): no cudaMalloc / cudaFree on the path of a call; two halves so that the
  // copies of chunk k+1 overlap the kernels of chunk k
  const size_t half = 3 * in_b + ((ch * 8 + 255) & ~(size_t)255) + ((ch * 4 + 255) & ~(size_t)255);
  if (ctx->hostio_bytes < 2 * half) {
    FRCS_CUDA_CHECK(cudaDeviceSynchronize());
    cudaFree(ctx->hostio);
    ctx->hostio = nullptr;
    ctx->hostio_bytes = 0;
    FRCS_CUDA_CHECK(cudaMalloc(&ctx->hostio, 2 * half));
    ctx->hostio_bytes = 2 * half;
  }
  if (!ctx->io_stream) {
    FRCS_CUDA_CHECK(cudaStreamCreateWithFlags(&ctx->io_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; i++) {
      FRCS_CUDA_CHECK(cudaEventCreateWithFlags(&ctx->io_in[i], cudaEventDisableTiming));
      FRCS_CUDA_CHECK(cudaEventCreateWithFlags(&ctx->io_done[i], cudaEventDisableTiming));
    }
  }
  cudaStream_t st = ctx->stream, io = ctx->io_stream;
  int k = 0;
  for (uint64_t i0 = 0; i0 < n; i0 += ch, k++) {
    const uint64_t m = std::min(ch, n - i0);
    const size_t ib = m * ctx->L.n * 2;
    uint8_t* base = (uint8_t*)ctx->hostio + (size_t)(k & 1) * half;
    uint16_t *d_sig = (uint16_t*)base, *d_pk = (uint16_t*)(base + in_b), *d_hm = (uint16_t*)(base + 2 * in_b);
    int64_t* d_fu = (int64_t*)(base + 3 * in_b);
    int32_t* d_st = (int32_t*)(base + 3 * in_b + ((ch * 8 + 255) & ~(size_t)255));
    // inputs of this chunk on the copy stream (after the kernels that last read this half: two chunks ago)
    if (k >= 2) FRCS_CUDA_CHECK(cudaStreamWaitEvent(io, ctx->io_done[k & 1], 0));
    FRCS_CUDA_CHECK(cudaMemcpyAsync(d_sig, sig + i0 * ctx->L.n, ib, cudaMemcpyHostToDevice, io));
    FRCS_CUDA_CHECK(cudaMemcpyAsync(d_pk, pk + i0 * ctx->L.n, ib, cudaMemcpyHostToDevice, io));
    FRCS_CUDA_CHECK(cudaMemcpyAsync(d_hm, hm + i0 * ctx->L.n, ib, cudaMemcpyHostToDevice, io));
    FRCS_CUDA_CHECK(cudaEventRecord(ctx->io_in[k & 1], io));
    FRCS_CUDA_CHECK(cudaStreamWaitEvent(st, ctx->io_in[k & 1], 0));
    int32_t rc = frcs_witness_check_batch_dev(ctx, m, d_sig, d_pk, d_hm, d_fu, d_st, st);
    if (rc) return rc;
    FRCS_CUDA_CHECK(cudaMemcpyAsync(first_unsat + i0, d_fu, m * 8, cudaMemcpyDeviceToHost, st));
    FRCS_CUDA_CHECK(cudaMemcpyAsync(status + i0, d_st, m * 4, cudaMemcpyDeviceToHost, st));
    FRCS_CUDA_CHECK(cudaEventRecord(ctx->io_done[k & 1], st));
  }
  FRCS_CUDA_CHECK(cudaStreamSynchronize(st));
  return FRCS_OK;
}

int32_t frcs_r1cs_eval_batch_dev(frcs_ctx* ctx, uint64_t n, const uint64_t* d_z, uint64_t* d_az, uint64_t* d_bz,
                                 uint64_t* d_cz, int64_t* d_first_unsat, void* stream) {
  if (!ctx || !d_z) return FRCS_E_INVALID_ARG;
  FRCS_CUDA_CHECK(cudaSetDevice(ctx->device));
  return launch_r1cs_eval(ctx, n, d_z, d_az, d_bz, d_cz, d_first_unsat, (cudaStream_t)stream);
}

int32_t frcs_r1cs_eval_batch(frcs_ctx* ctx, uint64_t n, const uint64_t* z, uint64_t* az, uint64_t* bz, uint64_t* cz,
                             int64_t* first_unsat) {
  if (!ctx || !z) return FRCS_E_INVALID_ARG;
  FRCS_CUDA_CHECK(cudaSetDevice(ctx->device));
  // bounded device buffers: z + up to three result vectors per signature
  const size_t per = (size_t)ctx->L.n_z * 32 + (size_t)ctx->L.n_cons * 32 * ((az ? 1 : 0) + (bz ? 1 : 0) + (cz ? 1 : 0));
  uint64_t ch = (8ull << 30) / per;
  ch = std::min<uint64_t>(std::max<uint64_t>(ch, 1), n ? n : 1);
  const size_t z_b = ch * (size_t)ctx->L.n_z * 32, o_b = ch * (size_t)ctx->L.n_cons * 32;
  DevBuf d_z, d_a, d_b, d_c, d_fu;
  cudaStream_t st = ctx->stream;
  FRCS_CUDA_CHECK(d_z.alloc(z_b));
  if (az) FRCS_CUDA_CHECK(d_a.alloc(o_b));
  if (bz) FRCS_CUDA_CHECK(d_b.alloc(o_b));
  if (cz) FRCS_CUDA_CHECK(d_c.alloc(o_b));
  if (first_unsat) FRCS_CUDA_CHECK(d_fu.alloc(ch * 8));
  for (uint64_t i0 = 0; i0 < n; i0 += ch) {
    const uint64_t m = std::min(ch, n - i0);
    const size_t zb = m * (size_t)ctx->L.n_z * 32, ob = m * (size_t)ctx->L.n_cons * 32;
    const uint64_t oo = i0 * ctx->L.n_cons * 4;
    FRCS_CUDA_CHECK(cudaMemcpyAsync(d_z.p, z + i0 * ctx->L.n_z * 4, zb, cudaMemcpyHostToDevice, st));
    int32_t rc = launch_r1cs_eval(ctx, m, d_z.as<uint64_t>(), d_a.as<uint64_t>(), d_b.as<uint64_t>(),
                                  d_c.as<uint64_t>(), d_fu.as<int64_t>(), st);
    if (rc) return rc;
    if (az) FRCS_CUDA_CHECK(cudaMemcpyAsync(az + oo, d_a.p, ob, cudaMemcpyDeviceToHost, st));
    if (bz) FRCS_CUDA_CHECK(cudaMemcpyAsync(bz + oo, d_b.p, ob, cudaMemcpyDeviceToHost, st));
    if (cz) FRCS_CUDA_CHECK(cudaMemcpyAsync(cz + oo, d_c.p, ob, cudaMemcpyDeviceToHost, st));
    if (first_unsat) FRCS_CUDA_CHECK(cudaMemcpyAsync(first_unsat + i0, d_fu.p, m * 8, cudaMemcpyDeviceToHost, st));
    FRCS_CUDA_CHECK(cudaStreamSynchronize(st));
  }
  return FRCS_OK;
}

uint64_t frcs_launch_count(const frcs_ctx* ctx) { return ctx ? ctx->launches : 0; }

int32_t frcs_profile_enable(frcs_ctx* ctx, int32_t on) {
  if (!ctx) return FRCS_E_INVALID_ARG;
  ctx->prof.on = on != 0;
  return FRCS_OK;
}
// Drains the recorded spans (synchronising on them) and returns the accumulated device
// time, launch count and work counter of stage `id`; reset != 0 clears the accumulators.
int32_t frcs_profile_get(frcs_ctx* ctx, int32_t id, double* ms_total, uint64_t* count, uint64_t* work, int32_t reset) {
  if (!ctx || id < 0 || id >= PROF_IDS) return FRCS_E_INVALID_ARG;
  FRCS_CUDA_CHECK(cudaSetDevice(ctx->device));
  Profiler& P = ctx->prof;
  for (auto& sp : P.spans) {
    FRCS_CUDA_CHECK(cudaEventSynchronize(sp.b));
    float ms = 0;
    FRCS_CUDA_CHECK(cudaEventElapsedTime(&ms, sp.a, sp.b));
    P.ms[sp.id] += ms;
    P.count[sp.id]++;
    P.pool.push_back(sp.a);
    P.pool.push_back(sp.b);
  }
  P.spans.clear();
  if (P.work_dev[id]) {
    uint32_t w = 0;
    FRCS_CUDA_CHECK(cudaMemcpy(&w, P.work_dev[id], 4, cudaMemcpyDeviceToHost));
    P.work[id] = (uint64_t)w * (P.work_mul[id] ? P.work_mul[id] : 1);
  }
  if (ms_total) *ms_total = P.ms[id];
  if (count) *count = P.count[id];
  if (work) *work = P.work[id];
  if (reset) {
    P.ms[id] = 0;
    P.count[id] = 0;
  }
  return FRCS_OK;
}

}  // extern "C"
