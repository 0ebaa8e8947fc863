// Internal: the context object behind the C ABI, and the launcher prototypes of each
// kernel group.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <string>
#include <vector>

#include "../../include/falcon_r1cs_b200.h"
#include "circuit.hpp"

#define FRCS_MAX_NORM_OPS 32

struct DevCSR {
  uint32_t* row_ptr = nullptr;
  uint32_t* col = nullptr;
  uint32_t* val = nullptr;  // nnz x 8 u32, Montgomery
  uint64_t nnz = 0;
};

// Classified terms of one circuit matrix for the fast R1CS evaluation (spmv.cu): per non-zero a
// column and a code.  code bit 30 clear: coefficient = +-mag (bit 31 = negative, mag < 2^30), col = z
// column.  code bit 30 set: full-width coefficient fval[code & 0x3fffffff] (Montgomery), col = index into
// the context's small-column table (the multiplicand is expected to be a small integer).
struct DevBundles {  // per bundle: 4 row ids, term range, wide matrix, multiplicand limit, 4 x 2 extra terms; per term: column, 20 digits
  uint32_t *rows = nullptr, *ptr = nullptr, *cols = nullptr, *wide = nullptr, *limit = nullptr, *extra = nullptr, *dbl = nullptr;
  uint64_t* rec_off = nullptr;
  void* rec = nullptr;
  uint32_t n = 0, max_terms = 0;
  uint32_t n_rest = 0;  // the first n_rest bundles hold no row of an ntt_circuit block
};
struct DevTerms {
  uint32_t* row_ptr = nullptr;
  uint32_t* col = nullptr;
  uint32_t* code = nullptr;
  uint32_t* fval = nullptr;
  uint32_t* full_end = nullptr;  // per row: index of its first term that is not in the full-coefficient form
  uint64_t nnz = 0, n_full = 0;
};

// Precomputed MSM bases: for every base P_i and window k, 2^(cb k) P_i in affine form (cb = window bits, msm.hpp).
struct DevBases {
  void* pts = nullptr;  // [windows][n] affine (G1: 24 u32, G2: 48 u32 each)
  uint64_t n = 0;
  int windows = 0;
  int cb = 16;
  bool g2 = false;
};

// Twiddle / scaling tables of one Radix2EvaluationDomain size (ntt.cu)
struct NttPlan {
  uint32_t L = 0;
  uint32_t* consts = nullptr;  // w, w^-1, 1/n, 1/(g^n-1), g, g^-1
  uint32_t* tw_fwd = nullptr;  // w^i,  i < n/2
  uint32_t* tw_inv = nullptr;  // w^-i, i < n/2
  uint32_t* cp = nullptr;      // g^i / n
  uint32_t* cpi = nullptr;     // g^-i / n
  uint32_t* cpz = nullptr;     // g^i / (g^n - 1)
};

// A, B, C of one stand-alone gadget circuit (gadgets.cu), built on first use
struct DevGadget {
  DevCSR m[3];
  uint32_t n_rows = 0, n_wit_total = 0;  // witnesses after the operands (incl. the optional `expected`)
  bool ready = false;
};

struct NormOpsDev {
  uint8_t kind[FRCS_MAX_NORM_OPS], a[FRCS_MAX_NORM_OPS], b[FRCS_MAX_NORM_OPS];
};

// u64 words of the MSM results of one proof: A | B1 | L+H | (unused, infinity) (G1 XYZZ, 24 each) | B2 (G2 XYZZ, 48)
#define PROOF_MSM_WORDS 144

struct ProverState {
  uint32_t cap = 0;  // proofs per group the buffers below are sized for
  cudaStream_t streams[5] = {};  // [0] z sort + a/b1, [1] b2, [2] l+h accumulation (low priority), [3] l+h sort
  cudaEvent_t done[2][3] = {}, fork = nullptr, sorted_z = nullptr, sorted_zb = nullptr, sorted_lh = nullptr, z_ready = nullptr,
              copied[2] = {};
  void* ntt_work = nullptr;   // cap x 3 x domain Fr
  void* h = nullptr;          // cap x domain Fr
  void* extras = nullptr;     // 2 slots x cap x 5 scalars
  void* results = nullptr;    // 2 slots x cap x PROOF_MSM_WORDS u64 (device)
  uint64_t* h_results = nullptr;  // pinned host copy
  // MSM work (cap problems each): [0] sort of z (x 2: group parity), [1] G1 accumulation of a+b1, [2] G2 accumulation
  // of b2, [3] sort of (w | -rs | h) (x 2), [4] G1 accumulation of l+h, [5] sort of z without the scalars whose b_g1 /
  // b_g2 base is the point at infinity (x 2; only when the key has enough of them, frcs_ctx::b_skip)
  void* msm_work[6] = {};
  size_t sort_stride[2] = {};  // bytes between the two copies of msm_work[0] / msm_work[3]
  // staging of the host entry points, grown on demand
  void* io = nullptr;
  size_t io_bytes = 0;
};

// Optional per-stage timing with CUDA events on the launching stream (frcs_profile_*).
enum {
  PROF_WITNESS = 0, PROF_R1CS = 1, PROF_WITNESS_MAP = 2, PROF_MSM_H_ACCUM = 3, PROF_MSM_H = 4, PROF_MSM_A = 5,
  PROF_MSM_B1 = 6, PROF_MSM_L = 7, PROF_MSM_B2 = 8, PROF_HOST_TAIL = 9, PROF_NTT = 10, PROF_SORT_Z = 11, PROF_SORT_LH = 12,
  PROF_GROUP = 13, PROF_SORT_DIGITS = 14, PROF_SORT_SCATTER = 15, PROF_IDS = 16
};
struct Profiler {
  bool on = false;
  struct Span { int id; cudaEvent_t a, b; };
  std::vector<cudaEvent_t> pool;
  std::vector<Span> spans;
  double ms[PROF_IDS] = {};
  uint64_t count[PROF_IDS] = {};
  uint64_t work[PROF_IDS] = {};           // id-specific work counter (e.g. bucket additions)
  const uint32_t* work_dev[PROF_IDS] = {};  // device location of the last launch's work counter
  uint64_t work_mul[PROF_IDS] = {};         // problems per launch (the counter is read for problem 0)
};

struct frcs_ctx {
  int device = 0;
  circuit::Layout L;
  NormOpsDev norm_ops;
  uint32_t domain_log2 = 0;
  uint64_t launches = 0;
  cudaStream_t stream = nullptr;  // internal stream for the host entry points

  // circuit matrices
  DevCSR A, B, C;
  std::vector<uint32_t> long_rows_host;  // rows of A handled one-warp-per-row
  uint32_t* long_rows = nullptr;
  uint32_t n_long_rows = 0;
  DevTerms TA, TB, TC;
  // long rows whose wide matrix has integer coefficients below 2^159 in magnitude (the inlined NTT rows): signed
  // base-2^32 digit records (r1cs_signed_long_kernel); gl_rows = the other long rows (generic warp-per-row kernel)
  uint32_t *sl_rows = nullptr, *sl_ptr = nullptr, *sl_rec = nullptr, *sl_wide = nullptr, *sl_limit = nullptr, *gl_rows = nullptr;
  uint32_t n_sl_rows = 0, n_gl_rows = 0;
  // the same rows bundled four at a time by identical column lists (r1cs_bundle_kernel, batches of >= 64 signatures):
  DevBundles bd;
  uint32_t* bd_sums = nullptr;  // integer row sums between the two passes: [bundle slot][signature][8 words]
  size_t bd_sums_bytes = 0;
  bool bundles_usable = false;
  uint32_t* is_long = nullptr;     // bitmap over rows: handled by the warp-per-row kernel
  uint32_t* small_cols = nullptr;  // z columns multiplied by full-width coefficients (sig / v inputs, One)
  uint32_t n_small = 0;
  uint32_t* xs = nullptr;          // [signatures][n_small] canonical values of those columns (0xffffffff: not small)
  size_t xs_bytes = 0;
  uint32_t* r_perm = nullptr;  // short rows in class order
  uint32_t n_short_rows = 0;
  uint32_t* r_pm1 = nullptr;   // rows whose terms are all +-1 with at most one of each sign per matrix: 8 words per row
  uint32_t n_pm1_rows = 0;
  uint32_t *r_hdr = nullptr, *r_mterm = nullptr, *r_mfval = nullptr;  // merged short-row program
  // the ntt_circuit blocks of the circuit (circuit::NttBlock) and the tables of r1cs_ntt_rows_kernel: the N twiddles
  // and the LOG_N bound constants 2^(l+1) q^(l+2) as field elements (Montgomery)
  std::vector<circuit::NttBlock> ntt_blocks;
  uint32_t *ntt_tw_mont = nullptr, *ntt_cst_mont = nullptr;
  uint32_t n_sl_rest = 0, n_gl_rest = 0;  // the first n_*_rest long rows of each list are not rows of an ntt_circuit block
  bool ntt_rows_usable = false;
  // plan of the streaming short-row kernel (spmv.cu: r1cs_stream_kernel): per-window row programs
  void *stream_wins = nullptr, *stream_desc = nullptr;
  uint32_t n_stream_win = 0, stream_slots = 0, stream_desc_max = 0, stream_qmax = 0;
  size_t stream_smem = 0;
  bool stream_usable = false;
  // witness-gen tables
  uint32_t* ntt_tab = nullptr;  // [N] forward twiddles, [N] inverse twiddles
  void* check_z = nullptr;      // assignments of frcs_witness_check_batch (never leave the device)
  size_t check_z_bytes = 0;
  void* wit_scratch = nullptr;  // per-CTA parking space of the witness kernel (mod_q quotients)
  size_t wit_scratch_bytes = 0;
  uint32_t* mont_tab = nullptr; // [2][2^14] Fr: mont(j), mont(2^14 j)  (schoolbook witness kernel)
  std::vector<NttPlan> plans;  // Fr NTT tables per domain size
  DevGadget gadgets[circuit::GADGET_COUNT][2];  // [gadget][with the test macros' expected-output row]
  // proving key
  bool has_pk = false;
  // pre-processed base tables: a, b_g1, b_g2 = query ++ (1-base, r-base, s-base); lh = l_query ++ delta_1 ++ h_query
  DevBases pk_a, pk_b1, pk_b2, pk_lh;
  // bit i set: base i of the b_g1 and b_g2 tables is the point at infinity (the variable has no entry in the B matrix:
  // ~41 % of the Falcon circuits' columns).  The B MSMs then sort z without those scalars (their own sort) instead of
  // sharing the a_query sort: a lane that meets an infinity base idles while its warp performs a full addition.
  // nullptr: too few to pay for the second sort.
  uint32_t* b_skip = nullptr;
  uint64_t b_skip_count = 0;
  uint32_t* red_corr[2] = {nullptr, nullptr};  // -RED_CORR * generator (G1, G2), see msm_impl.cuh
  // base-range shard of the proving key held by this context (single-proof multi-GPU mode); {0, 1} = all
  struct Shard {
    uint32_t idx = 0, n = 1;
    uint64_t z_lo = 0, z_n = 0, l_lo = 0, l_n = 0, h_lo = 0, h_n = 0;
  } shard;
  ProverState prover;
  Profiler prof;
  bool prover_ready = false;
  // scratch, grown on demand
  void* scratch = nullptr;
  size_t scratch_bytes = 0;
  // staging of frcs_witness_check_batch (host entry point): two halves, copy stream, events
  void* hostio = nullptr;
  size_t hostio_bytes = 0;
  cudaStream_t io_stream = nullptr;
  cudaEvent_t io_in[2] = {}, io_done[2] = {};
};

// Own bounds checks for the kernels written this round (compute-sanitizer is closed on the GPU pool): a build with
// -DFRCS_BOUNDS turns these into device-side asserts; the GPU test-suite is then run against that build
// (tools/gpu_bounds.sh, profiles/r02_bounds_build_tests.txt).
#ifdef FRCS_BOUNDS
#include <cassert>
#define FRCS_ASSERT(c) assert(c)
#else
#define FRCS_ASSERT(c) ((void)0)
#endif

void frcs_set_error(const std::string& msg);
#define FRCS_CUDA_CHECK(expr)                                                                   \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess) {                                                                    \
      frcs_set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));                       \
      return FRCS_E_CUDA;                                                                       \
    }                                                                                           \
  } while (0)

int prof_begin(frcs_ctx* ctx, int id, cudaStream_t st);
void prof_end(frcs_ctx* ctx, int handle, cudaStream_t st);

// witness.cu
int32_t ensure_mont_table(frcs_ctx* ctx, cudaStream_t st);
int32_t launch_witness(frcs_ctx* ctx, uint64_t n, const uint16_t* d_sig, const uint16_t* d_pk, const uint16_t* d_hm,
                       uint64_t* d_z, int32_t* d_status, cudaStream_t st);
// ntt.cu
int32_t ensure_scratch(frcs_ctx* ctx, size_t bytes);
// nb assignments (z_stride u64 words apart) -> nb h vectors (2^domain_log2 Fr each, contiguous); work: nb x 3 x domain Fr
int32_t launch_witness_map(frcs_ctx* ctx, uint32_t nb, const uint64_t* d_z, uint64_t* d_h, uint32_t* work, cudaStream_t st);
int32_t launch_witness_map_head(frcs_ctx* ctx, const uint64_t* d_z, uint32_t* work, uint32_t vec_mask, cudaStream_t st);
int32_t launch_witness_map_tail(frcs_ctx* ctx, uint32_t* work, uint64_t* d_h, cudaStream_t st);
// spmv.cu
int32_t build_fast_r1cs(frcs_ctx* ctx, const circuit::Matrices& m);
void free_fast_r1cs(frcs_ctx* ctx);
int32_t launch_to_montgomery(frcs_ctx* ctx, uint32_t* d_vals, uint64_t count, cudaStream_t st);
int32_t launch_matvec3(frcs_ctx* ctx, const DevCSR* m, uint32_t n_rows, const uint32_t* d_long, uint32_t n_long,
                       const uint32_t* d_x, uint32_t* ya, uint32_t* yb, uint32_t* yc, cudaStream_t st);
int32_t launch_r1cs_eval(frcs_ctx* ctx, uint64_t n, const uint64_t* d_z, uint64_t* d_az, uint64_t* d_bz,
                         uint64_t* d_cz, int64_t* d_first_unsat, cudaStream_t st, uint64_t out_stride = 0);
