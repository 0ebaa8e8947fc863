// Subsystem (3): the Groth16 witness map over the BLS12-381 scalar field.
// Replaces ark-poly 0.3.0's Radix2EvaluationDomain::{ifft,coset_fft,coset_ifft}_in_place
// and ark-groth16 0.3.0's R1CStoQAP::witness_map ([EXT]; reached from
// examples/pok_sig.rs:32 through create_proof).  SURVEY.md App. B.4/B.5.
//
// Layout: vectors of n = 2^L Fr elements (32 B each, Montgomery) in HBM/L2.  An NTT is
// 2-3 launches; each launch runs up to 10 radix-2 stages on a 1024-element tile held in
// shared memory (limb-major, conflict-free).  Forward and inverse transforms are paired
// as DIF (natural -> bit-reversed) and DIT (bit-reversed -> natural) so that no
// permutation pass is needed: ifft leaves coefficients bit-reversed, the coset scaling
// g^i/n is applied through bitrev(i), and the forward DIT returns to natural order.
#define FF_OPAQUE_M0  // (ff32.cuh: keeps the Fr multiplications of the butterflies on IMAD.WIDE)
#include "ctx.hpp"
#include "nvtx.hpp"
#define FF_INLINE_MUL
#include "ff32.cuh"

using ff::Fr;

namespace {

constexpr uint32_t TILE_LOG = 10, TILE = 1u << TILE_LOG, NT = 256;

__device__ __forceinline__ Fr ld_fr(const uint32_t* p) {
  Fr r;
  uint64_t a, b, c, d;
  asm volatile("ld.global.v4.b64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
  r.v[0] = (uint32_t)a; r.v[1] = (uint32_t)(a >> 32);
  r.v[2] = (uint32_t)b; r.v[3] = (uint32_t)(b >> 32);
  r.v[4] = (uint32_t)c; r.v[5] = (uint32_t)(c >> 32);
  r.v[6] = (uint32_t)d; r.v[7] = (uint32_t)(d >> 32);
  return r;
}
__device__ __forceinline__ void st_fr(uint32_t* p, const Fr& x) {
  uint64_t a = (uint64_t)x.v[0] | ((uint64_t)x.v[1] << 32), b = (uint64_t)x.v[2] | ((uint64_t)x.v[3] << 32);
  uint64_t c = (uint64_t)x.v[4] | ((uint64_t)x.v[5] << 32), d = (uint64_t)x.v[6] | ((uint64_t)x.v[7] << 32);
  asm volatile("st.global.v4.b64 [%0], {%1,%2,%3,%4};" ::"l"(p), "l"(a), "l"(b), "l"(c), "l"(d) : "memory");
}
__device__ __forceinline__ Fr lds_fr(const uint32_t* sm, uint32_t e) {
  Fr r;
#pragma unroll
  for (int k = 0; k < 8; k++) r.v[k] = sm[k * TILE + e];
  return r;
}
__device__ __forceinline__ void sts_fr(uint32_t* sm, uint32_t e, const Fr& x) {
#pragma unroll
  for (int k = 0; k < 8; k++) sm[k * TILE + e] = x.v[k];
}

// consts: [0] w_n  [1] w_n^-1  [2] 1/n  [3] 1/(g^n - 1)  [4] g  [5] g^-1  [6] n
__global__ void ntt_consts_kernel(uint32_t* consts, uint32_t L) {
  Fr w, wi, g, gi;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    w.v[i] = FrParams::ROOT_2_32(i);
    wi.v[i] = FrParams::ROOT_2_32_INV(i);
    g.v[i] = FrParams::GEN(i);
    gi.v[i] = FrParams::GEN_INV(i);
  }
  for (uint32_t i = L; i < 32; i++) {
    w = w.sqr();
    wi = wi.sqr();
  }
  Fr ninv = Fr::from_u32(1u << L).inverse();
  Fr gn = g;
  for (uint32_t i = 0; i < L; i++) gn = gn.sqr();
  Fr zinv = (gn - Fr::one()).inverse();
  st_fr(consts, w);
  st_fr(consts + 8, wi);
  st_fr(consts + 16, ninv);
  st_fr(consts + 24, zinv);
  st_fr(consts + 32, g);
  st_fr(consts + 40, gi);
  st_fr(consts + 48, Fr::from_u32(1u << L));
}
// out[i] = scale * base^i, i < count (scale may be null = 1)
__global__ void pow_table_kernel(uint32_t* out, const uint32_t* base, const uint32_t* scale, uint32_t count) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  Fr b = ld_fr(base), r = scale ? ld_fr(scale) : Fr::one();
  for (uint32_t e = i; e; e >>= 1) {
    if (e & 1) r = r * b;
    b = b.sqr();
  }
  st_fr(out + 8 * (uint64_t)i, r);
}

struct ChunkArgs {
  uint32_t L, s0, S, stride_log, g_log;
  uint64_t batch_stride;    // in u32 words
  const uint32_t* tw;       // w^i (or w^-i), i < n/2
  const uint32_t* scale;    // optional epilogue table, indexed by bitrev_L(position) (DIF) / position (DIT)
  uint32_t* scatter_out;    // optional: DIF epilogue writes element to bitrev_L(position) of this buffer
  uint64_t scatter_stride;  // batch stride of scatter_out (u32 words)
  // witness map without the coset transform of c (launch_witness_map):
  const uint32_t* scale_c;  // optional: epilogue table of every third vector (y % 3 == 2) instead of `scale`
  uint32_t ab_only;         // the batch is vectors 3k and 3k + 1 only: vector of block row y = (y / 2) * 3 + y % 2
  uint64_t sub_off;         // non-zero: the DIF epilogue subtracts the element at the same position sub_off words further on
};
struct NttExtra {
  const uint32_t* scale_c = nullptr;
  uint32_t ab_only = 0;
  uint64_t sub_off = 0;
};

// Runs stages [s0, s0+S) of a radix-2 NTT on the tile
//   i = outer * 2^(S+stride_log) + hi * 2^stride_log + inner0 + g,  hi < 2^S, g < 2^g_log.
// DIT = false: Gentleman-Sande (DIF), stage s pairs i, i + n/2^(s+1); stride_log = L-s0-S
// DIT = true : Cooley-Tukey (DIT), stage s pairs i, i + 2^s;          stride_log = s0
template <bool DIT>
__global__ void __launch_bounds__(NT) ntt_chunk_kernel(uint32_t* data_all, ChunkArgs A) {
  extern __shared__ uint32_t sm[];
  uint32_t* data = data_all + (A.ab_only ? (blockIdx.y >> 1) * 3 + (blockIdx.y & 1) : blockIdx.y) * A.batch_stride;
  const uint32_t tid = threadIdx.x;
  const uint32_t tile_elems = 1u << (A.S + A.g_log);
  const uint32_t tpo_log = A.stride_log - A.g_log;
  const uint32_t outer = blockIdx.x >> tpo_log;
  const uint32_t inner0 = (blockIdx.x & ((1u << tpo_log) - 1)) << A.g_log;
  const uint32_t base = (outer << (A.S + A.stride_log)) + inner0;
  const uint32_t gmask = (1u << A.g_log) - 1;

  for (uint32_t e = tid; e < tile_elems; e += NT) {
    uint32_t gi = base + ((e >> A.g_log) << A.stride_log) + (e & gmask);
    Fr x = ld_fr(data + 8 * (uint64_t)gi);
    if (DIT && A.scale && A.s0 == 0) x = x * ld_fr(A.scale + 8 * (uint64_t)(__brev(gi) >> (32 - A.L)));
    sts_fr(sm, e, x);
  }
  uint32_t t = 0;
#ifndef FRCS_NTT_RADIX2
  // Two stages at a time in registers: a thread takes the four elements that differ in the two pair bits (pl, pl + 1)
  // of `hi`, so a tile goes through shared memory and a barrier once per two stages (three twiddles per four
  // elements instead of four).  Same butterflies, same twiddles as the single stages below.
  for (; t + 1 < A.S; t += 2) {
    __syncthreads();
    const uint32_t pl = DIT ? t : (A.S - 2 - t);
    const uint32_t sh1 = DIT ? (A.L - A.s0 - t - 1) : (A.s0 + t);  // twiddle shift of the first of the two stages
    for (uint32_t qd = tid; qd < tile_elems / 4; qd += NT) {
      const uint32_t g = qd & gmask, hq = qd >> A.g_log;
      const uint32_t low = hq & ((1u << pl) - 1);
      const uint32_t hb0 = ((hq >> pl) << (pl + 2)) | low;
      const uint32_t e00 = (hb0 << A.g_log) | g, da = 1u << (pl + A.g_log), db = da << 1;
      const uint32_t tw0 = (low << A.stride_log) + inner0 + g, tw1 = ((low | (1u << pl)) << A.stride_log) + inner0 + g;
      Fr x00 = lds_fr(sm, e00), x10 = lds_fr(sm, e00 + da), x01 = lds_fr(sm, e00 + db), x11 = lds_fr(sm, e00 + da + db);
      if (DIT) {
        {  // stage t: pairs differ in bit pl, one twiddle for both
          const Fr w = ld_fr(A.tw + 8 * (uint64_t)(tw0 << sh1));
          Fr u = x10 * w;
          x10 = x00 - u;
          x00 = x00 + u;
          u = x11 * w;
          x11 = x01 - u;
          x01 = x01 + u;
        }
        {  // stage t + 1: pairs differ in bit pl + 1
          Fr u = x01 * ld_fr(A.tw + 8 * (uint64_t)(tw0 << (sh1 - 1)));
          x01 = x00 - u;
          x00 = x00 + u;
          u = x11 * ld_fr(A.tw + 8 * (uint64_t)(tw1 << (sh1 - 1)));
          x11 = x10 - u;
          x10 = x10 + u;
        }
      } else {
        {  // stage t: pairs differ in bit pl + 1
          Fr d = x00 - x01;
          x00 = x00 + x01;
          x01 = d * ld_fr(A.tw + 8 * (uint64_t)(tw0 << sh1));
          d = x10 - x11;
          x10 = x10 + x11;
          x11 = d * ld_fr(A.tw + 8 * (uint64_t)(tw1 << sh1));
        }
        {  // stage t + 1: pairs differ in bit pl, one twiddle for both
          const Fr w = ld_fr(A.tw + 8 * (uint64_t)(tw0 << (sh1 + 1)));
          Fr d = x00 - x10;
          x00 = x00 + x10;
          x10 = d * w;
          d = x01 - x11;
          x01 = x01 + x11;
          x11 = d * w;
        }
      }
      sts_fr(sm, e00, x00);
      sts_fr(sm, e00 + da, x10);
      sts_fr(sm, e00 + db, x01);
      sts_fr(sm, e00 + da + db, x11);
    }
  }
#endif
  for (; t < A.S; t++) {
    __syncthreads();
    const uint32_t pb = DIT ? t : (A.S - 1 - t);  // bit of `hi` that distinguishes the pair
    const uint32_t tw_shift = DIT ? (A.L - A.s0 - t - 1) : (A.s0 + t);
    for (uint32_t b = tid; b < tile_elems / 2; b += NT) {
      uint32_t g = b & gmask, hb = b >> A.g_log;
      uint32_t low = hb & ((1u << pb) - 1);
      uint32_t hi0 = ((hb >> pb) << (pb + 1)) | low;
      uint32_t e0 = (hi0 << A.g_log) | g, e1 = e0 + (1u << (pb + A.g_log));
      uint32_t ex = ((low << A.stride_log) + inner0 + g) << tw_shift;
      Fr w = ld_fr(A.tw + 8 * (uint64_t)ex);
      Fr x = lds_fr(sm, e0), y = lds_fr(sm, e1);
      if (DIT) {
        Fr tt = y * w;
        sts_fr(sm, e0, x + tt);
        sts_fr(sm, e1, x - tt);
      } else {
        sts_fr(sm, e0, x + y);
        sts_fr(sm, e1, (x - y) * w);
      }
    }
  }
  __syncthreads();
  const bool last = A.s0 + A.S == A.L;
  for (uint32_t e = tid; e < tile_elems; e += NT) {
    uint32_t gi = base + ((e >> A.g_log) << A.stride_log) + (e & gmask);
    Fr x = lds_fr(sm, e);
    if (!DIT && last && (A.scale || A.scatter_out || A.sub_off)) {
      uint32_t nat = __brev(gi) >> (32 - A.L);
      if (A.sub_off) x = x - ld_fr(data + A.sub_off + 8 * (uint64_t)gi);
      const uint32_t* sc = (A.scale_c && blockIdx.y % 3 == 2) ? A.scale_c : A.scale;
      if (sc) x = x * ld_fr(sc + 8 * (uint64_t)nat);
      if (A.scatter_out) {
        st_fr(A.scatter_out + blockIdx.y * A.scatter_stride + 8 * (uint64_t)nat, x);
        continue;
      }
    }
    st_fr(data + 8 * (uint64_t)gi, x);
  }
}

// a = (a*b - c) * zinv   (mul_polynomials_in_evaluation_domain, -= c, divide_by_vanishing_poly_on_coset)
__global__ void pointwise_kernel(uint32_t* a, const uint32_t* b, const uint32_t* c, const uint32_t* zinv, uint32_t n,
                                 uint64_t batch_stride) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  a += blockIdx.y * batch_stride;
  b += blockIdx.y * batch_stride;
  c += blockIdx.y * batch_stride;
  Fr x = ld_fr(a + 8 * (uint64_t)i) * ld_fr(b + 8 * (uint64_t)i) - ld_fr(c + 8 * (uint64_t)i);
  st_fr(a + 8 * (uint64_t)i, x * ld_fr(zinv));
}
// a = a * b * zinv  (the c term joins in coefficient form, see launch_witness_map)
__global__ void pointwise_ab_kernel(uint32_t* a, const uint32_t* b, const uint32_t* zinv, uint32_t n, uint64_t batch_stride) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  a += blockIdx.y * batch_stride;
  b += blockIdx.y * batch_stride;
  st_fr(a + 8 * (uint64_t)i, ld_fr(a + 8 * (uint64_t)i) * ld_fr(b + 8 * (uint64_t)i) * ld_fr(zinv));
}
__global__ void scale_kernel(uint32_t* a, const uint32_t* table, const uint32_t* cst, uint32_t n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fr x = ld_fr(a + 8 * (uint64_t)i);
  if (table) x = x * ld_fr(table + 8 * (uint64_t)i);
  if (cst) x = x * ld_fr(cst);
  st_fr(a + 8 * (uint64_t)i, x);
}
__global__ void bitrev_permute_kernel(const uint32_t* src, uint32_t* dst, uint32_t L) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (1u << L)) return;
  st_fr(dst + 8 * (uint64_t)(__brev(i) >> (32 - L)), ld_fr(src + 8 * (uint64_t)i));
}
// a[nc + i] = z[i] for i < ni (input-consistency rows of R1CStoQAP::witness_map)
__global__ void copy_instance_kernel(uint32_t* a, const uint32_t* z, uint32_t nc, uint32_t ni, uint64_t a_stride,
                                     uint64_t z_stride) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  a += blockIdx.y * a_stride;
  z += blockIdx.y * z_stride;
  if (i < ni) st_fr(a + 8 * (uint64_t)(nc + i), ld_fr(z + 8 * (uint64_t)i));
}

}  // namespace

static int32_t get_plan(frcs_ctx* ctx, uint32_t L, cudaStream_t st, NttPlan** out) {
  for (auto& p : ctx->plans)
    if (p.L == L) {
      *out = &p;
      return FRCS_OK;
    }
  if (L < 1 || L > 26) {
    frcs_set_error("NTT size out of range");
    return FRCS_E_INVALID_ARG;
  }
  NttPlan p;
  p.L = L;
  const uint32_t n = 1u << L;
  FRCS_CUDA_CHECK(cudaMalloc(&p.consts, 7 * 32));
  FRCS_CUDA_CHECK(cudaMalloc(&p.tw_fwd, (size_t)(n / 2 + 1) * 32));
  FRCS_CUDA_CHECK(cudaMalloc(&p.tw_inv, (size_t)(n / 2 + 1) * 32));
  FRCS_CUDA_CHECK(cudaMalloc(&p.cp, (size_t)n * 32));
  FRCS_CUDA_CHECK(cudaMalloc(&p.cpi, (size_t)n * 32));
  FRCS_CUDA_CHECK(cudaMalloc(&p.cpz, (size_t)n * 32));
  ntt_consts_kernel<<<1, 1, 0, st>>>(p.consts, L);
  unsigned g2 = (n / 2 + 255) / 256, g1 = (n + 255) / 256;
  if (n >= 2) {
    pow_table_kernel<<<g2, 256, 0, st>>>(p.tw_fwd, p.consts, nullptr, n / 2);
    pow_table_kernel<<<g2, 256, 0, st>>>(p.tw_inv, p.consts + 8, nullptr, n / 2);
  }
  pow_table_kernel<<<g1, 256, 0, st>>>(p.cp, p.consts + 32, p.consts + 16, n);   // g^i / n
  pow_table_kernel<<<g1, 256, 0, st>>>(p.cpi, p.consts + 40, p.consts + 16, n);  // g^-i / n
  pow_table_kernel<<<g1, 256, 0, st>>>(p.cpz, p.consts + 32, p.consts + 24, n);  // g^i / (g^n - 1)
  ctx->launches += 6;
  FRCS_CUDA_CHECK(cudaGetLastError());
  ctx->plans.push_back(p);
  *out = &ctx->plans.back();
  return FRCS_OK;
}

// One full transform of `batch` vectors (stride batch_stride words).  dit=false: natural
// in, bit-reversed out; dit=true: bit-reversed in, natural out.
static int32_t run_ntt(frcs_ctx* ctx, const NttPlan& p, uint32_t* data, uint32_t batch, uint64_t batch_stride, bool dit,
                       bool inverse, const uint32_t* scale, uint32_t* scatter_out, cudaStream_t st,
                       uint64_t scatter_stride = 0, NttExtra ex = NttExtra()) {
  if (scatter_stride == 0) scatter_stride = batch_stride;
  const uint32_t L = p.L;
  static bool attr_set = false;
  if (!attr_set) {
    FRCS_CUDA_CHECK(cudaFuncSetAttribute(ntt_chunk_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE * 32));
    FRCS_CUDA_CHECK(cudaFuncSetAttribute(ntt_chunk_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE * 32));
    attr_set = true;
  }
  // chunking: the stride-1 chunk takes min(L, TILE_LOG) stages, the rest is split evenly
  uint32_t s_last = L < TILE_LOG ? L : TILE_LOG;
  uint32_t rest = L - s_last;
  uint32_t n_chunks = rest ? (rest + TILE_LOG - 2) / (TILE_LOG - 1) : 0;  // <= 9 stages each (g_log >= 1)
  std::vector<uint32_t> sizes;
  for (uint32_t i = 0; i < n_chunks; i++) sizes.push_back(rest / n_chunks + (i < rest % n_chunks ? 1 : 0));
  // stage order: DIF = big strides first (rest chunks, then the contiguous chunk); DIT = reverse
  std::vector<ChunkArgs> plan;
  if (!dit) {
    uint32_t s0 = 0;
    for (uint32_t S : sizes) {
      ChunkArgs a{L, s0, S, L - s0 - S, TILE_LOG - S, batch_stride, inverse ? p.tw_inv : p.tw_fwd, nullptr, nullptr, 0};
      plan.push_back(a);
      s0 += S;
    }
    ChunkArgs a{L, s0, s_last, 0, 0, batch_stride, inverse ? p.tw_inv : p.tw_fwd, scale, scatter_out, scatter_stride};
    plan.push_back(a);
  } else {
    ChunkArgs a{L, 0, s_last, 0, 0, batch_stride, inverse ? p.tw_inv : p.tw_fwd, scale, nullptr, 0};
    plan.push_back(a);
    uint32_t s0 = s_last;
    for (uint32_t S : sizes) {
      ChunkArgs b{L, s0, S, s0, TILE_LOG - S, batch_stride, inverse ? p.tw_inv : p.tw_fwd, nullptr, nullptr, 0};
      plan.push_back(b);
      s0 += S;
    }
  }
  for (auto& a : plan) a.ab_only = ex.ab_only;
  if (!dit) {  // the epilogue extras belong to the chunk that ends the transform
    plan.back().scale_c = ex.scale_c;
    plan.back().sub_off = ex.sub_off;
  }
  for (auto& a : plan) {
    uint32_t tile_log = a.S + a.g_log;
    dim3 grid(1u << (L - tile_log), batch);
    if (dit)
      ntt_chunk_kernel<true><<<grid, NT, TILE * 32, st>>>(data, a);
    else
      ntt_chunk_kernel<false><<<grid, NT, TILE * 32, st>>>(data, a);
    ctx->launches++;
  }
  FRCS_CUDA_CHECK(cudaGetLastError());
  return FRCS_OK;
}

int32_t get_ntt_plan(frcs_ctx* ctx, uint32_t L, cudaStream_t st, NttPlan** out) { return get_plan(ctx, L, st, out); }

int32_t ensure_scratch(frcs_ctx* ctx, size_t bytes) {
  if (ctx->scratch_bytes >= bytes) return FRCS_OK;
  if (ctx->scratch) cudaFree(ctx->scratch);
  ctx->scratch = nullptr;
  ctx->scratch_bytes = 0;
  FRCS_CUDA_CHECK(cudaMalloc(&ctx->scratch, bytes));
  ctx->scratch_bytes = bytes;
  return FRCS_OK;
}

// nb assignments z (device, n_z Fr each, consecutive) -> nb vectors h (device, n x 32 B each, natural
// order).  work: nb x 3n Fr of scratch, laid out [problem][a | b | c][n].
int32_t launch_witness_map(frcs_ctx* ctx, uint32_t nb, const uint64_t* d_z, uint64_t* d_h, uint32_t* work,
                           cudaStream_t st) {
  NttPlan* p;
  int32_t rc = get_plan(ctx, ctx->domain_log2, st, &p);
  if (rc) return rc;
  if (nb == 0) return FRCS_OK;
  NvtxRange nvtx("frcs:witness_map");
  const uint32_t L = p->L, n = 1u << L, nc = ctx->L.n_cons, ni = ctx->L.n_inst;
  uint32_t *a = work, *b = work + 8ull * n, *c = work + 16ull * n;
  int ph = prof_begin(ctx, PROF_WITNESS_MAP, st);
  FRCS_CUDA_CHECK(cudaMemsetAsync(work, 0, (size_t)nb * 3ull * n * 32, st));
  rc = launch_r1cs_eval(ctx, nb, d_z, (uint64_t*)a, (uint64_t*)b, (uint64_t*)c, nullptr, st, 3ull * n);
  if (rc) return rc;
  copy_instance_kernel<<<dim3((ni + 255) / 256, nb), 256, 0, st>>>(a, (const uint32_t*)d_z, nc, ni, 24ull * n,
                                                                   8ull * ctx->L.n_z);
  ctx->launches++;
  int pn = prof_begin(ctx, PROF_NTT, st);
  // Six transforms per proof instead of arkworks' seven, same h: R1CStoQAP::witness_map computes
  //   h = coset_ifft((a_g . b_g - c_g) / (g^n - 1)),   x_g = coset_fft(ifft(x)),
  // and coset_ifft is linear with coset_ifft(c_g) = the coefficients of c (degree < n), so
  //   h = (coset_ifft(a_g . b_g) - ifft(c)) / (g^n - 1)
  // for every assignment, satisfying or not: c never goes to the coset.
  // (1) ifft of a, b, c -> bit-reversed coefficients; a, b scaled by g^i / n (the coset shift), c by g^i / (g^n - 1)
  //     over the raw (un-normalised) output, i.e. c_i g^i n / (g^n - 1)
  NttExtra e1;
  e1.scale_c = p->cpz;
  if ((rc = run_ntt(ctx, *p, work, 3 * nb, 8ull * n, false, true, p->cp, nullptr, st, 0, e1))) return rc;
  // (2) coset fft of a and b back to natural order
  NttExtra e2;
  e2.ab_only = 1;
  if ((rc = run_ntt(ctx, *p, work, 2 * nb, 8ull * n, true, false, nullptr, nullptr, st, 0, e2))) return rc;
  // (3) a = a_g . b_g / (g^n - 1)
  pointwise_ab_kernel<<<dim3((n + 255) / 256, nb), 256, 0, st>>>(a, b, p->consts + 24, n, 24ull * n);
  ctx->launches++;
  // (4) coset_ifft: DIF inverse; epilogue (raw - c's scaled coefficient at the same bit-reversed position) * g^-i / n,
  //     un-bit-reversed on the way out
  NttExtra e4;
  e4.sub_off = 16ull * n;
  if ((rc = run_ntt(ctx, *p, a, nb, 24ull * n, false, true, p->cpi, (uint32_t*)d_h, st, 8ull * n, e4))) return rc;
  prof_end(ctx, pn, st);
  prof_end(ctx, ph, st);
  FRCS_CUDA_CHECK(cudaGetLastError());
  return FRCS_OK;
}

// The witness map of ONE proof in two halves, for a proving key split over several GPUs (frcs_prove_split_*): every
// shard evaluates a, b, c from z (0.25 ms), but takes only the vectors in `vec_mask` (bit v = vector v of a | b | c)
// through ifft + coset fft; the other shards' vectors arrive over NVLink before the tail.  work = [a | b | c][n].
int32_t launch_witness_map_head(frcs_ctx* ctx, const uint64_t* d_z, uint32_t* work, uint32_t vec_mask, cudaStream_t st) {
  NttPlan* p;
  int32_t rc = get_plan(ctx, ctx->domain_log2, st, &p);
  if (rc) return rc;
  NvtxRange nvtx("frcs:witness_map_head");
  const uint32_t L = p->L, n = 1u << L, nc = ctx->L.n_cons, ni = ctx->L.n_inst;
  int ph = prof_begin(ctx, PROF_WITNESS_MAP, st);
  FRCS_CUDA_CHECK(cudaMemsetAsync(work, 0, 3ull * n * 32, st));
  rc = launch_r1cs_eval(ctx, 1, d_z, (uint64_t*)work, (uint64_t*)(work + 8ull * n), (uint64_t*)(work + 16ull * n),
                        nullptr, st, 3ull * n);
  if (rc) return rc;
  copy_instance_kernel<<<dim3((ni + 255) / 256, 1), 256, 0, st>>>(work, (const uint32_t*)d_z, nc, ni, 24ull * n,
                                                                  8ull * ctx->L.n_z);
  ctx->launches++;
  for (uint32_t v = 0; v < 3; v++) {
    if (!(vec_mask >> v & 1)) continue;
    uint32_t* x = work + 8ull * n * v;
    if ((rc = run_ntt(ctx, *p, x, 1, 8ull * n, false, true, p->cp, nullptr, st))) return rc;
    if ((rc = run_ntt(ctx, *p, x, 1, 8ull * n, true, false, nullptr, nullptr, st))) return rc;
  }
  prof_end(ctx, ph, st);
  FRCS_CUDA_CHECK(cudaGetLastError());
  return FRCS_OK;
}

// second half: h = coset_ifft((a . b - c) / Z) from the three coset evaluation vectors in work (a is overwritten)
int32_t launch_witness_map_tail(frcs_ctx* ctx, uint32_t* work, uint64_t* d_h, cudaStream_t st) {
  NttPlan* p;
  int32_t rc = get_plan(ctx, ctx->domain_log2, st, &p);
  if (rc) return rc;
  NvtxRange nvtx("frcs:witness_map_tail");
  const uint32_t n = 1u << p->L;
  int pn = prof_begin(ctx, PROF_NTT, st);
  pointwise_kernel<<<dim3((n + 255) / 256, 1), 256, 0, st>>>(work, work + 8ull * n, work + 16ull * n, p->consts + 24, n,
                                                             24ull * n);
  ctx->launches++;
  if ((rc = run_ntt(ctx, *p, work, 1, 24ull * n, false, true, p->cpi, (uint32_t*)d_h, st, 8ull * n))) return rc;
  prof_end(ctx, pn, st);
  FRCS_CUDA_CHECK(cudaGetLastError());
  return FRCS_OK;
}

extern "C" {

int32_t frcs_witness_map_dev(frcs_ctx* ctx, const uint64_t* d_z, uint64_t* d_h, void* stream) {
  if (!ctx || !d_z || !d_h) return FRCS_E_INVALID_ARG;
  FRCS_CUDA_CHECK(cudaSetDevice(ctx->device));
  int32_t rc = ensure_scratch(ctx, 3ull * 32 << ctx->domain_log2);
  if (rc) return rc;
  return launch_witness_map(ctx, 1, d_z, d_h, (uint32_t*)ctx->scratch, (cudaStream_t)stream);
}

int32_t frcs_witness_map(frcs_ctx* ctx, const uint64_t* z, uint64_t* h_out) {
  if (!ctx || !z || !h_out) return FRCS_E_INVALID_ARG;
  FRCS_CUDA_CHECK(cudaSetDevice(ctx->device));
  const size_t zb = (size_t)ctx->L.n_z * 32, hb = (size_t)32 << ctx->domain_log2;
  void *d_z = nullptr, *d_h = nullptr;
  FRCS_CUDA_CHECK(cudaMalloc(&d_z, zb));
  FRCS_CUDA_CHECK(cudaMalloc(&d_h, hb));
  cudaStream_t st = ctx->stream;
  FRCS_CUDA_CHECK(cudaMemcpyAsync(d_z, z, zb, cudaMemcpyHostToDevice, st));
  int32_t rc = frcs_witness_map_dev(ctx, (const uint64_t*)d_z, (uint64_t*)d_h, st);
  if (rc == FRCS_OK) {
    FRCS_CUDA_CHECK(cudaMemcpyAsync(h_out, d_h, hb, cudaMemcpyDeviceToHost, st));
    FRCS_CUDA_CHECK(cudaStreamSynchronize(st));
  }
  cudaFree(d_z);
  cudaFree(d_h);
  return rc;
}

// op: 0 fft, 1 ifft, 2 coset_fft, 3 coset_ifft (natural order in and out, like ark-poly);
//     4 = DIF forward then DIT inverse (round trip through both kernels, scaled by 1/n)
int32_t frcs_domain_op(frcs_ctx* ctx, uint32_t log_size, int32_t op, uint64_t* data) {
  if (!ctx || !data || op < 0 || op > 4) return FRCS_E_INVALID_ARG;
  FRCS_CUDA_CHECK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  NttPlan* p;
  int32_t rc = get_plan(ctx, log_size, st, &p);
  if (rc) return rc;
  const uint32_t n = 1u << log_size;
  uint32_t *d = nullptr, *tmp = nullptr;
  FRCS_CUDA_CHECK(cudaMalloc(&d, (size_t)n * 32));
  FRCS_CUDA_CHECK(cudaMalloc(&tmp, (size_t)n * 32));
  FRCS_CUDA_CHECK(cudaMemcpyAsync(d, data, (size_t)n * 32, cudaMemcpyHostToDevice, st));
  unsigned g = (n + 255) / 256;
  uint32_t* result = d;
  if (op == 0 || op == 2) {
    if (op == 2) {  // distribute powers of g: the table holds g^i/n, so multiply by n afterwards
      scale_kernel<<<g, 256, 0, st>>>(d, p->cp, p->consts + 48, n);
      rc = run_ntt(ctx, *p, d, 1, 8ull * n, false, false, nullptr, nullptr, st);
      bitrev_permute_kernel<<<g, 256, 0, st>>>(d, tmp, log_size);
      result = tmp;
    } else {
      rc = run_ntt(ctx, *p, d, 1, 8ull * n, false, false, nullptr, nullptr, st);
      bitrev_permute_kernel<<<g, 256, 0, st>>>(d, tmp, log_size);
      result = tmp;
    }
  } else if (op == 1 || op == 3) {
    rc = run_ntt(ctx, *p, d, 1, 8ull * n, false, true, nullptr, nullptr, st);
    bitrev_permute_kernel<<<g, 256, 0, st>>>(d, tmp, log_size);
    if (op == 3)
      scale_kernel<<<g, 256, 0, st>>>(tmp, p->cpi, nullptr, n);  // g^-i / n
    else
      scale_kernel<<<g, 256, 0, st>>>(tmp, nullptr, p->consts + 16, n);  // 1/n
    result = tmp;
  } else {
    rc = run_ntt(ctx, *p, d, 1, 8ull * n, false, false, nullptr, nullptr, st);
    if (!rc) rc = run_ntt(ctx, *p, d, 1, 8ull * n, true, true, nullptr, nullptr, st);
    scale_kernel<<<g, 256, 0, st>>>(d, nullptr, p->consts + 16, n);
  }
  if (rc == FRCS_OK) {
    FRCS_CUDA_CHECK(cudaMemcpyAsync(data, result, (size_t)n * 32, cudaMemcpyDeviceToHost, st));
    FRCS_CUDA_CHECK(cudaStreamSynchronize(st));
    FRCS_CUDA_CHECK(cudaGetLastError());
  }
  cudaFree(d);
  cudaFree(tmp);
  return rc;
}

}  // extern "C"
