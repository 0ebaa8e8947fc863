// Subsystem (1): batched witness generation for FalconNTTVerificationCircuit.
// One signature per thread block.  Replaces `generate_constraints` in Prove mode
// (circuits/falcon_ntt.rs:26-123 and the value closures of every gadget it calls):
// the clear-text prelude (falcon_ntt.rs:44-51), ntt_circuit's unreduced butterflies and
// mod_q quotients (gadgets/poly.rs:104-159, gadgets/arithmetics.rs:105-149), add_mod
// (arithmetics.rs:214-262), the range-proof bits (gadgets/range_proofs.rs) and
// l2_norm_var (gadgets/misc.rs:30-51).  Output: the full assignment
// z = [1 | pk_ntt | hm_ntt | witnesses] in allocation order,
// every entry an Fr in Montgomery form, written with 32-byte stores.
//
// HBM-write bound: 32*(n_inst+n_wit) bytes per signature (5,080,736 B for Falcon-1024).
#include <algorithm>

#include "ctx.hpp"
#include "nvtx.hpp"
#define FF_INLINE_MUL
#include "ff32.cuh"
#include "witness_dev.cuh"

using ff::Fr;
using namespace wdev;

namespace {

constexpr int WT = 512;  // threads per block

struct WitnessParams {
  circuit::Layout L;
  NormOpsDev ops;
  uint32_t cst[11][5];  // 2^(l+1) q^(l+2) for layer l (falcon_ntt.rs:31-39)
  uint32_t n_inv;       // N^-1 mod q
};

// Coalesced sweep over records [c0, c0 + ch) of one section, STRIDE entries per record: bits [0, NBITS) of the
// record's mask at positions BIT_LO.., and NVAL non-boolean entries at positions VAL_LO.. whose Montgomery
// values were staged in shared memory (slot = record * NVAL + k).  Every entry of the chunk is written here,
// in address order: one warp = 1 KiB of consecutive bytes.
template <int STRIDE, int BIT_LO, int NBITS, int VAL_LO, int NVAL, class MaskFn>
__device__ __forceinline__ void sweep_chunk(uint64_t* zc, uint32_t c0, uint32_t ch, int tid, const uint32_t* stage,
                                            MaskFn mask_of) {
  for (uint32_t r = tid; r < ch * STRIDE; r += WT) {
    const uint32_t rr = r / STRIDE, pos = r - rr * STRIDE;
    const uint32_t jb = pos - BIT_LO, jv = pos - VAL_LO;  // unsigned wrap when below the range
    if (NVAL > 0 && jv < (uint32_t)NVAL) {
      const uint32_t* src = stage + (rr * NVAL + jv) * 8;
      Fr v;
#pragma unroll
      for (int k = 0; k < 8; k++) v.v[k] = src[k];
      store_fr(zc + 4 * (uint64_t)r, v);
    } else if (jb < (uint32_t)NBITS) {
      store_bit(zc + 4 * (uint64_t)r, (mask_of(c0 + rr) >> jb) & 1u);
    }
  }
}

template <int LOGN>
__global__ void __launch_bounds__(WT, 2)
    witness_kernel(WitnessParams P, uint64_t n_sig, const uint16_t* __restrict__ g_sig, const uint16_t* __restrict__ g_pk,
                   const uint16_t* __restrict__ g_hm, const uint32_t* __restrict__ g_tab,
                   const uint32_t* __restrict__ g_mont, uint32_t* __restrict__ g_tq, uint64_t* __restrict__ g_z,
                   int32_t* __restrict__ g_status) {
  constexpr int N = 1 << LOGN;
  extern __shared__ uint32_t smem[];
  uint32_t* tq = g_tq + (size_t)blockIdx.x * 2 * N * 8;  // this CTA's mod_q quotients (Montgomery), [2][N]
  uint32_t* s_tab = smem;           // [N] forward twiddles
  uint32_t* s_itab = s_tab + N;     // [N] inverse twiddles
  uint32_t* s_sig = s_itab + N;     // sig (over Z_q)
  uint32_t* s_v = s_sig + N;        // v = hm - sig*pk
  uint32_t* s_pkn = s_v + N;        // pk_ntt
  uint32_t* s_hmn = s_pkn + N;      // hm_ntt
  uint32_t* s_sign = s_hmn + N;     // sig_ntt (= mod_q outputs b)
  uint32_t* s_vn = s_sign + N;      // v_ntt
  uint32_t* s_pwp = s_vn + N;       // sig_ntt*pk_ntt
  uint32_t* s_pwt = s_pwp + N;      // add_mod quotient
  uint32_t* s_pwc = s_pwt + N;      // add_mod remainder
  uint32_t* s_l2s = s_pwc + N;      // [2N] lifted |e|
  uint32_t* s_l2p = s_l2s + 2 * N;  // [2N] squares
  uint32_t* s_lazy = s_l2p + 2 * N; // [5][N] unreduced NTT values
  uint32_t* s_norm = s_lazy + 5 * N;  // [64] norm gadget witnesses
  uint32_t* s_stage = s_norm + 64;    // [384][8] Montgomery values of the chunk being written
  __shared__ unsigned long long s_acc;
  __shared__ int s_bad;

  const int tid = threadIdx.x;
  const circuit::Layout& L = P.L;

  for (int i = tid; i < N; i += WT) {
    s_tab[i] = g_tab[i];
    s_itab[i] = g_tab[N + i];
  }

  for (uint64_t sid = blockIdx.x; sid < n_sig; sid += gridDim.x) {
    __syncthreads();
    if (tid == 0) {
      s_acc = 0;
      s_bad = 0;
    }
    uint64_t* z = g_z + sid * (uint64_t)L.n_z * 4;
    // ---- load inputs ----
    int bad = 0;
    for (int i = tid; i < N; i += WT) {
      uint32_t a = g_sig[sid * N + i], b = g_pk[sid * N + i], c = g_hm[sid * N + i];
      bad |= (a >= Q) | (b >= Q) | (c >= Q);
      s_sig[i] = modq(a);
      s_pkn[i] = modq(b);
      s_hmn[i] = modq(c);
      s_sign[i] = modq(a);
    }
    __syncthreads();
    if (bad) s_bad = 1;
    // ---- clear-text NTTs of pk, hm, sig (NTTPolynomial::from, falcon_ntt.rs:45-51) ----
    {
      int t = N;
#pragma unroll 1
      for (int l = 0; l < LOGN; l++) {
        int ht = t >> 1;
        for (int idx = tid; idx < N / 2; idx += WT) {
          int i = idx / ht, j = idx - i * ht;
          int p0 = i * t + j, p1 = p0 + ht;
          uint32_t s = s_tab[(1 << l) + i];
          uint32_t u, v;
          u = s_pkn[p0]; v = modq(s_pkn[p1] * s); s_pkn[p0] = modq(u + v); s_pkn[p1] = modq(u + Q - v);
          u = s_hmn[p0]; v = modq(s_hmn[p1] * s); s_hmn[p0] = modq(u + v); s_hmn[p1] = modq(u + Q - v);
          u = s_sign[p0]; v = modq(s_sign[p1] * s); s_sign[p0] = modq(u + v); s_sign[p1] = modq(u + Q - v);
        }
        t = ht;
        __syncthreads();
      }
    }
    // v_ntt = hm_ntt - sig_ntt * pk_ntt; v = INTT(v_ntt)   (v = hm - sig*pk, falcon_ntt.rs:47-49)
    for (int i = tid; i < N; i += WT) {
      uint32_t x = modq(s_hmn[i] + Q - modq(s_sign[i] * s_pkn[i]));
      s_vn[i] = x;
      s_v[i] = x;
    }
    __syncthreads();
    {
      int t = 1;
#pragma unroll 1
      for (int l = LOGN - 1; l >= 0; l--) {
        for (int idx = tid; idx < N / 2; idx += WT) {
          int i = idx / t, j = idx - i * t;
          int p0 = i * 2 * t + j, p1 = p0 + t;
          uint32_t si = s_itab[(1 << l) + i];
          uint32_t u = s_v[p0], v = s_v[p1];
          s_v[p0] = modq(u + v);
          s_v[p1] = modq((u + Q - v) * si);
        }
        t <<= 1;
        __syncthreads();
      }
      for (int i = tid; i < N; i += WT) s_v[i] = modq(s_v[i] * P.n_inv);
      __syncthreads();
    }
    // ---- ntt_circuit: unreduced butterflies on integers (poly.rs:115-149) + mod_q quotients ----
#pragma unroll 1
    for (int pass = 0; pass < 2; pass++) {
      const uint32_t* src = pass == 0 ? s_sig : s_v;

      for (int i = tid; i < N; i += WT) {
        s_lazy[i] = src[i];
#pragma unroll
        for (int k = 1; k < 5; k++) s_lazy[k * N + i] = 0;
      }
      __syncthreads();
      int t = N;
#pragma unroll 1
      for (int l = 0; l < LOGN; l++) {
        int ht = t >> 1;
        for (int idx = tid; idx < N / 2; idx += WT) {
          int i = idx / ht, j = idx - i * ht;
          int p0 = i * t + j, p1 = p0 + ht;
          uint32_t s = s_tab[(1 << l) + i];
          uint32_t u[5], sv[5];
          uint64_t c = 0;
#pragma unroll
          for (int k = 0; k < 5; k++) {
            u[k] = s_lazy[k * N + p0];
            c += (uint64_t)s_lazy[k * N + p1] * s;
            sv[k] = (uint32_t)c;
            c >>= 32;
          }
          // out[j] = u + v ; out[j+ht] = u + (const[l+1] - v)
          uint64_t ca = 0, cb = 0;
          int64_t br = 0;
#pragma unroll
          for (int k = 0; k < 5; k++) {
            ca += (uint64_t)u[k] + sv[k];
            s_lazy[k * N + p0] = (uint32_t)ca;
            ca >>= 32;
            int64_t d = (int64_t)P.cst[l][k] - sv[k] - br;
            br = d < 0;
            cb += (uint64_t)u[k] + (uint32_t)d;
            s_lazy[k * N + p1] = (uint32_t)cb;
            cb >>= 32;
          }
        }
        t = ht;
        __syncthreads();
      }
      // mod_q: t = a / q, b = a % q on the canonical integer (arithmetics.rs:127-134); the quotient (up to
      // 146 bits) is parked in a per-CTA scratch (L2) in Montgomery form until its chunk of z is written
      for (int i = tid; i < N; i += WT) {
        uint64_t rem = 0;
        Fr t = Fr::zero();
#pragma unroll
        for (int k = 4; k >= 0; k--) {
          uint64_t cur = (rem << 32) | s_lazy[k * N + i];
          uint64_t qk = cur / Q;
          rem = cur - qk * Q;
          t.v[k] = (uint32_t)qk;
        }
        store_fr(reinterpret_cast<uint64_t*>(tq + (size_t)(pass * N + i) * 8), t.to_mont());
        // rem == clear-text NTT value; keep the one derived from the wide value
        if (pass == 0)
          s_sign[i] = (uint32_t)rem;
        else
          s_vn[i] = (uint32_t)rem;
      }
      __syncthreads();
    }
    // ---- pointwise products and add_mod (falcon_ntt.rs:94-111, arithmetics.rs:239-246) ----
    unsigned long long local_norm = 0;
    for (int i = tid; i < N; i += WT) {
      uint32_t p = s_sign[i] * s_pkn[i];
      uint32_t sum = s_vn[i] + p;
      s_pwp[i] = p;
      s_pwt[i] = sum / Q;
      s_pwc[i] = sum % Q;
    }
    // ---- l2_norm_var over v ++ sig (misc.rs:30-51) ----
    for (int k = tid; k < 2 * N; k += WT) {
      uint32_t e = k < N ? s_v[k] : s_sig[k - N];
      uint32_t s = e < 6144 ? e : Q - e;
      s_l2s[k] = s;
      s_l2p[k] = s * s;
      local_norm += (unsigned long long)s * s;
    }
    for (int o = 16; o > 0; o >>= 1) local_norm += __shfl_xor_sync(0xffffffffu, local_norm, o);
    if ((tid & 31) == 0) atomicAdd(&s_acc, local_norm);
    __syncthreads();
    // ---- enforce_less_than_norm_bound witnesses (range_proofs.rs:100-186 / 192-272) ----
    if (tid == 0) {
      unsigned long long norm = s_acc;
      int st = 0;
      if (s_bad) st = FRCS_E_COEFF_RANGE;
      if (st == 0 && norm >= L.l2_bound) st = FRCS_E_NORM_BOUND;
      g_status[sid] = st;
      for (uint32_t i = 0; i < L.norm_bits; i++) s_norm[i] = (uint32_t)((norm >> i) & 1);
      for (uint32_t k = 0; k < L.norm_ops; k++) {
        uint32_t a = s_norm[P.ops.a[k]], b = s_norm[P.ops.b[k]], r;
        switch (P.ops.kind[k]) {
          case circuit::OP_AND: r = a & b; break;
          case circuit::OP_OR: r = a | b; break;
          case circuit::OP_AND_NOT: r = a & (b ^ 1); break;
          default: r = (a ^ 1) & (b ^ 1); break;
        }
        s_norm[L.norm_bits + k] = r;
      }
    }
    __syncthreads();

    // ================= output =================
    // z is written front to back in chunks of 128 gadget records (~120 KB): within a chunk first the few
    // non-boolean entries (strided 32-byte stores), then the boolean entries as a coalesced sweep (one warp =
    // 1 KiB).  Both land in the same lines within microseconds, so L2 hands HBM whole lines in address order;
    // writing all strided entries of a signature first costs 30 % (partially written lines get evicted).
    // (0) One, pk_ntt, hm_ntt, sig, v: contiguous values
    for (int d = tid; d < 4 * N + 1; d += WT) {
      if (d == 4 * N) {
        store_bit(z, true);  // z[0] = One
        continue;
      }
      int g = d >> LOGN, i = d & (N - 1);
      uint32_t x = g == 0 ? s_pkn[i] : g == 1 ? s_hmn[i] : g == 2 ? s_sig[i] : s_v[i];
      uint32_t pos = g == 0 ? 1 + i : g == 1 ? 1 + N + i : g == 2 ? L.n_inst + L.w_sig + i : L.n_inst + L.w_v + i;
      store_fr(z + 4 * (uint64_t)pos, mont_small(g_mont, x));
    }
    constexpr uint32_t CH = 128;  // records per chunk
#pragma unroll 1
    for (int sec = 0; sec < 5; sec++) {
      // section: first witness, record stride, records, value entries per record
      const uint32_t w0 = sec == 0 ? L.w_vrange : sec == 1 ? L.w_nttsig : sec == 2 ? L.w_nttv : sec == 3 ? L.w_pw : L.w_l2;
      const uint32_t stride = sec == 0 ? 27 : sec <= 2 ? 29 : sec == 3 ? 30 : 18;
      const uint32_t nrec = sec == 4 ? 2 * N : N, nval = sec == 0 ? 0 : sec == 3 ? 3 : 2;
#pragma unroll 1
      for (uint32_t c0 = 0; c0 < nrec; c0 += CH) {
        uint64_t* zc = z + 4 * (uint64_t)(L.n_inst + w0 + c0 * stride);
        // the chunk's non-boolean entries (at most one per thread: CH * nval <= 384 < WT) are brought to Montgomery
        // form and staged in shared memory; the sweep then writes every entry of the chunk in address order
        if ((uint32_t)tid < CH * nval) {
          const uint32_t rr = (uint32_t)tid / nval, j = (uint32_t)tid - rr * nval, i = c0 + rr;
          Fr v;
          if (sec <= 2 && j == 0) {
            v = ld_tab(tq + (size_t)((sec - 1) * N + i) * 8);  // t: the parked Montgomery value
          } else {
            uint32_t x;
            if (sec <= 2)
              x = sec == 1 ? s_sign[i] : s_vn[i];
            else if (sec == 3)
              x = j == 0 ? s_pwp[i] : j == 1 ? s_pwt[i] : s_pwc[i];
            else
              x = j == 0 ? s_l2s[i] : s_l2p[i];
            v = mont_small(g_mont, x);
          }
#pragma unroll
          for (int k = 0; k < 8; k++) s_stage[tid * 8 + k] = v.v[k];
        }
        __syncthreads();
        if (sec == 0)
          sweep_chunk<27, 0, 27, 0, 0>(zc, c0, CH, tid, s_stage, [&](uint32_t i) { return ltq_mask(s_v[i]); });
        else if (sec == 1)
          sweep_chunk<29, 2, 27, 0, 2>(zc, c0, CH, tid, s_stage, [&](uint32_t i) { return ltq_mask(s_sign[i]); });
        else if (sec == 2)
          sweep_chunk<29, 2, 27, 0, 2>(zc, c0, CH, tid, s_stage, [&](uint32_t i) { return ltq_mask(s_vn[i]); });
        else if (sec == 3)
          sweep_chunk<30, 3, 27, 0, 3>(zc, c0, CH, tid, s_stage, [&](uint32_t i) { return ltq_mask(s_pwc[i]); });
        else
          sweep_chunk<18, 0, 16, 16, 2>(zc, c0, CH, tid, s_stage,
                                        [&](uint32_t i) { return l2_mask(i < (uint32_t)N ? s_v[i] : s_sig[i - N]); });
        __syncthreads();
      }
    }
    for (uint32_t w = tid; w < L.norm_bits + L.norm_ops; w += WT)
      store_bit(z + 4 * (uint64_t)(L.n_inst + L.w_norm + w), s_norm[w]);
  }
}

// ---------------------------------------------------------------------------------------------
// FalconDualNTTVerificationCircuit (circuits/falcon_dual_ntt.rs:26-132, gadgets/dual_poly.rs:8-52): the signature
// and v = hm - sig * pk are split into (pos, neg) pairs of polynomials with coefficients below 6144
// (DualPolynomial::from, [EXT] falcon-rust: e < 6144 -> pos = e, else neg = q - e), each half goes through
// ntt_circuit, and per index  mod_q(hm_ntt + v_neg_ntt + sig_neg_ntt * pk_ntt) == mod_q(v_pos_ntt + sig_pos_ntt * pk_ntt).
// z = [1 | pk_ntt | hm_ntt | sig pair | v pair | 4 x ntt_circuit | N x (left, right) | 4N squares | norm];
// a pair = pos[N], neg[N], the N products pos_i * neg_i (all zero), then `ne` = 0 and `multiplier` = 1 of is_zero.
// Same structure as witness_kernel: clear-text prelude, lazy butterflies, then z front to back in chunks.
template <int LOGN>
__global__ void __launch_bounds__(WT, 2)
    witness_dual_kernel(WitnessParams P, uint64_t n_sig, const uint16_t* __restrict__ g_sig,
                        const uint16_t* __restrict__ g_pk, const uint16_t* __restrict__ g_hm,
                        const uint32_t* __restrict__ g_tab, const uint32_t* __restrict__ g_mont,
                        uint32_t* __restrict__ g_tq, uint64_t* __restrict__ g_z, int32_t* __restrict__ g_status) {
  constexpr int N = 1 << LOGN;
  extern __shared__ uint32_t smem[];
  uint32_t* tq = g_tq + (size_t)blockIdx.x * 4 * N * 8;  // this CTA's mod_q quotients (Montgomery), [4][N]
  uint32_t* s_tab = smem;             // [N] forward twiddles
  uint32_t* s_itab = s_tab + N;       // [N] inverse twiddles
  uint32_t* s_pkn = s_itab + N;       // pk_ntt
  uint32_t* s_hmn = s_pkn + N;        // hm_ntt
  uint32_t* s_x = s_hmn + N;          // [4][N] sig.pos, sig.neg, v.pos, v.neg
  uint32_t* s_y = s_x + 4 * N;        // [4][N] their NTTs (the mod_q remainders)
  uint32_t* s_lazy = s_y + 4 * N;     // [5][N] unreduced NTT values; the prelude parks sig, sig_ntt, v here
  uint32_t* s_norm = s_lazy + 5 * N;  // [64] norm gadget witnesses
  uint32_t* s_stage = s_norm + 64;    // [384][8] Montgomery values of the chunk being written
  uint32_t *t_sig = s_lazy, *t_sign = s_lazy + N, *t_v = s_lazy + 2 * N;
  __shared__ unsigned long long s_acc;
  __shared__ int s_bad;

  const int tid = threadIdx.x;
  const circuit::Layout& L = P.L;
  for (int i = tid; i < N; i += WT) {
    s_tab[i] = g_tab[i];
    s_itab[i] = g_tab[N + i];
  }
  for (uint64_t sid = blockIdx.x; sid < n_sig; sid += gridDim.x) {
    __syncthreads();
    if (tid == 0) {
      s_acc = 0;
      s_bad = 0;
    }
    uint64_t* z = g_z + sid * (uint64_t)L.n_z * 4;
    int bad = 0;
    for (int i = tid; i < N; i += WT) {
      uint32_t a = g_sig[sid * N + i], b = g_pk[sid * N + i], c = g_hm[sid * N + i];
      bad |= (a >= Q) | (b >= Q) | (c >= Q);
      t_sig[i] = modq(a);
      t_sign[i] = modq(a);
      s_pkn[i] = modq(b);
      s_hmn[i] = modq(c);
    }
    __syncthreads();
    if (bad) s_bad = 1;
    // ---- clear-text NTTs of pk, hm, sig; v = INTT(hm_ntt - sig_ntt * pk_ntt)   (falcon_dual_ntt.rs:44-53) ----
    {
      int t = N;
#pragma unroll 1
      for (int l = 0; l < LOGN; l++) {
        int ht = t >> 1;
        for (int idx = tid; idx < N / 2; idx += WT) {
          int i = idx / ht, j = idx - i * ht;
          int p0 = i * t + j, p1 = p0 + ht;
          uint32_t s = s_tab[(1 << l) + i];
          uint32_t u, v;
          u = s_pkn[p0]; v = modq(s_pkn[p1] * s); s_pkn[p0] = modq(u + v); s_pkn[p1] = modq(u + Q - v);
          u = s_hmn[p0]; v = modq(s_hmn[p1] * s); s_hmn[p0] = modq(u + v); s_hmn[p1] = modq(u + Q - v);
          u = t_sign[p0]; v = modq(t_sign[p1] * s); t_sign[p0] = modq(u + v); t_sign[p1] = modq(u + Q - v);
        }
        t = ht;
        __syncthreads();
      }
    }
    for (int i = tid; i < N; i += WT) t_v[i] = modq(s_hmn[i] + Q - modq(t_sign[i] * s_pkn[i]));
    __syncthreads();
    {
      int t = 1;
#pragma unroll 1
      for (int l = LOGN - 1; l >= 0; l--) {
        for (int idx = tid; idx < N / 2; idx += WT) {
          int i = idx / t, j = idx - i * t;
          int p0 = i * 2 * t + j, p1 = p0 + t;
          uint32_t si = s_itab[(1 << l) + i];
          uint32_t u = t_v[p0], v = t_v[p1];
          t_v[p0] = modq(u + v);
          t_v[p1] = modq((u + Q - v) * si);
        }
        t <<= 1;
        __syncthreads();
      }
    }
    // ---- DualPolynomial::from: split sig and v; the norm over the four halves (misc.rs:55-65) ----
    unsigned long long local_norm = 0;
    for (int i = tid; i < N; i += WT) {
      const uint32_t e = t_sig[i], w = modq(t_v[i] * P.n_inv);
      const uint32_t ep = e < 6144 ? e : 0, en = e < 6144 ? 0 : Q - e;
      const uint32_t wp = w < 6144 ? w : 0, wn = w < 6144 ? 0 : Q - w;
      s_x[i] = ep;
      s_x[N + i] = en;
      s_x[2 * N + i] = wp;
      s_x[3 * N + i] = wn;
      local_norm += (unsigned long long)ep * ep + (unsigned long long)en * en + (unsigned long long)wp * wp +
                    (unsigned long long)wn * wn;
    }
    for (int o = 16; o > 0; o >>= 1) local_norm += __shfl_xor_sync(0xffffffffu, local_norm, o);
    if ((tid & 31) == 0) atomicAdd(&s_acc, local_norm);
    __syncthreads();
    // ---- 4 x ntt_circuit: unreduced butterflies on integers (poly.rs:115-149) + mod_q quotients ----
#pragma unroll 1
    for (int pass = 0; pass < 4; pass++) {
      const uint32_t* src = s_x + pass * N;
      for (int i = tid; i < N; i += WT) {
        s_lazy[i] = src[i];
#pragma unroll
        for (int k = 1; k < 5; k++) s_lazy[k * N + i] = 0;
      }
      __syncthreads();
      int t = N;
#pragma unroll 1
      for (int l = 0; l < LOGN; l++) {
        int ht = t >> 1;
        for (int idx = tid; idx < N / 2; idx += WT) {
          int i = idx / ht, j = idx - i * ht;
          int p0 = i * t + j, p1 = p0 + ht;
          uint32_t s = s_tab[(1 << l) + i];
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
          for (int k = 0; k < 5; k++) {
            ca += (uint64_t)u[k] + sv[k];
            s_lazy[k * N + p0] = (uint32_t)ca;
            ca >>= 32;
            int64_t d = (int64_t)P.cst[l][k] - sv[k] - br;
            br = d < 0;
            cb += (uint64_t)u[k] + (uint32_t)d;
            s_lazy[k * N + p1] = (uint32_t)cb;
            cb >>= 32;
          }
        }
        t = ht;
        __syncthreads();
      }
      for (int i = tid; i < N; i += WT) {
        uint64_t rem = 0;
        Fr t = Fr::zero();
#pragma unroll
        for (int k = 4; k >= 0; k--) {
          uint64_t cur = (rem << 32) | s_lazy[k * N + i];
          uint64_t qk = cur / Q;
          rem = cur - qk * Q;
          t.v[k] = (uint32_t)qk;
        }
        store_fr(reinterpret_cast<uint64_t*>(tq + (size_t)(pass * N + i) * 8), t.to_mont());
        s_y[pass * N + i] = (uint32_t)rem;
      }
      __syncthreads();
    }
    // ---- enforce_less_than_norm_bound witnesses ----
    if (tid == 0) {
      unsigned long long norm = s_acc;
      int st = 0;
      if (s_bad) st = FRCS_E_COEFF_RANGE;
      if (st == 0 && norm >= L.l2_bound) st = FRCS_E_NORM_BOUND;
      g_status[sid] = st;
      for (uint32_t i = 0; i < L.norm_bits; i++) s_norm[i] = (uint32_t)((norm >> i) & 1);
      for (uint32_t k = 0; k < L.norm_ops; k++) {
        uint32_t a = s_norm[P.ops.a[k]], b = s_norm[P.ops.b[k]], r;
        switch (P.ops.kind[k]) {
          case circuit::OP_AND: r = a & b; break;
          case circuit::OP_OR: r = a | b; break;
          case circuit::OP_AND_NOT: r = a & (b ^ 1); break;
          default: r = (a ^ 1) & (b ^ 1); break;
        }
        s_norm[L.norm_bits + k] = r;
      }
    }
    __syncthreads();

    // ================= output =================
    // (0) One, pk_ntt, hm_ntt; the two (pos, neg, products, ne, multiplier) pairs: contiguous values
    for (int d = tid; d < 2 * N + 1; d += WT) {
      if (d == 2 * N) {
        store_bit(z, true);
        continue;
      }
      store_fr(z + 4 * (uint64_t)(1 + d), mont_small(g_mont, d < N ? s_pkn[d] : s_hmn[d - N]));
    }
    for (int d = tid; d < 2 * (3 * N + 2); d += WT) {
      const int pair = d >= 3 * N + 2, j = d - pair * (3 * N + 2);
      uint64_t* dst = z + 4 * (uint64_t)(L.n_inst + (pair ? L.w_v : L.w_sig) + j);
      if (j < 2 * N)
        store_fr(dst, mont_small(g_mont, s_x[pair * 2 * N + j]));
      else
        store_bit(dst, j == 3 * N + 1);  // products and `ne` are zero, `multiplier` is one
    }
    constexpr uint32_t CH = 128;  // records per chunk
    // (1) the four ntt_circuit blocks: N x [t, b, 27 range witnesses of b]
#pragma unroll 1
    for (int blk = 0; blk < 4; blk++) {
      const uint32_t* yb = s_y + blk * N;
#pragma unroll 1
      for (uint32_t c0 = 0; c0 < (uint32_t)N; c0 += CH) {
        uint64_t* zc = z + 4 * (uint64_t)(L.n_inst + L.w_ntt4 + 29 * N * blk + c0 * 29);
        if ((uint32_t)tid < CH * 2) {
          const uint32_t rr = (uint32_t)tid >> 1, j = (uint32_t)tid & 1, i = c0 + rr;
          Fr v = j == 0 ? ld_tab(tq + (size_t)(blk * N + i) * 8) : mont_small(g_mont, yb[i]);
#pragma unroll
          for (int k = 0; k < 8; k++) s_stage[tid * 8 + k] = v.v[k];
        }
        __syncthreads();
        sweep_chunk<29, 2, 27, 0, 2>(zc, c0, CH, tid, s_stage, [&](uint32_t i) { return ltq_mask(yb[i]); });
        __syncthreads();
      }
    }
    // (2) pointwise: 2N records [product, t, b, 27 range witnesses of b]; record 2i = left, 2i + 1 = right of index i
    auto pw = [&](uint32_t r, uint32_t& p, uint32_t& t, uint32_t& b) {
      const uint32_t i = r >> 1;
      uint32_t sum;
      if ((r & 1) == 0) {  // hm_ntt + v.neg_ntt + sig.neg_ntt * pk_ntt
        p = s_y[N + i] * s_pkn[i];
        sum = s_hmn[i] + s_y[3 * N + i] + p;
      } else {  // v.pos_ntt + sig.pos_ntt * pk_ntt
        p = s_y[i] * s_pkn[i];
        sum = s_y[2 * N + i] + p;
      }
      t = sum / Q;
      b = sum - t * Q;
    };
#pragma unroll 1
    for (uint32_t c0 = 0; c0 < 2u * N; c0 += CH) {
      uint64_t* zc = z + 4 * (uint64_t)(L.n_inst + L.w_pw + c0 * 30);
      if ((uint32_t)tid < CH * 3) {
        const uint32_t rr = (uint32_t)tid / 3, j = (uint32_t)tid - rr * 3;
        uint32_t p, t, b;
        pw(c0 + rr, p, t, b);
        Fr v = mont_small(g_mont, j == 0 ? p : j == 1 ? t : b);
#pragma unroll
        for (int k = 0; k < 8; k++) s_stage[tid * 8 + k] = v.v[k];
      }
      __syncthreads();
      sweep_chunk<30, 3, 27, 0, 3>(zc, c0, CH, tid, s_stage, [&](uint32_t r) {
        uint32_t p, t, b;
        pw(r, p, t, b);
        return ltq_mask(b);
      });
      __syncthreads();
    }
    // (3) the 4N squares, in the order v.pos, v.neg, sig.pos, sig.neg (falcon_dual_ntt.rs:121-129)
    for (int k = tid; k < 4 * N; k += WT) {
      const uint32_t e = s_x[k < 2 * N ? 2 * N + k : k - 2 * N];
      store_fr(z + 4 * (uint64_t)(L.n_inst + L.w_l2 + k), mont_small(g_mont, e * e));
    }
    for (uint32_t w = tid; w < L.norm_bits + L.norm_ops; w += WT)
      store_bit(z + 4 * (uint64_t)(L.n_inst + L.w_norm + w), s_norm[w]);
  }
}

// ---------------------------------------------------------------------------------------------
// FalconSchoolBookVerificationCircuit (circuits/falcon_schoolbook.rs:26-132).
// z = [1 | pk | hm | sig | (v_i, 27 range witnesses) x N | column_0 .. column_{N-1} | l2 | norm]; column i
// holds t, c, the N products sig_k * buf_k of inner_product_mod (arithmetics.rs:34-100), the range
// witnesses of c and the two is_eq gadgets.  36.9 MB per Falcon-1024 signature, 91 % of it the N^2
// products: HBM-write bound.  Values below 2^28 are brought to Montgomery form with two table
// look-ups, mont(x) = T0[x mod 2^14] + T1[x >> 14] (T1[j] = mont(2^14 j)), instead of a multiplication.
struct SbParams {
  circuit::Layout L;
  NormOpsDev ops;
  uint32_t inv_q[8], inv_neg_q[8];  // q^-1 and (-q)^-1 mod r, Montgomery form (the is_neq multipliers)
};

__global__ void mont_table_kernel(uint32_t* tab) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 32768u) return;
  Fr x = Fr::zero();
  if (i < 16384u) {
    x.v[0] = i;
  } else {
    uint64_t v = (uint64_t)(i - 16384u) << 14;
    x.v[0] = (uint32_t)v;
    x.v[1] = (uint32_t)(v >> 32);
  }
  x = x.to_mont();
#pragma unroll
  for (int k = 0; k < 8; k++) tab[8 * i + k] = x.v[k];
}

template <int LOGN>
__global__ void __launch_bounds__(WT, 2)
    witness_sb_kernel(SbParams P, uint64_t n_sig, const uint16_t* __restrict__ g_sig, const uint16_t* __restrict__ g_pk,
                      const uint16_t* __restrict__ g_hm, const uint32_t* __restrict__ g_mont, uint64_t* __restrict__ g_z,
                      int32_t* __restrict__ g_status) {
  constexpr int N = 1 << LOGN;
  __shared__ uint32_t s_sig[N], s_pk[N], s_hm[N], s_v[N], s_c[N], s_t[N], s_rhs_ge[N / 32 + 1];
  __shared__ uint32_t s_norm[64];
  __shared__ unsigned long long s_acc;
  __shared__ int s_bad;
  const int tid = threadIdx.x;
  const circuit::Layout& L = P.L;
  for (uint64_t sid = blockIdx.x; sid < n_sig; sid += gridDim.x) {
    __syncthreads();
    if (tid == 0) {
      s_acc = 0;
      s_bad = 0;
    }
    if (tid < N / 32 + 1) s_rhs_ge[tid] = 0;
    uint64_t* z = g_z + sid * (uint64_t)L.n_z * 4;
    int bad = 0;
    for (int i = tid; i < N; i += WT) {
      uint32_t a = g_sig[sid * N + i], b = g_pk[sid * N + i], c = g_hm[sid * N + i];
      bad |= (a >= Q) | (b >= Q) | (c >= Q);
      s_sig[i] = modq(a);
      s_pk[i] = modq(b);
      s_hm[i] = modq(c);
    }
    __syncthreads();
    if (bad) s_bad = 1;
    // ---- column sums ab_i = sum_k sig_k * buf_k (integers), t = ab / q, c = ab % q; v = hm - sig*pk lifted
    for (int i = tid; i < N; i += WT) {
      unsigned long long ab = 0;
      for (int k = 0; k <= i; k++) ab += (unsigned long long)(s_sig[k] * s_pk[i - k]);
      for (int k = i + 1; k < N; k++) ab += (unsigned long long)(s_sig[k] * (Q - s_pk[N + i - k]));
      uint32_t t = (uint32_t)(ab / Q), c = (uint32_t)(ab - (unsigned long long)t * Q);
      s_t[i] = t;
      s_c[i] = c;
      uint32_t rhs = s_hm[i] + Q - c;  // in [1, 2q)
      bool ge = rhs >= Q;
      s_v[i] = ge ? rhs - Q : rhs;
      if (ge) atomicOr(&s_rhs_ge[i >> 5], 1u << (i & 31));
    }
    __syncthreads();
    // ---- l2 norm over v ++ sig (misc.rs:30-51) and the norm-bound witnesses
    unsigned long long local_norm = 0;
    for (int k = tid; k < 2 * N; k += WT) {
      uint32_t e = k < N ? s_v[k] : s_sig[k - N];
      uint32_t sft = e < 6144 ? e : Q - e;
      local_norm += (unsigned long long)sft * sft;
    }
    for (int o = 16; o > 0; o >>= 1) local_norm += __shfl_xor_sync(0xffffffffu, local_norm, o);
    if ((tid & 31) == 0) atomicAdd(&s_acc, local_norm);
    __syncthreads();
    if (tid == 0) {
      unsigned long long norm = s_acc;
      int st = 0;
      if (s_bad) st = FRCS_E_COEFF_RANGE;
      if (st == 0 && norm >= L.l2_bound) st = FRCS_E_NORM_BOUND;
      if (blockIdx.y == 0) g_status[sid] = st;
      for (uint32_t i = 0; i < L.norm_bits; i++) s_norm[i] = (uint32_t)((norm >> i) & 1);
      for (uint32_t k = 0; k < L.norm_ops; k++) {
        uint32_t a = s_norm[P.ops.a[k]], b = s_norm[P.ops.b[k]], r;
        switch (P.ops.kind[k]) {
          case circuit::OP_AND: r = a & b; break;
          case circuit::OP_OR: r = a | b; break;
          case circuit::OP_AND_NOT: r = a & (b ^ 1); break;
          default: r = (a ^ 1) & (b ^ 1); break;
        }
        s_norm[L.norm_bits + k] = r;
      }
    }
    __syncthreads();
    // ================= output =================
    // gridDim.y CTAs share a signature (each repeats the cheap prelude above): part p writes the columns
    // [p N / parts, (p + 1) N / parts); part 0 also writes everything that is not a column
    const int parts = gridDim.y, part = blockIdx.y;
    const int col_lo = (int)((long long)N * part / parts), col_hi = (int)((long long)N * (part + 1) / parts);
    if (part == 0) {
    // One, pk, hm (instance), sig
    for (int d = tid; d < 3 * N + 1; d += WT) {
      if (d == 3 * N) {
        store_bit(z, true);
        continue;
      }
      int g = d >> LOGN, i = d & (N - 1);
      uint32_t x = g == 0 ? s_pk[i] : g == 1 ? s_hm[i] : s_sig[i];
      uint32_t pos = g == 0 ? 1 + i : g == 1 ? 1 + N + i : L.n_inst + L.w_sig + i;
      store_fr(z + 4 * (uint64_t)pos, mont_small(g_mont, x));
    }
    // v[i] and its 27 range witnesses, interleaved
    for (uint32_t r = tid; r < 28u * N; r += WT) {
      uint32_t i = r / 28, j = r - 28 * i;
      uint64_t* dst = z + 4 * (uint64_t)(L.n_inst + L.w_v + r);
      if (j == 0)
        store_fr(dst, mont_small(g_mont, s_v[i]));
      else
        store_bit(dst, ltq_bit(s_v[i], j - 1));
    }
    }
    // the N columns
#pragma unroll 1
    for (int i = col_lo; i < col_hi; i++) {
      uint64_t* col = z + 4 * (uint64_t)(L.n_inst + L.w_cols + (uint64_t)L.sb_col * i);
      const uint32_t c = s_c[i];
      const bool ge = (s_rhs_ge[i >> 5] >> (i & 31)) & 1;  // rhs = v + q: ne1 = 1, ne2 = 0; else ne1 = 0, ne2 = 1
      for (uint32_t e = tid; e < L.sb_col; e += WT) {
        uint64_t* dst = col + 4 * (uint64_t)e;
        if (e >= 2 && e < 2u + N) {
          int k = e - 2;
          uint32_t m = k <= i ? s_sig[k] * s_pk[i - k] : s_sig[k] * (Q - s_pk[N + i - k]);
          store_fr(dst, mont_small(g_mont, m));
        } else if (e < 2) {
          store_fr(dst, mont_small(g_mont, e == 0 ? s_t[i] : c));
        } else {
          uint32_t j = e - 2 - N;
          if (j < 27) {
            store_bit(dst, ltq_bit(c, j));
          } else if (j == 27) {
            store_bit(dst, ge);  // ne1
          } else if (j == 29) {
            store_bit(dst, !ge);  // ne2
          } else if (j == 31) {
            store_bit(dst, false);  // ne2 & ne1
          } else {
            // multiplier of AllocatedFp::is_neq: (x - y)^-1 if x != y, else 1;  x - y is q (j = 28) or -q (j = 30)
            const bool ne = j == 28 ? ge : !ge;
            Fr m = Fr::one();
            if (ne) {
#pragma unroll
              for (int q = 0; q < 8; q++) m.v[q] = j == 28 ? P.inv_q[q] : P.inv_neg_q[q];
            }
            store_fr(dst, m);
          }
        }
      }
    }
    if (part == parts - 1) {
    // l2 elements (18 witnesses each) over v ++ sig
    for (uint32_t r = tid; r < 36u * N; r += WT) {
      uint32_t k = r / 18, j = r - 18 * k;
      uint32_t e = k < (uint32_t)N ? s_v[k] : s_sig[k - N];
      uint64_t* dst = z + 4 * (uint64_t)(L.n_inst + L.w_l2 + r);
      bool y1 = ((e >> 11) & 1) && ((e >> 12) & 1);
      if (j < 16) {
        store_bit(dst, j < 14 ? ((e >> j) & 1) : (j == 14 ? y1 : (!((e >> 13) & 1) && !y1)));
      } else {
        uint32_t sft = e < 6144 ? e : Q - e;
        store_fr(dst, mont_small(g_mont, j == 16 ? sft : sft * sft));
      }
    }
    for (uint32_t w = tid; w < L.norm_bits + L.norm_ops; w += WT)
      store_bit(z + 4 * (uint64_t)(L.n_inst + L.w_norm + w), s_norm[w]);
    }
  }
}

}  // namespace

// the table of Montgomery images (made once per context, at creation; complete when this returns)
int32_t ensure_mont_table(frcs_ctx* ctx, cudaStream_t st) {
  if (ctx->mont_tab) return FRCS_OK;
  FRCS_CUDA_CHECK(cudaMalloc(&ctx->mont_tab, 32768 * 32));
  mont_table_kernel<<<128, 256, 0, st>>>(ctx->mont_tab);
  ctx->launches++;
  FRCS_CUDA_CHECK(cudaStreamSynchronize(st));
  return FRCS_OK;
}

int32_t launch_witness(frcs_ctx* ctx, uint64_t n, const uint16_t* d_sig, const uint16_t* d_pk, const uint16_t* d_hm,
                       uint64_t* d_z, int32_t* d_status, cudaStream_t st) {
  if (n == 0) return FRCS_OK;
  NvtxRange nvtx("frcs:witness_generation");
  int32_t rc_tab = ensure_mont_table(ctx, st);
  if (rc_tab) return rc_tab;
  if (ctx->L.kind == FRCS_KIND_SCHOOLBOOK) {
    SbParams P;
    P.L = ctx->L;
    P.ops = ctx->norm_ops;
    // q^-1 and (-q)^-1 mod r in Montgomery form, by exponentiation on the host build of the field code
    {
      static const Fr iq = Fr::from_u32(Q).inverse(), inq = Fr::from_u32(Q).neg().inverse();
      for (int k = 0; k < 8; k++) {
        P.inv_q[k] = iq.v[k];
        P.inv_neg_q[k] = inq.v[k];
      }
    }
    int sms = 0;
    FRCS_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device));
    unsigned grid = (unsigned)(n < (uint64_t)sms * 2 ? n : (uint64_t)sms * 2);
    // few signatures (the single split proof of BASELINE configs[3]): several CTAs per signature, by column range
    unsigned parts = (unsigned)std::min<uint64_t>(64, std::max<uint64_t>(1, (uint64_t)sms * 2 / grid));
    int ph = prof_begin(ctx, PROF_WITNESS, st);
    if (ctx->L.logn == 10)
      witness_sb_kernel<10><<<dim3(grid, parts), WT, 0, st>>>(P, n, d_sig, d_pk, d_hm, ctx->mont_tab, d_z, d_status);
    else
      witness_sb_kernel<9><<<dim3(grid, parts), WT, 0, st>>>(P, n, d_sig, d_pk, d_hm, ctx->mont_tab, d_z, d_status);
    prof_end(ctx, ph, st);
    ctx->launches++;
    FRCS_CUDA_CHECK(cudaGetLastError());
    return FRCS_OK;
  }
  const bool dual = ctx->L.kind == FRCS_KIND_DUAL_NTT;
  WitnessParams P;
  P.L = ctx->L;
  P.ops = ctx->norm_ops;
  const uint32_t logn = ctx->L.logn, N = ctx->L.n;
  for (uint32_t l = 0; l < logn; l++) {
    circuit::U256 c = circuit::u256_pow2(l + 1);
    for (uint32_t e = 0; e < l + 2; e++) c = circuit::u256_mul_small(c, Q);
    for (int k = 0; k < 5; k++) P.cst[l][k] = c.v[k];
  }
  P.n_inv = circuit::powmod_q(N, Q - 2);
  // 92 KB for N = 1024 (dual circuit: 80 KB): two CTAs per SM
  size_t smem = (size_t)(dual ? 2 + 2 + 4 + 4 + 5 : 2 + 9 + 4 + 5) * N * 4 + 64 * 4 + 384 * 32;
  int sms = 0;
  FRCS_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device));
  unsigned grid = (unsigned)(n < (uint64_t)sms * 2 ? n : (uint64_t)sms * 2);  // persistent: 2 CTAs per SM, grid-stride
  const size_t tq_bytes = (size_t)grid * (dual ? 4 : 2) * N * 32;
  if (ctx->wit_scratch_bytes < tq_bytes) {
    FRCS_CUDA_CHECK(cudaDeviceSynchronize());
    cudaFree(ctx->wit_scratch);
    ctx->wit_scratch = nullptr;
    ctx->wit_scratch_bytes = 0;
    FRCS_CUDA_CHECK(cudaMalloc(&ctx->wit_scratch, tq_bytes));
    ctx->wit_scratch_bytes = tq_bytes;
  }
  int ph = prof_begin(ctx, PROF_WITNESS, st);
  if (dual) {
    if (logn == 10) {
      FRCS_CUDA_CHECK(cudaFuncSetAttribute(witness_dual_kernel<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      witness_dual_kernel<10><<<grid, WT, smem, st>>>(P, n, d_sig, d_pk, d_hm, ctx->ntt_tab, ctx->mont_tab, (uint32_t*)ctx->wit_scratch, d_z, d_status);
    } else {
      FRCS_CUDA_CHECK(cudaFuncSetAttribute(witness_dual_kernel<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      witness_dual_kernel<9><<<grid, WT, smem, st>>>(P, n, d_sig, d_pk, d_hm, ctx->ntt_tab, ctx->mont_tab, (uint32_t*)ctx->wit_scratch, d_z, d_status);
    }
  } else if (logn == 10) {
    FRCS_CUDA_CHECK(cudaFuncSetAttribute(witness_kernel<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    witness_kernel<10><<<grid, WT, smem, st>>>(P, n, d_sig, d_pk, d_hm, ctx->ntt_tab, ctx->mont_tab, (uint32_t*)ctx->wit_scratch, d_z, d_status);
  } else {
    FRCS_CUDA_CHECK(cudaFuncSetAttribute(witness_kernel<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    witness_kernel<9><<<grid, WT, smem, st>>>(P, n, d_sig, d_pk, d_hm, ctx->ntt_tab, ctx->mont_tab, (uint32_t*)ctx->wit_scratch, d_z, d_status);
  }
  prof_end(ctx, ph, st);
  ctx->launches++;
  FRCS_CUDA_CHECK(cudaGetLastError());
  return FRCS_OK;
}
