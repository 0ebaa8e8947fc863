// frcs_selftest: element-wise field / curve operations through the same source on the
// host (on_device = 0) or as one thread per item on the GPU (on_device = 1).  Used by
// tests/ to pin ff32.cuh and ec.cuh against Python big-integer arithmetic and the oracle.
//
// op: 0 Fr mul, 1 Fr add, 2 Fr sub, 3 Fr inverse, 4 Fr to_mont, 5 Fr from_mont
//     10 Fq mul, 11 Fq add, 12 Fq sub, 13 Fq inverse;  6 / 16: Fr / Fq inverse by the windowed power (inverse_w4);
//     7 / 17: Fr / Fq square (on the device: the dedicated squaring of ff32.cuh)
//     18: Fq a*b - c*d through the fused single-reduction form (mul_sub2_inline; in: a, b, c, d)
//     24 G1 affine + affine through pair_classify / pair_finish with its own inversion (in: two points, out: sum)
//     34 the same for G2
//     20 G1 add (affine+affine), 21 G1 double, 22 G1 scalar mul (point | 4-limb canonical scalar)
//     30 G2 add, 31 G2 double, 32 G2 scalar mul
//     23 / 33: G1 / G2 window multiples 2^(16k) P, k = 1..15 (MSM base pre-processing)
// in/out are arrays of uint64 (Montgomery field elements, little-endian limbs).
#include "ctx.hpp"
#include "ec.cuh"

using namespace ff;

namespace {

template <class F>
FF_HD F load_f(const uint64_t* p) {
  F r;
#pragma unroll
  for (int i = 0; i < F::N / 2; i++) {
    r.v[2 * i] = (uint32_t)p[i];
    r.v[2 * i + 1] = (uint32_t)(p[i] >> 32);
  }
  return r;
}
template <class F>
FF_HD void store_f(uint64_t* p, const F& x) {
#pragma unroll
  for (int i = 0; i < F::N / 2; i++) p[i] = (uint64_t)x.v[2 * i] | ((uint64_t)x.v[2 * i + 1] << 32);
}
FF_HD Fq2 load_fq2(const uint64_t* p) { return {load_f<Fq>(p), load_f<Fq>(p + 6)}; }
FF_HD void store_fq2(uint64_t* p, const Fq2& x) {
  store_f<Fq>(p, x.c0);
  store_f<Fq>(p + 6, x.c1);
}
FF_HD ec::G1Affine load_g1(const uint64_t* p) { return {load_f<Fq>(p), load_f<Fq>(p + 6)}; }
FF_HD void store_g1(uint64_t* p, const ec::G1Affine& a) {
  store_f<Fq>(p, a.x);
  store_f<Fq>(p + 6, a.y);
}
FF_HD ec::G2Affine load_g2(const uint64_t* p) { return {load_fq2(p), load_fq2(p + 12)}; }
FF_HD void store_g2(uint64_t* p, const ec::G2Affine& a) {
  store_fq2(p, a.x);
  store_fq2(p + 12, a.y);
}

template <class F>
FF_HD void field_op(int op, const uint64_t* in, uint64_t* out) {
  constexpr int W = F::N / 2;
  F a = load_f<F>(in);
  switch (op) {
    case 0: store_f<F>(out, a * load_f<F>(in + W)); break;
    case 1: store_f<F>(out, a + load_f<F>(in + W)); break;
    case 2: store_f<F>(out, a - load_f<F>(in + W)); break;
    case 3: store_f<F>(out, a.inverse()); break;
    case 4: store_f<F>(out, a.to_mont()); break;
    case 5: store_f<F>(out, a.from_mont()); break;
    case 6: store_f<F>(out, F::inverse_w4(a)); break;
    case 7:
#ifdef __CUDA_ARCH__
      store_f<F>(out, a.sqr_dev());
#else
      store_f<F>(out, a.sqr());
#endif
      break;
  }
}

FF_HD void one_item(int op, const uint64_t* in, uint64_t* out, uint64_t i) {
  if (op < 10) {
    int w = (op <= 2) ? 8 : 4;  // two operands or one
    field_op<Fr>(op, in + i * w, out + i * 4);
  } else if (op == 18) {
    Fq r;
    Fq::mul_sub2_inline(r, load_f<Fq>(in + i * 24), load_f<Fq>(in + i * 24 + 6), load_f<Fq>(in + i * 24 + 12),
                        load_f<Fq>(in + i * 24 + 18));
    store_f<Fq>(out + i * 6, r);
  } else if (op < 20) {
    int w = (op - 10 <= 2) ? 12 : 6;
    field_op<Fq>(op - 10, in + i * w, out + i * 6);
  } else if (op == 20) {
    ec::G1 p = ec::G1::from_affine(load_g1(in + i * 24));
    p.add_mixed(load_g1(in + i * 24 + 12));
    store_g1(out + i * 12, p.to_affine());
  } else if (op == 24) {
    const ec::G1Affine p1 = load_g1(in + i * 24), p2 = load_g1(in + i * 24 + 12);
    Fq den = Fq::one();
    const int kind = ec::pair_classify<Fq>(p1, p2, true, den);
    store_g1(out + i * 12, ec::pair_finish<Fq>(kind, p1, p2, Fq::inverse_w4(den)));
  } else if (op == 34) {
    const ec::G2Affine p1 = load_g2(in + i * 48), p2 = load_g2(in + i * 48 + 24);
    Fq2 den = Fq2::one();
    const int kind = ec::pair_classify<Fq2>(p1, p2, true, den);
    store_g2(out + i * 24, ec::pair_finish<Fq2>(kind, p1, p2, Fq2::inverse_w4(den)));
  } else if (op == 21) {
    store_g1(out + i * 12, ec::G1::dbl_affine(load_g1(in + i * 12)).to_affine());
  } else if (op == 22) {
    uint32_t k[8];
    for (int j = 0; j < 4; j++) {
      k[2 * j] = (uint32_t)in[i * 16 + 12 + j];
      k[2 * j + 1] = (uint32_t)(in[i * 16 + 12 + j] >> 32);
    }
    store_g1(out + i * 12, ec::G1::from_affine(load_g1(in + i * 16)).mul(k, 255).to_affine());
  } else if (op == 23) {
    ec::G1Affine w[15];
    ec::window_multiples<Fq, 16>(load_g1(in + i * 12), w);
    for (int k = 0; k < 15; k++) store_g1(out + (i * 15 + k) * 12, w[k]);
  } else if (op == 33) {
    ec::G2Affine w[15];
    ec::window_multiples<Fq2, 16>(load_g2(in + i * 24), w);
    for (int k = 0; k < 15; k++) store_g2(out + (i * 15 + k) * 24, w[k]);
  } else if (op == 30) {
    ec::G2 p = ec::G2::from_affine(load_g2(in + i * 48));
    p.add_mixed(load_g2(in + i * 48 + 24));
    store_g2(out + i * 24, p.to_affine());
  } else if (op == 31) {
    store_g2(out + i * 24, ec::G2::dbl_affine(load_g2(in + i * 24)).to_affine());
  } else if (op == 32) {
    uint32_t k[8];
    for (int j = 0; j < 4; j++) {
      k[2 * j] = (uint32_t)in[i * 28 + 24 + j];
      k[2 * j + 1] = (uint32_t)(in[i * 28 + 24 + j] >> 32);
    }
    store_g2(out + i * 24, ec::G2::from_affine(load_g2(in + i * 28)).mul(k, 255).to_affine());
  }
}

__global__ void selftest_kernel(int op, const uint64_t* in, uint64_t* out, uint64_t n) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i < n) one_item(op, in, out, i);
}

int in_words(int op) {
  switch (op) {
    case 0: case 1: case 2: return 8;
    case 3: case 4: case 5: case 6: case 7: return 4;
    case 10: case 11: case 12: return 12;
    case 13: case 16: case 17: return 6;
    case 18: case 20: case 24: return 24;
    case 21: return 12;
    case 23: return 12;
    case 33: return 24;
    case 22: return 16;
    case 30: case 34: return 48;
    case 31: return 24;
    case 32: return 28;
  }
  return 0;
}
int out_words(int op) { return op == 23 ? 180 : op == 33 ? 360 : op < 10 ? 4 : op < 20 ? 6 : op < 30 ? 12 : 24; }

}  // namespace

extern "C" int32_t frcs_selftest(int32_t op, int32_t on_device, const uint64_t* in, uint64_t n, uint64_t* out) {
  int iw = in_words(op), ow = out_words(op);
  if (!iw || !in || !out) {
    frcs_set_error("frcs_selftest: bad op or null buffer");
    return FRCS_E_INVALID_ARG;
  }
  if (!on_device) {
    for (uint64_t i = 0; i < n; i++) one_item(op, in, out, i);
    return FRCS_OK;
  }
  uint64_t *d_in = nullptr, *d_out = nullptr;
  FRCS_CUDA_CHECK(cudaMalloc(&d_in, n * iw * 8));
  FRCS_CUDA_CHECK(cudaMalloc(&d_out, n * ow * 8));
  FRCS_CUDA_CHECK(cudaMemcpy(d_in, in, n * iw * 8, cudaMemcpyHostToDevice));
  selftest_kernel<<<(unsigned)((n + 63) / 64), 64>>>(op, d_in, d_out, n);
  FRCS_CUDA_CHECK(cudaGetLastError());
  FRCS_CUDA_CHECK(cudaMemcpy(out, d_out, n * ow * 8, cudaMemcpyDeviceToHost));
  cudaFree(d_in);
  cudaFree(d_out);
  return FRCS_OK;
}
