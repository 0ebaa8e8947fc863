// Pipe-rate microbenchmarks for the field-arithmetic design (sm_100a).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench tools/ubench.cu && tools/ubench
// Each kernel keeps ILP independent dependent chains per thread whose operands depend on
// their own previous results, so ptxas cannot fold the products.
// Output: operations per clock per SM (at the SM clock measured with clock64) and ops/s.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>

#define ILP 8
#define ITERS 4096

template <int MODE>
__global__ void __launch_bounds__(256) k(uint64_t* out, uint32_t seed, unsigned long long* cyc) {
  uint32_t b = seed * 2654435761u + blockIdx.x * 977u + threadIdx.x;
  uint64_t acc[ILP];
  uint32_t lo[ILP], hi[ILP], x[ILP], y[ILP];
  double d[ILP], e = 1.0 + 1e-9 * threadIdx.x;
#pragma unroll
  for (int i = 0; i < ILP; i++) {
    acc[i] = (uint64_t)(b + i) * 0x9e3779b97f4a7c15ull;
    lo[i] = b + i * 7;
    hi[i] = b ^ (i * 13);
    x[i] = b + i;
    y[i] = b - i;
    d[i] = 1.0 + i + 1e-6 * threadIdx.x;
  }
  unsigned long long t0 = clock64();
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) {
      const int j = (i + 1) % ILP;
      if (MODE == 0) {  // IMAD.WIDE.U32, 64-bit accumulate, no carry in/out
        asm volatile("{.reg .u64 t; mov.b64 t, {%0,%1}; mad.wide.u32 t, %2, %3, t; mov.b64 {%0,%1}, t;}"
                     : "+r"(lo[i]), "+r"(hi[i]) : "r"(hi[j]), "r"(b));
      } else if (MODE == 1) {  // mad.lo.cc + madc.hi pair (carry inside the pair only)
        asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;" : "+r"(lo[i]), "+r"(hi[i]) : "r"(hi[j]), "r"(b));
      } else if (MODE == 8) {  // carry chains of 4 pairs, as in a field-multiplication row (IMAD.WIDE.U32.X)
        if ((i & 3) == 0)
          asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo[i]), "+r"(hi[i]) : "r"(hi[j]), "r"(b));
        else if ((i & 3) == 3)
          asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;" : "+r"(lo[i]), "+r"(hi[i]) : "r"(hi[j]), "r"(b));
        else
          asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo[i]), "+r"(hi[i]) : "r"(hi[j]), "r"(b));
      } else if (MODE == 2) {  // 32-bit IMAD (lo)
        asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(lo[i]) : "r"(lo[j]), "r"(b));
      } else if (MODE == 3) {  // IMAD.HI
        asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(lo[i]) : "r"(lo[j]), "r"(b));
      } else if (MODE == 4) {  // DFMA
        asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(d[i]) : "d"(d[j]), "d"(e));
      } else if (MODE == 5) {  // 64-bit add: IADD3 + IADD3.X
        asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(lo[i]), "+r"(hi[i]) : "r"(lo[j]), "r"(hi[j]));
      } else if (MODE == 6) {  // IMAD.WIDE + DFMA interleaved
        asm volatile("{.reg .u64 t; mov.b64 t, {%0,%1}; mad.wide.u32 t, %2, %3, t; mov.b64 {%0,%1}, t;}"
                     : "+r"(lo[i]), "+r"(hi[i]) : "r"(hi[j]), "r"(b));
        asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(d[i]) : "d"(d[j]), "d"(e));
      } else if (MODE == 7) {  // IMAD.WIDE + 64-bit add interleaved
        asm volatile("{.reg .u64 t; mov.b64 t, {%0,%1}; mad.wide.u32 t, %2, %3, t; mov.b64 {%0,%1}, t;}"
                     : "+r"(lo[i]), "+r"(hi[i]) : "r"(hi[j]), "r"(b));
        asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(x[i]), "+r"(y[i]) : "r"(x[j]), "r"(y[j]));
      } else if (MODE == 9) {  // DFMA + 64-bit add interleaved
        asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(d[i]) : "d"(d[j]), "d"(e));
        asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(x[i]), "+r"(y[i]) : "r"(x[j]), "r"(y[j]));
      } else if (MODE == 10) {  // 2 x 32-bit IMAD (lo + hi halves separately)
        asm volatile("mad.lo.u32 %0, %2, %3, %0; mad.hi.u32 %1, %2, %3, %1;" : "+r"(lo[i]), "+r"(hi[i]) : "r"(hi[j]), "r"(b));
      }
    }
  }
  unsigned long long t1 = clock64();
  uint64_t s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s ^= acc[i] ^ x[i] ^ y[i] ^ lo[i] ^ ((uint64_t)hi[i] << 32) ^ (uint64_t)__double_as_longlong(d[i]);
  if (s == 0x1234567) out[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int MODE>
void run(const char* name, double ops_per_inner, int sms) {
  uint64_t* out;
  unsigned long long* cyc;
  cudaMalloc(&out, 8);
  cudaMalloc(&cyc, 8);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  int blocks = sms * 8;
  double best = 1e30;
  unsigned long long c = 0;
  for (int rep = 0; rep < 4; rep++) {
    cudaEventRecord(e0);
    k<MODE><<<blocks, 256>>>(out, 12345 + rep, cyc);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep && ms < best) {
      best = ms;
      cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    }
  }
  double ops = (double)blocks * 256 * ITERS * ILP * ops_per_inner;
  // per SM per clock, using the in-kernel cycle count of one CTA (8 CTAs of 256 threads share an SM)
  double per_clk_sm = 8.0 * 256 * ITERS * ILP * ops_per_inner / (double)c;
  printf("%-44s %8.3f ms  %10.3e ops/s  %7.2f ops/clk/SM (clock64)  err=%s\n", name, best, ops / (best * 1e-3), per_clk_sm,
         cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
  cudaFree(cyc);
}

// write-only and copy bandwidth (32-byte stores per thread, like the witness kernel's z writes)
__global__ void __launch_bounds__(512) wr_kernel(uint64_t* dst, size_t n32) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n32; i += (size_t)gridDim.x * blockDim.x) {
    uint64_t v = i;
    asm volatile("st.global.v4.b64 [%0], {%1,%2,%3,%4};" ::"l"(dst + 4 * i), "l"(v), "l"(v), "l"(v), "l"(v) : "memory");
  }
}
__global__ void __launch_bounds__(512) cp_kernel(uint64_t* dst, const uint64_t* src, size_t n32) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n32; i += (size_t)gridDim.x * blockDim.x) {
    uint64_t a, b, c, d;
    asm volatile("ld.global.nc.v4.b64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(src + 4 * i));
    asm volatile("st.global.v4.b64 [%0], {%1,%2,%3,%4};" ::"l"(dst + 4 * i), "l"(a), "l"(b), "l"(c), "l"(d) : "memory");
  }
}
void run_bw(int sms) {
  size_t bytes = (size_t)3 << 30, n32 = bytes / 32;
  uint64_t *a, *b;
  cudaMalloc(&a, bytes);
  cudaMalloc(&b, bytes);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int mode = 0; mode < 3; mode++) {
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
      cudaEventRecord(e0);
      if (mode == 0) wr_kernel<<<sms * 4, 512>>>(a, n32);
      if (mode == 1) cp_kernel<<<sms * 4, 512>>>(b, a, n32);
      if (mode == 2) cudaMemsetAsync(a, 1, bytes);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (rep && ms < best) best = ms;
    }
    double gb = (mode == 1 ? 2.0 : 1.0) * bytes / 1e9;
    printf("%-44s %8.3f ms  %8.1f GB/s\n", mode == 0 ? "write-only 32 B stores (3 GiB)" : mode == 1 ? "copy (read+write bytes)" : "cudaMemset", best, gb / (best * 1e-3));
  }
  cudaFree(a);
  cudaFree(b);
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  printf("SMs %d\n", sms);
  run_bw(sms);
  run<0>("IMAD.WIDE.U32 (64-bit acc, no carry)", 1, sms);
  run<1>("mad.lo.cc+madc.hi.cc pair (1 LP)", 1, sms);
  run<8>("carry chains of 4 pairs (IMAD.WIDE.U32.X)", 1, sms);
  run<2>("IMAD lo 32", 1, sms);
  run<3>("IMAD.HI.U32", 1, sms);
  run<4>("DFMA", 1, sms);
  run<5>("IADD3 + IADD3.X pair", 1, sms);
  run<6>("IMAD.WIDE + DFMA (pairs)", 1, sms);
  run<7>("IMAD.WIDE + 64-bit add (pairs)", 1, sms);
  run<9>("DFMA + 64-bit add (pairs)", 1, sms);
  run<10>("IMAD lo + IMAD.HI (1 LP as two halves)", 1, sms);
  return 0;
}
