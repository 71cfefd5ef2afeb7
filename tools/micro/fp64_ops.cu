// Throughput of the arithmetic the few-groups aggregate leans on (development aid): cycles per warp instruction per SM.
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void bench(double* out, int iters, float seed) {
  float f0 = seed + threadIdx.x, f1 = f0 * 0.5f, f2 = f0 * 0.25f, f3 = f0 * 0.125f;
  double a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0, a5 = 0, a6 = 0, a7 = 0;
  int g = threadIdx.x & 3;
  for (int i = 0; i < iters; ++i) {
    if (OP == 0) {  // 4 conversions f32 -> f64 (+ 4 f32 adds to keep inputs changing)
      a0 += 0; f0 += 1.0f; f1 += 1.0f; f2 += 1.0f; f3 += 1.0f;
      double d0 = f0, d1 = f1, d2 = f2, d3 = f3;
      asm volatile("" :: "d"(d0), "d"(d1), "d"(d2), "d"(d3));
    }
    if (OP == 1) {  // 8 independent DADD
      a0 += 1.5; a1 += 1.5; a2 += 1.5; a3 += 1.5; a4 += 1.5; a5 += 1.5; a6 += 1.5; a7 += 1.5;
      asm volatile("" : "+d"(a0), "+d"(a1), "+d"(a2), "+d"(a3), "+d"(a4), "+d"(a5), "+d"(a6), "+d"(a7));
    }
    if (OP == 2) {  // 8 predicated DADD, one in four lanes active
      double v = 1.5;
      asm volatile("{.reg .pred p; setp.eq.s32 p, %8, 0; @p add.rn.f64 %0, %0, %9; @p add.rn.f64 %1, %1, %9; @p add.rn.f64 %2, %2, %9; @p add.rn.f64 %3, %3, %9;"
                   "setp.eq.s32 p, %8, 1; @p add.rn.f64 %4, %4, %9; @p add.rn.f64 %5, %5, %9; @p add.rn.f64 %6, %6, %9; @p add.rn.f64 %7, %7, %9;}"
                   : "+d"(a0), "+d"(a1), "+d"(a2), "+d"(a3), "+d"(a4), "+d"(a5), "+d"(a6), "+d"(a7) : "r"(g), "d"(v));
    }
    if (OP == 3) {  // 8 FFMA
      f0 = fmaf(f0, 1.0001f, 0.5f); f1 = fmaf(f1, 1.0001f, 0.5f); f2 = fmaf(f2, 1.0001f, 0.5f); f3 = fmaf(f3, 1.0001f, 0.5f);
      f0 = fmaf(f0, 1.0001f, 0.5f); f1 = fmaf(f1, 1.0001f, 0.5f); f2 = fmaf(f2, 1.0001f, 0.5f); f3 = fmaf(f3, 1.0001f, 0.5f);
    }
    if (OP == 4) {  // f32 -> f64 by integer ops (normal numbers and zero only): 4 values
      f0 += 1.0f; f1 += 1.0f; f2 += 1.0f; f3 += 1.0f;
      float fs[4] = {f0, f1, f2, f3};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        unsigned b = __float_as_uint(fs[k]);
        unsigned hi = (b & 0x80000000u) | (((b >> 3) & 0x0FFFFFFFu) + ((b & 0x7FFFFFFFu) ? 0x38000000u : 0u));
        unsigned lo = b << 29;
        double d = __hiloint2double(hi, lo);
        asm volatile("" :: "d"(d));
      }
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + f0 + f1 + f2 + f3;
}
int main() {
  double* out; cudaMalloc(&out, 148 * 8 * 256 * 8);
  const char* names[] = {"4x F2F.F64.F32 (+4 FADD)", "8x DADD", "8x predicated DADD (1/4 lanes) + 2 SETP", "8x FFMA", "4x int-op f32->f64 (+4 FADD)"};
  const int per_iter[] = {4, 8, 8, 8, 4};
  const int iters = 4096;
  for (int op = 0; op < 5; ++op) {
    for (int blocks_per_sm : {1, 4}) {
      cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
      auto launch = [&]() {
        switch (op) {
          case 0: bench<0><<<148 * blocks_per_sm, 256>>>(out, iters, 1.f); break; case 1: bench<1><<<148 * blocks_per_sm, 256>>>(out, iters, 1.f); break;
          case 2: bench<2><<<148 * blocks_per_sm, 256>>>(out, iters, 1.f); break; case 3: bench<3><<<148 * blocks_per_sm, 256>>>(out, iters, 1.f); break;
          default: bench<4><<<148 * blocks_per_sm, 256>>>(out, iters, 1.f); break;
        }
      };
      launch(); cudaDeviceSynchronize();
      cudaEventRecord(a); launch(); cudaEventRecord(b); cudaEventSynchronize(b);
      float ms; cudaEventElapsedTime(&ms, a, b);
      double ops_per_sm = double(iters) * 8 * blocks_per_sm * per_iter[op];
      printf("%-44s blocks/SM %d: %.3f ms, %.2f cycles per warp-op per SM\n", names[op], blocks_per_sm, ms, ms * 1e-3 * 1.965e9 / ops_per_sm);
    }
  }
  return 0;
}
