// Throughput of the warp primitives the join's multi-split leans on (development aid): cycles per warp-instruction per SM.
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void bench(unsigned* out, int iters, unsigned seed) {
  __shared__ unsigned hist[256];
  if (threadIdx.x < 256) hist[threadIdx.x] = 0;
  __syncthreads();
  unsigned lane = threadIdx.x & 31;
  unsigned x = seed + (threadIdx.x >> 2);  // runs of 4 equal values, ~8 distinct per warp
  unsigned acc = 0;
  for (int i = 0; i < iters; ++i) {
    unsigned p = (x + i * 7) & 255;
    if (OP == 0) acc += __shfl_xor_sync(0xffffffffu, p, 1);
    if (OP == 1) acc += __match_any_sync(0xffffffffu, p);
    if (OP == 2) acc += __ballot_sync(0xffffffffu, p & 1);
    if (OP == 3) atomicAdd(&hist[p], 1u);                        // all lanes, 4-way same address
    if (OP == 4) { if ((lane & 3) == 0) atomicAdd(&hist[p], 4u); }  // leaders only (8 lanes)
    if (OP == 5) acc += atomicAdd(&hist[p], 1u);                  // with return value
    if (OP == 6) { unsigned prev = __shfl_up_sync(0xffffffffu, p, 1); unsigned b = __ballot_sync(0xffffffffu, prev != p || lane == 0); acc += b; }
    if (OP == 7) acc += __popc(p) + __ffs(p);
    if (OP == 8) acc += __reduce_add_sync(0xffffffffu, p);
  }
  __syncthreads();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + hist[threadIdx.x & 255];
}
int main() {
  unsigned* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
  const char* names[] = {"shfl", "match_any", "ballot", "atoms all lanes 4-way", "atoms 8 leaders", "atoms with return", "shfl_up+ballot runs", "popc+ffs", "redux.add"};
  const int iters = 4096;
  for (int op = 0; op < 9; ++op) {
    for (int blocks_per_sm : {1, 4}) {
      cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
      auto launch = [&](int o) {
        switch (o) {
          case 0: bench<0><<<148 * blocks_per_sm, 256>>>(out, iters, 1); break; case 1: bench<1><<<148 * blocks_per_sm, 256>>>(out, iters, 1); break;
          case 2: bench<2><<<148 * blocks_per_sm, 256>>>(out, iters, 1); break; case 3: bench<3><<<148 * blocks_per_sm, 256>>>(out, iters, 1); break;
          case 4: bench<4><<<148 * blocks_per_sm, 256>>>(out, iters, 1); break; case 5: bench<5><<<148 * blocks_per_sm, 256>>>(out, iters, 1); break;
          case 6: bench<6><<<148 * blocks_per_sm, 256>>>(out, iters, 1); break; case 7: bench<7><<<148 * blocks_per_sm, 256>>>(out, iters, 1); break;
          default: bench<8><<<148 * blocks_per_sm, 256>>>(out, iters, 1); break;
        }
      };
      launch(op); cudaDeviceSynchronize();
      cudaEventRecord(a); launch(op); cudaEventRecord(b); cudaEventSynchronize(b);
      float ms; cudaEventElapsedTime(&ms, a, b);
      double warp_instr_per_sm = double(iters) * 8 * blocks_per_sm;
      printf("%-24s blocks/SM %d: %.3f ms, %.2f cycles per warp-op per SM (at 1.965 GHz)\n", names[op], blocks_per_sm, ms,
             ms * 1e-3 * 1.965e9 / warp_instr_per_sm);
    }
  }
  return 0;
}
