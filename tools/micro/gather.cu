// Cost of the per-row dictionary gather (l_extendedprice: ~60 K float entries per chunk, random value-IDs) through the
// three paths an SM has: L1/L2 (ld.global.nc), shared memory, and a cluster partner's shared memory (DSMEM).
// Prints SM cycles per warp-level gather instruction (32 random 4-byte reads). Development aid.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather gather.cu && ./gather
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>
namespace cg = cooperative_groups;

constexpr int kEntries = 61440;        // 240 KB
constexpr int kHalf = kEntries / 2;    // 120 KB
constexpr int kThreads = 512;

__device__ __forceinline__ unsigned next(unsigned& state) {
  state = state * 1664525u + 1013904223u;
  return (state >> 8) % kEntries;
}

// mode 0: all through L1/L2; 1: all from shared (index folded into the staged half); 2: lower half shared, upper half L1
__global__ void __launch_bounds__(kThreads) gather_kernel(const float* __restrict__ table, float* out, int iters, int mode) {
  extern __shared__ float s_table[];
  if (mode != 0) {
    for (int i = threadIdx.x; i < kHalf; i += kThreads) s_table[i] = table[i];
  }
  __syncthreads();
  unsigned state = blockIdx.x * 7919u + threadIdx.x * 104729u + 1u;
  float acc = 0.f;
  for (int i = 0; i < iters; ++i) {
    const unsigned index = next(state);
    if (mode == 0) {
      acc += __ldg(table + index);
    } else if (mode == 1) {
      acc += s_table[index >= kHalf ? index - kHalf : index];
    } else {
      acc += index < kHalf ? s_table[index] : __ldg(table + index);
    }
  }
  out[blockIdx.x * kThreads + threadIdx.x] = acc;
}

// cluster of 2: CTA r stages half r; a gather goes to the local or the partner's shared memory
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads) gather_dsmem_kernel(const float* __restrict__ table, float* out,
                                                                                         int iters) {
  extern __shared__ float s_table[];
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned rank = cluster.block_rank();
  for (int i = threadIdx.x; i < kHalf; i += kThreads) s_table[i] = table[rank * kHalf + i];
  cluster.sync();
  const float* halves[2] = {cluster.map_shared_rank(s_table, 0), cluster.map_shared_rank(s_table, 1)};
  unsigned state = blockIdx.x * 7919u + threadIdx.x * 104729u + 1u;
  float acc = 0.f;
  for (int i = 0; i < iters; ++i) {
    const unsigned index = next(state);
    const unsigned half = index >= kHalf;
    acc += halves[half][index - half * kHalf];
  }
  cluster.sync();
  out[blockIdx.x * kThreads + threadIdx.x] = acc;
}

int main() {
  std::vector<float> host(kEntries);
  for (int i = 0; i < kEntries; ++i) host[i] = 900.f + i * 0.5f;
  float *table, *out;
  cudaMalloc(&table, kEntries * sizeof(float));
  cudaMalloc(&out, 148 * 2 * kThreads * sizeof(float));
  cudaMemcpy(table, host.data(), kEntries * sizeof(float), cudaMemcpyHostToDevice);
  const int iters = 2048;
  const size_t shared_bytes = kHalf * sizeof(float);
  cudaFuncSetAttribute(gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shared_bytes);
  cudaFuncSetAttribute(gather_dsmem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shared_bytes);
  const char* names[] = {"L1/L2 (ld.global.nc)", "shared memory", "half shared / half L1", "cluster pair (DSMEM)"};
  for (int mode = 0; mode < 4; ++mode) {
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    auto launch = [&]() {
      if (mode < 3) {
        gather_kernel<<<148, kThreads, mode == 0 ? 0 : shared_bytes>>>(table, out, iters, mode);
      } else {
        gather_dsmem_kernel<<<148, kThreads, shared_bytes>>>(table, out, iters);
      }
    };
    launch();
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    launch();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    const double warp_gathers_per_sm = double(iters) * kThreads / 32;
    printf("%-24s %.3f ms  %.1f SM cycles per warp gather (at 1.965 GHz)  [%s]\n", names[mode], ms,
           ms * 1e-3 * 1.965e9 / warp_gathers_per_sm, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
