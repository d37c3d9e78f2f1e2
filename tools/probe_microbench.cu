// probe_microbench.cu -- what can this GPU do for the access pattern of the probe stage?
// Random 32-byte sector gathers from a table far larger than L2, one or several independent gathers in flight per
// thread. Prints G gathers/s and the equivalent GB/s of 32-byte sectors; run under ncu for dram bytes per gather.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/probe_microbench tools/probe_microbench.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t mix(uint64_t x) {
  x += 0x9e3779b97f4a7c15ull; x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
  x = (x ^ (x >> 27)) * 0x94d049bb133111ebull; return x ^ (x >> 31);
}

// 64-byte gathers: is a second sector of the same 64-byte half line free?
__global__ void gather64(const ulonglong2* __restrict__ table, uint64_t n_buckets, uint64_t per_thread, uint64_t* out) {
  uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, acc = 0;
  for (uint64_t i = 0; i < per_thread; i++) {
    uint64_t bkt = __umul64hi(mix(t * per_thread + i), n_buckets / 2);
    ulonglong2 a = __ldg(table + bkt * 4), b = __ldg(table + bkt * 4 + 1), c = __ldg(table + bkt * 4 + 2), d = __ldg(table + bkt * 4 + 3);
    acc += a.x ^ a.y ^ b.x ^ b.y ^ c.x ^ c.y ^ d.x ^ d.y;
  }
  if (acc == 0x1234567) out[0] = acc;
}

template <int ILP>
__global__ void gather(const ulonglong2* __restrict__ table, uint64_t n_buckets, uint64_t per_thread, uint64_t* out) {
  uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, acc = 0;
  for (uint64_t i = 0; i < per_thread; i += ILP) {
    ulonglong2 a[ILP], b[ILP];
#pragma unroll
    for (int u = 0; u < ILP; u++) {
      uint64_t bkt = __umul64hi(mix(t * per_thread + i + u), n_buckets);
      a[u] = __ldg(table + bkt * 2);
      b[u] = __ldg(table + bkt * 2 + 1);
    }
#pragma unroll
    for (int u = 0; u < ILP; u++) acc += a[u].x ^ a[u].y ^ b[u].x ^ b[u].y;
  }
  if (acc == 0x1234567) out[0] = acc;
}

template <int ILP>
static void run(const ulonglong2* table, uint64_t n_buckets, int threads_per_sm, uint64_t* out) {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int block = 128, grid = sms * threads_per_sm / block;
  uint64_t per_thread = 2048;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  gather<ILP><<<grid, block>>>(table, n_buckets, 64, out);
  cudaEventRecord(e0);
  gather<ILP><<<grid, block>>>(table, n_buckets, per_thread, out);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  double n = (double)grid * block * per_thread;
  printf("ILP %d threads/SM %4d : %7.2f G gathers/s  = %7.1f GB/s of 32-byte sectors  (%.2f ms)\n", ILP, threads_per_sm,
         n / ms / 1e6, n * 32 / ms / 1e6, ms);
}

int main(int argc, char** argv) {
  double gb = argc > 1 ? atof(argv[1]) : 20.0;
  uint64_t n_buckets = (uint64_t)(gb * 1e9 / 32);
  ulonglong2* table; uint64_t* out;
  cudaMalloc(&table, n_buckets * 32);
  cudaMalloc(&out, 8);
  cudaMemset(table, 1, n_buckets * 32);
  if (argc > 2) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, atoi(argv[2]));
  size_t g = 0;
  cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity);
  printf("table %.1f GB, L2 fetch granularity limit %zu\n", gb, g);
  for (int tps : {256, 512, 1024}) run<1>(table, n_buckets, tps, out);
  {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int grid = sms * 512 / 128;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    gather64<<<grid, 128>>>(table, n_buckets, 64, out);
    cudaEventRecord(e0);
    gather64<<<grid, 128>>>(table, n_buckets, 2048, out);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    double n = (double)grid * 128 * 2048;
    printf("64-byte gathers, 512 threads/SM: %7.2f G gathers/s = %7.1f GB/s (%.2f ms)\n", n / ms / 1e6, n * 64 / ms / 1e6, ms);
  }
  return cudaDeviceSynchronize() != cudaSuccess;
}
