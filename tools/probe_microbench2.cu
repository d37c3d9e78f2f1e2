// probe_microbench2.cu -- which load flavour fetches single 32-byte sectors for random probes?
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
__device__ __forceinline__ uint64_t mix(uint64_t x) {
  x += 0x9e3779b97f4a7c15ull; x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
  x = (x ^ (x >> 27)) * 0x94d049bb133111ebull; return x ^ (x >> 31);
}
template <int MODE> __device__ __forceinline__ ulonglong2 ld16(const ulonglong2* p) {
  ulonglong2 v;
  if (MODE == 0) v = __ldg(p);
  else if (MODE == 1) v = *p;
  else if (MODE == 2) asm volatile("ld.global.cg.v2.u64 {%0,%1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p));
  else if (MODE == 3) asm volatile("ld.global.cv.v2.u64 {%0,%1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p));
  else if (MODE == 4) asm volatile("ld.global.cs.v2.u64 {%0,%1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p));
  else if (MODE == 5) asm volatile("ld.global.L1::no_allocate.v2.u64 {%0,%1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p));
  else asm volatile("ld.global.nc.L1::no_allocate.v2.u64 {%0,%1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p));
  return v;
}
template <int MODE>
__global__ void gather(const ulonglong2* table, uint64_t n_buckets, uint64_t per_thread, uint64_t* out) {
  uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, acc = 0;
  for (uint64_t i = 0; i < per_thread; i++) {
    uint64_t bkt = __umul64hi(mix(t * per_thread + i), n_buckets);
    ulonglong2 a = ld16<MODE>(table + bkt * 2), b = ld16<MODE>(table + bkt * 2 + 1);
    acc += a.x ^ a.y ^ b.x ^ b.y;
  }
  if (acc == 0x1234567) out[0] = acc;
}
// one 32-byte load instruction per gather (sm_100: ld.global.v4.u64)
__global__ void gather256(const ulonglong2* table, uint64_t n_buckets, uint64_t per_thread, uint64_t* out) {
  uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, acc = 0;
  for (uint64_t i = 0; i < per_thread; i++) {
    uint64_t bkt = __umul64hi(mix(t * per_thread + i), n_buckets);
    uint64_t a, b, c, d;
    asm volatile("ld.global.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(table + bkt * 2));
    acc += a ^ b ^ c ^ d;
  }
  if (acc == 0x1234567) out[0] = acc;
}
template <class K> static void timeit(const char* name, K k, const ulonglong2* table, uint64_t nb, uint64_t* out) {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int grid = sms * 512 / 128; uint64_t per = 2048;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<<<grid, 128>>>(table, nb, 64, out);
  cudaEventRecord(e0); k<<<grid, 128>>>(table, nb, per, out); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  double n = (double)grid * 128 * per;
  printf("%-40s %7.2f G gathers/s (%.2f ms) %s\n", name, n / ms / 1e6, ms, cudaGetErrorString(cudaGetLastError()));
}
int main(int argc, char** argv) {
  double gb = argc > 1 ? atof(argv[1]) : 20.0;
  uint64_t nb = (uint64_t)(gb * 1e9 / 32);
  ulonglong2* table; uint64_t* out;
  cudaMalloc(&table, nb * 32); cudaMalloc(&out, 8); cudaMemset(table, 1, nb * 32);
  if (argc > 2) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, atoi(argv[2]));
  size_t g = 0; cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity);
  printf("table %.1f GB, L2 fetch granularity limit %zu\n", gb, g);
  timeit("__ldg (ld.global.nc)", gather<0>, table, nb, out);
  timeit("plain ld.global", gather<1>, table, nb, out);
  timeit("ld.global.cg", gather<2>, table, nb, out);
  timeit("ld.global.cv", gather<3>, table, nb, out);
  timeit("ld.global.cs", gather<4>, table, nb, out);
  timeit("ld.global.L1::no_allocate", gather<5>, table, nb, out);
  timeit("ld.global.nc.L1::no_allocate", gather<6>, table, nb, out);
  timeit("ld.global.v4.u64 (one 32-byte load)", gather256, table, nb, out);
  return cudaDeviceSynchronize() != cudaSuccess;
}
