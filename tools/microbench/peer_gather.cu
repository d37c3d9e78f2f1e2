// Random 32-byte gathers from a PEER GPU's memory over NVLink (the lookup pattern of a table that is sharded over GPUs),
// against the same gathers from local HBM.  One process, two devices.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o peer_gather peer_gather.cu && ./peer_gather
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint64_t mix(uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33; return x; }

template <int ILP, int BYTES>
__global__ void __launch_bounds__(128, 4) gather(const uint64_t* __restrict__ table, uint64_t n_sectors, uint64_t per_thread, uint64_t seed,
                                                 unsigned long long* sink) {
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t acc = 0, s = mix(seed + tid);
  for (uint64_t i = 0; i < per_thread; i += ILP) {
    uint64_t v[ILP][4];
#pragma unroll
    for (int j = 0; j < ILP; j++) {
      s = mix(s + j);
      const uint64_t* p = table + (s % n_sectors) * 4;
      if (BYTES == 32)
        asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(v[j][0]), "=l"(v[j][1]), "=l"(v[j][2]), "=l"(v[j][3]) : "l"(p));
      else {
        asm volatile("ld.global.nc.L1::no_allocate.v2.u64 {%0,%1}, [%2];" : "=l"(v[j][0]), "=l"(v[j][1]) : "l"(p));
        v[j][2] = v[j][3] = 0;
      }
    }
#pragma unroll
    for (int j = 0; j < ILP; j++) acc += v[j][0] ^ v[j][1] ^ v[j][2] ^ v[j][3];
  }
  if (acc == 0x1234567) atomicAdd(sink, 1ull);
}

template <int ILP, int BYTES>
static double run(int dev, const uint64_t* table, uint64_t n_sectors, unsigned long long* sink, cudaStream_t st, int other_dev = -1,
                  const uint64_t* other_table = nullptr, unsigned long long* other_sink = nullptr, cudaStream_t other_st = nullptr) {
  const uint64_t threads = 148ull * 4 * 128 * 8, per = 256;
  cudaEvent_t a, b;
  CK(cudaSetDevice(dev)); CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  for (int rep = 0; rep < 2; rep++) {   // the first is the warm-up
    CK(cudaSetDevice(dev));
    CK(cudaEventRecord(a, st));
    gather<ILP, BYTES><<<(unsigned)(threads / 128), 128, 0, st>>>(table, n_sectors, per, 17 + rep, sink);
    CK(cudaEventRecord(b, st));
    if (other_dev >= 0) {
      CK(cudaSetDevice(other_dev));
      gather<ILP, BYTES><<<(unsigned)(threads / 128), 128, 0, other_st>>>(other_table, n_sectors, per, 99 + rep, other_sink);
      CK(cudaStreamSynchronize(other_st));
      CK(cudaSetDevice(dev));
    }
    CK(cudaStreamSynchronize(st));
  }
  float ms = 0; CK(cudaEventElapsedTime(&ms, a, b));
  return (double)threads * per / (ms * 1e-3) / 1e9;
}

int main() {
  int n = 0; CK(cudaGetDeviceCount(&n));
  if (n < 2) { printf("needs 2 GPUs\n"); return 0; }
  const uint64_t bytes = 16ull << 30, n_sectors = bytes / 32;
  uint64_t* t[2]; unsigned long long* sink[2]; cudaStream_t st[2];
  for (int d = 0; d < 2; d++) {
    CK(cudaSetDevice(d)); CK(cudaDeviceEnablePeerAccess(1 - d, 0));
    CK(cudaMalloc(&t[d], bytes)); CK(cudaMemset(t[d], 1, bytes)); CK(cudaMalloc(&sink[d], 8)); CK(cudaStreamCreate(&st[d]));
  }
  for (int d = 0; d < 2; d++) { CK(cudaSetDevice(d)); CK(cudaDeviceSynchronize()); }
  printf("random gathers out of a 16 GB table, 16 warps/SM, G gathers/s on GPU 0\n");
  printf("local  32 B, 4 in flight per thread: %.2f\n", run<4, 32>(0, t[0], n_sectors, sink[0], st[0]));
  printf("peer   32 B, 4 in flight per thread: %.2f\n", run<4, 32>(0, t[1], n_sectors, sink[0], st[0]));
  printf("peer   32 B, 8 in flight per thread: %.2f\n", run<8, 32>(0, t[1], n_sectors, sink[0], st[0]));
  printf("peer   16 B, 8 in flight per thread: %.2f\n", run<8, 16>(0, t[1], n_sectors, sink[0], st[0]));
  printf("peer   32 B, 8 in flight, both GPUs gathering from each other: %.2f\n",
         run<8, 32>(0, t[1], n_sectors, sink[0], st[0], 1, t[0], sink[1], st[1]));
  printf("peer   32 B, 4 in flight, both GPUs gathering from each other: %.2f\n",
         run<4, 32>(0, t[1], n_sectors, sink[0], st[0], 1, t[0], sink[1], st[1]));
  return 0;
}
