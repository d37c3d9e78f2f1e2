// probe_microbench4.cu -- asynchronous random 32-byte gathers into shared memory at LOW occupancy (the shape of a
// thread-per-read kernel whose probes are issued while the scan goes on): cp.async (LDGSTS) with commit groups and a
// double-buffered tile per thread vs TMA bulk copies with an mbarrier per (warp, tile).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/probe_microbench4.bin tools/probe_microbench4.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
__device__ __forceinline__ uint64_t mix(uint64_t x) {
  x += 0x9e3779b97f4a7c15ull; x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
  x = (x ^ (x >> 27)) * 0x94d049bb133111ebull; return x ^ (x >> 31);
}
// MODE 0: cp.async.ca 2 x 16 B; 1: cp.async.cg 2 x 16 B; 2: cp.async.ca with L2::128B prefetch hint
template <int CAP, int MODE, int THREADS>
__global__ void __launch_bounds__(THREADS) gather_ldgsts(const ulonglong2* table, uint64_t n_buckets, uint64_t per_thread, uint64_t* out, int spin) {
  extern __shared__ __align__(16) ulonglong2 buf[];   // [2][CAP][2][THREADS]
  uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, acc = 0;
  auto slot = [&](int tile, int e, int half) { return &buf[((tile * CAP + e) * 2 + half) * THREADS + threadIdx.x]; };
  int cur = 0;
  uint64_t i = 0;
  bool first = true;
  while (i < per_thread || !first) {
    if (i < per_thread) {
#pragma unroll
      for (int e = 0; e < CAP; e++) {
        uint64_t bkt = __umul64hi(mix(t * per_thread + i + e), n_buckets);
        uint32_t s0 = (uint32_t)__cvta_generic_to_shared(slot(cur, e, 0)), s1 = (uint32_t)__cvta_generic_to_shared(slot(cur, e, 1));
        if (MODE == 0) {
          asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(s0), "l"(table + bkt * 2));
          asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(s1), "l"(table + bkt * 2 + 1));
        } else if (MODE == 1) {
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s0), "l"(table + bkt * 2));
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s1), "l"(table + bkt * 2 + 1));
        } else {
          asm volatile("cp.async.ca.shared.global.L2::64B [%0], [%1], 16;" ::"r"(s0), "l"(table + bkt * 2));
          asm volatile("cp.async.ca.shared.global.L2::64B [%0], [%1], 16;" ::"r"(s1), "l"(table + bkt * 2 + 1));
        }
        for (int s = 0; s < spin; s++) acc = mix(acc);   // stand-in for the scan work between two super-mers
      }
      i += CAP;
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    if (!first || i >= per_thread) {
      if (i < per_thread || first) asm volatile("cp.async.wait_group 1;" ::: "memory");
      else asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    if (!first) {
#pragma unroll
      for (int e = 0; e < CAP; e++) { ulonglong2 a = *slot(cur ^ 1, e, 0), b = *slot(cur ^ 1, e, 1); acc += a.x ^ a.y ^ b.x ^ b.y; }
    }
    if (i >= per_thread && !first) {   // last tile
      asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
      for (int e = 0; e < CAP; e++) { ulonglong2 a = *slot(cur, e, 0), b = *slot(cur, e, 1); acc += a.x ^ a.y ^ b.x ^ b.y; }
      break;
    }
    first = false;
    cur ^= 1;
  }
  if (acc == 0x1234567) out[0] = acc;
}
// TMA bulk 32 B per gather, one mbarrier per (warp, tile), double-buffered
template <int CAP, int THREADS>
__global__ void __launch_bounds__(THREADS) gather_tma(const ulonglong2* table, uint64_t n_buckets, uint64_t per_thread, uint64_t* out, int spin) {
  extern __shared__ __align__(128) ulonglong2 buf[];   // [2][CAP][THREADS][2]
  __shared__ __align__(8) uint64_t bars[2][THREADS / 32];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&bars[0][warp])), "r"(32));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&bars[1][warp])), "r"(32));
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, acc = 0;
  auto slot = [&](int tile, int e) { return &buf[((tile * CAP + e) * THREADS + threadIdx.x) * 2]; };
  uint32_t phase[2] = {0, 0};
  auto wait = [&](int tile) {
    uint32_t b = (uint32_t)__cvta_generic_to_shared(&bars[tile][warp]), done = 0;
    while (!done)
      asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(b), "r"(phase[tile]) : "memory");
    phase[tile] ^= 1;
  };
  int cur = 0;
  bool first = true;
  for (uint64_t i = 0; i < per_thread; i += CAP) {
    uint32_t b = (uint32_t)__cvta_generic_to_shared(&bars[cur][warp]);
#pragma unroll
    for (int e = 0; e < CAP; e++) {
      uint64_t bkt = __umul64hi(mix(t * per_thread + i + e), n_buckets);
      uint32_t s0 = (uint32_t)__cvta_generic_to_shared(slot(cur, e));
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 32, [%2];" ::"r"(s0), "l"(table + bkt * 2), "r"(b) : "memory");
      for (int s = 0; s < spin; s++) acc = mix(acc);
    }
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(32 * CAP) : "memory");
    if (!first) {
      wait(cur ^ 1);
#pragma unroll
      for (int e = 0; e < CAP; e++) { ulonglong2 a = slot(cur ^ 1, e)[0], c = slot(cur ^ 1, e)[1]; acc += a.x ^ a.y ^ c.x ^ c.y; }
      __syncwarp();
    }
    first = false;
    cur ^= 1;
  }
  wait(cur ^ 1);
#pragma unroll
  for (int e = 0; e < CAP; e++) { ulonglong2 a = slot(cur ^ 1, e)[0], c = slot(cur ^ 1, e)[1]; acc += a.x ^ a.y ^ c.x ^ c.y; }
  if (acc == 0x1234567) out[0] = acc;
}
static int g_sms = 148;
template <class F> static void timeit(const char* name, int threads, int ctas_per_sm, uint64_t per, F launch) {
  int grid = g_sms * ctas_per_sm;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  launch(grid, (uint64_t)48);
  cudaEventRecord(e0); launch(grid, per); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  double n = (double)grid * threads * per;
  printf("%-44s %4d thr x %d CTA/SM : %7.2f G gathers/s (%.2f ms) %s\n", name, threads, ctas_per_sm, n / ms / 1e6, ms, cudaGetErrorString(cudaGetLastError()));
}
int main(int argc, char** argv) {
  double gb = argc > 1 ? atof(argv[1]) : 16.0;
  uint64_t per = argc > 2 ? strtoull(argv[2], 0, 10) : 1536;
  cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, 0);
  uint64_t nb = (uint64_t)(gb * 1e9 / 32) & ~3ull;
  ulonglong2* table; uint64_t* out;
  cudaMalloc(&table, nb * 32); cudaMalloc(&out, 8); cudaMemset(table, 1, nb * 32);
#define LDGSTS(CAP, MODE, THR, CTAS, SPIN, NAME) { auto k = gather_ldgsts<CAP, MODE, THR>; size_t sm = 2 * CAP * 2 * THR * 16; \
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm); \
    timeit(NAME, THR, CTAS, per, [&](int g, uint64_t p) { k<<<g, THR, sm>>>(table, nb, p, out, SPIN); }); }
#define TMA(CAP, THR, CTAS, SPIN, NAME) { auto k = gather_tma<CAP, THR>; size_t sm = 2 * CAP * 2 * THR * 16; \
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm); \
    timeit(NAME, THR, CTAS, per, [&](int g, uint64_t p) { k<<<g, THR, sm>>>(table, nb, p, out, SPIN); }); }
  for (int spin : {0, 20, 60}) {
    printf("-- spin %d (~%d dependent ALU instr between gathers)\n", spin, spin * 12);
    LDGSTS(8, 0, 128, 2, spin, "cp.async.ca cap 8");
    LDGSTS(8, 1, 128, 2, spin, "cp.async.cg cap 8");
    LDGSTS(8, 2, 128, 2, spin, "cp.async.ca.L2::64B cap 8");
    LDGSTS(6, 0, 128, 3, spin, "cp.async.ca cap 6");
    LDGSTS(4, 0, 128, 4, spin, "cp.async.ca cap 4");
    LDGSTS(8, 0, 128, 1, spin, "cp.async.ca cap 8");
    TMA(8, 128, 2, spin, "TMA bulk 32 B cap 8");
    TMA(6, 128, 3, spin, "TMA bulk 32 B cap 6");
    TMA(4, 128, 4, spin, "TMA bulk 32 B cap 4");
  }
  return cudaDeviceSynchronize() != cudaSuccess;
}
