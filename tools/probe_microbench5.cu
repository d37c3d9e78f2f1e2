// probe_microbench5.cu -- does a stream of random table gathers slow down OTHER warps of the same SM that only use shared
// memory and the ALU? (The fused classify kernel loses 40 % when its table moves from L2 to DRAM although every bucket has
// landed by the time it is used; ncu shows the shared-memory waits of the scanning code going up)
// One block of 16 warps per SM: G gather warps issue random 32-byte cp.async (or plain ld.global.nc) gathers in tiles of
// 192 per warp, like the classify kernel; the other warps run a fixed number of rounds of a "scan step stand-in":
//   kind 0: dependent shared-memory loads (pointer chase in the warp's own region) + a little ALU
//   kind 1: shared-memory stores (the append) + ALU, no load
//   kind 2: ALU only
// Reported: time the workers need for their rounds, alone and next to the gathers, with the table in L2 and in DRAM,
// and the gather rate reached meanwhile.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/probe_microbench5.bin tools/probe_microbench5.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
__device__ __forceinline__ uint64_t mix(uint64_t x) {
  x += 0x9e3779b97f4a7c15ull; x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
  x = (x ^ (x >> 27)) * 0x94d049bb133111ebull; return x ^ (x >> 31);
}
#define WARPS 16
#define TILE 192
// flags[0]: workers still running (gather warps stop when it reaches 0)
template <int GATHER_MODE>   // 0: cp.async.ca 2 x 16 B into shared memory, 1: ld.global.nc 2 x 16 B into registers
__global__ void __launch_bounds__(WARPS * 32) interfere(const ulonglong2* table, uint64_t n_buckets, int n_gather, int kind, int rounds,
                                                        unsigned long long* worker_cycles, unsigned long long* gathers, uint64_t* sink) {
  extern __shared__ __align__(16) uint8_t smem[];
  __shared__ int workers_left;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) workers_left = WARPS - n_gather;
  // per warp: 192 x 32 B staging (gather warps) or a 6 KB private region (workers)
  uint8_t* mine = smem + (size_t)warp * TILE * 32;
  uint32_t* region = reinterpret_cast<uint32_t*>(mine);
  for (int i = lane; i < TILE * 8; i += 32) region[i] = (uint32_t)((i * 167 + 13) % (TILE * 8));   // a permutation to chase
  __syncthreads();
  uint64_t acc = 0;
  if (warp < n_gather) {
    uint64_t seed = (((uint64_t)blockIdx.x * WARPS + warp) << 40) + lane, done = 0;   // every gather its own random bucket
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(mine);
    while (*(volatile int*)&workers_left > 0 || n_gather == WARPS) {
      for (int e = lane; e < TILE; e += 32) {
        const uint64_t bkt = __umul64hi(mix(seed), n_buckets);
        seed += 32;
        if (GATHER_MODE == 0) {
          asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(sbase + e * 32), "l"(table + bkt * 2));
          asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(sbase + e * 32 + 16), "l"(table + bkt * 2 + 1));
        } else {
          ulonglong2 a, b;
          asm volatile("ld.global.nc.v2.u64 {%0,%1}, [%2];" : "=l"(a.x), "=l"(a.y) : "l"(table + bkt * 2));
          asm volatile("ld.global.nc.v2.u64 {%0,%1}, [%2];" : "=l"(b.x), "=l"(b.y) : "l"(table + bkt * 2 + 1));
          acc += a.x ^ a.y ^ b.x ^ b.y;
        }
      }
      if (GATHER_MODE == 0) {
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        acc += *reinterpret_cast<uint64_t*>(mine + lane * 32);
      }
      done += TILE;
      if (n_gather == WARPS && done >= (uint64_t)rounds * 8) break;   // gather-only run: fixed amount
    }
    if (lane == 0) atomicAdd(gathers, (unsigned long long)done);
  } else {
    const long long t0 = clock64();
    uint32_t idx = lane;
    for (int r = 0; r < rounds; r++) {
      if (kind == 0) {
#pragma unroll
        for (int j = 0; j < 4; j++) { idx = region[idx]; acc += idx; }
      } else if (kind == 1) {
#pragma unroll
        for (int j = 0; j < 4; j++) region[(idx + j * 32) % (TILE * 8)] = (uint32_t)acc;
        idx = (idx + 1) % (TILE * 8);
      }
#pragma unroll
      for (int j = 0; j < 8; j++) acc = acc * 0x9e3779b97f4a7c15ull + (acc >> 29);
      if (kind == 2 && acc == 0x7654321) region[lane] = (uint32_t)acc;   // keeps the ALU-only loop alive
    }
    const long long t1 = clock64();
    if (lane == 0) { atomicMax(worker_cycles, (unsigned long long)(t1 - t0)); atomicSub(&workers_left, 1); }
  }
  if (acc == 0x1234567) sink[0] = acc;
}

int main(int argc, char** argv) {
  const int rounds = argc > 1 ? atoi(argv[1]) : 200000;
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount;
  const size_t smem = (size_t)WARPS * TILE * 32;
  cudaFuncSetAttribute(interfere<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(interfere<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  unsigned long long *d_cyc, *d_g; uint64_t* d_sink;
  cudaMalloc(&d_cyc, 8); cudaMalloc(&d_g, 8); cudaMalloc(&d_sink, 8);
  const double table_gb[2] = {0.0625, 16.0};
  const char* table_name[2] = {"64 MB (L2)", "16 GB (DRAM)"};
  const char* kind_name[3] = {"smem pointer chase + ALU", "smem stores + ALU", "ALU only"};
  printf("%d SMs, %d worker rounds; one block of %d warps per SM, G gather warps, %d workers\n", sms, rounds, WARPS, WARPS);
  for (int ti = 0; ti < 2; ti++) {
    const uint64_t n_buckets = (uint64_t)(table_gb[ti] * (1ull << 30)) / 32;
    ulonglong2* table;
    if (cudaMalloc(&table, n_buckets * 32) != cudaSuccess) { printf("table allocation failed\n"); return 1; }
    cudaMemset(table, 1, n_buckets * 32);
    for (int mode = 0; mode < 2; mode++)
      for (int kind = 0; kind < 3; kind++)
        for (int g : {0, 4, 8}) {
          cudaMemset(d_cyc, 0, 8); cudaMemset(d_g, 0, 8);
          cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
          for (int rep = 0; rep < 2; rep++) {   // the second run is the measured one
            cudaMemset(d_cyc, 0, 8); cudaMemset(d_g, 0, 8);
            cudaEventRecord(e0);
            if (mode == 0) interfere<0><<<sms, WARPS * 32, smem>>>(table, n_buckets, g, kind, rounds, d_cyc, d_g, d_sink);
            else interfere<1><<<sms, WARPS * 32, smem>>>(table, n_buckets, g, kind, rounds, d_cyc, d_g, d_sink);
            cudaEventRecord(e1);
            if (cudaEventSynchronize(e1) != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
          }
          float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
          unsigned long long cyc = 0, gat = 0;
          cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost); cudaMemcpy(&gat, d_g, 8, cudaMemcpyDeviceToHost);
          printf("table %-13s gathers %-22s workers: %-25s G=%d  kernel %8.3f ms  worker cycles/round %7.1f  gather rate %6.2f G/s\n",
                 table_name[ti], mode == 0 ? "cp.async.ca -> smem" : "ld.global.nc -> regs", kind_name[kind], g, ms,
                 (double)cyc / rounds, gat / (ms * 1e-3) / 1e9);
          if (mode == 1 && g == 0) {}   // (G = 0 rows of the two modes are the same baseline)
        }
    cudaFree(table);
  }
  return 0;
}
