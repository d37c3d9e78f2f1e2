// probe_microbench3.cu -- how many DRAM bytes does one random 32-byte gather cost on B200, and which access path
// (load flavour, prefetch-size hint, texture, cp.async, TMA bulk copy, allocation kind, gather width) changes it?
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/probe_microbench3.bin tools/probe_microbench3.cu -lcuda
//   ./probe_microbench3.bin [table GB] [per_thread]
// Run plain for rates; run under `ncu --metrics dram__bytes_read.sum,lts__t_sectors_srcunit_tex_op_read.sum` for bytes.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t mix(uint64_t x) {
  x += 0x9e3779b97f4a7c15ull; x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
  x = (x ^ (x >> 27)) * 0x94d049bb133111ebull; return x ^ (x >> 31);
}
enum { LDG = 0, L2_64, L2_128, L2_256, EVICT_FIRST, NO_ALLOC_POLICY, LD8, V4, NMODES };
static const char* mode_name[] = {"2 x ld.global.nc.v2.u64", "ld.global.nc.L2::64B", "ld.global.nc.L2::128B", "ld.global.nc.L2::256B",
                                  "ld.global.nc.L2::cache_hint evict_first", "ld.global.nc.L2::cache_hint evict_unchanged",
                                  "one 8-byte ld.global.nc.u64", "ld.global.v4.u64"};

template <int MODE> __device__ __forceinline__ uint64_t ld32(const ulonglong2* p, uint64_t pol) {
  ulonglong2 a, b;
  if (MODE == LDG) { a = __ldg(p); b = __ldg(p + 1); }
  else if (MODE == L2_64) {
    asm volatile("ld.global.nc.L2::64B.v2.u64 {%0,%1}, [%2];" : "=l"(a.x), "=l"(a.y) : "l"(p));
    asm volatile("ld.global.nc.L2::64B.v2.u64 {%0,%1}, [%2];" : "=l"(b.x), "=l"(b.y) : "l"(p + 1));
  } else if (MODE == L2_128) {
    asm volatile("ld.global.nc.L2::128B.v2.u64 {%0,%1}, [%2];" : "=l"(a.x), "=l"(a.y) : "l"(p));
    asm volatile("ld.global.nc.L2::128B.v2.u64 {%0,%1}, [%2];" : "=l"(b.x), "=l"(b.y) : "l"(p + 1));
  } else if (MODE == L2_256) {
    asm volatile("ld.global.nc.L2::256B.v2.u64 {%0,%1}, [%2];" : "=l"(a.x), "=l"(a.y) : "l"(p));
    asm volatile("ld.global.nc.L2::256B.v2.u64 {%0,%1}, [%2];" : "=l"(b.x), "=l"(b.y) : "l"(p + 1));
  } else if (MODE == EVICT_FIRST || MODE == NO_ALLOC_POLICY) {
    asm volatile("ld.global.nc.L2::cache_hint.v2.u64 {%0,%1}, [%2], %3;" : "=l"(a.x), "=l"(a.y) : "l"(p), "l"(pol));
    asm volatile("ld.global.nc.L2::cache_hint.v2.u64 {%0,%1}, [%2], %3;" : "=l"(b.x), "=l"(b.y) : "l"(p + 1), "l"(pol));
  } else if (MODE == LD8) {
    a.x = __ldg(reinterpret_cast<const unsigned long long*>(p)); a.y = 0; b.x = 0; b.y = 0;
  } else {
    asm volatile("ld.global.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a.x), "=l"(a.y), "=l"(b.x), "=l"(b.y) : "l"(p));
  }
  return a.x ^ a.y ^ b.x ^ b.y;
}
template <int MODE>
__global__ void gather(const ulonglong2* table, uint64_t n_buckets, uint64_t per_thread, uint64_t* out) {
  uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, acc = 0, pol = 0;
  if (MODE == EVICT_FIRST) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  if (MODE == NO_ALLOC_POLICY) asm volatile("createpolicy.fractional.L2::evict_unchanged.b64 %0, 1.0;" : "=l"(pol));
  for (uint64_t i = 0; i < per_thread; i++) {
    uint64_t bkt = __umul64hi(mix(t * per_thread + i), n_buckets);
    acc += ld32<MODE>(table + bkt * 2, pol);
  }
  if (acc == 0x1234567) out[0] = acc;
}
// WIDTH-byte gathers (64 / 128), each done by WIDTH/16 cooperating lanes: one coalesced request per gather
template <int WIDTH>
__global__ void gather_wide(const ulonglong2* table, uint64_t n_buckets, uint64_t per_thread, uint64_t* out) {
  constexpr int LANES = WIDTH / 16;
  uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, acc = 0;
  uint64_t grp = t / LANES, sub = t % LANES, nb = n_buckets / (WIDTH / 32);
  for (uint64_t i = 0; i < per_thread; i++) {   // per_thread gathers per GROUP
    uint64_t bkt = __umul64hi(mix(grp * per_thread + i), nb);
    ulonglong2 a = __ldg(table + bkt * LANES + sub);
    acc += a.x ^ a.y;
  }
  if (acc == 0x1234567) out[0] = acc;
}
// texture path
__global__ void gather_tex(cudaTextureObject_t tex, uint64_t n_buckets, uint64_t per_thread, uint64_t* out) {
  uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; uint32_t acc = 0;
  for (uint64_t i = 0; i < per_thread; i++) {
    uint64_t bkt = __umul64hi(mix(t * per_thread + i), n_buckets);
    uint4 a = tex1Dfetch<uint4>(tex, (int)(bkt * 2)), b = tex1Dfetch<uint4>(tex, (int)(bkt * 2 + 1));
    acc += a.x ^ a.y ^ a.z ^ a.w ^ b.x ^ b.y ^ b.z ^ b.w;
  }
  if (acc == 0x1234567) out[0] = acc;
}
// cp.async (LDGSTS) 16 bytes x 2 into shared memory, DEPTH gathers in flight per thread
template <int DEPTH>
__global__ void gather_cpasync(const ulonglong2* table, uint64_t n_buckets, uint64_t per_thread, uint64_t* out) {
  __shared__ ulonglong2 buf[DEPTH][2][128];
  uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, acc = 0;
  for (uint64_t i = 0; i < per_thread; i += DEPTH) {
#pragma unroll
    for (int d = 0; d < DEPTH; d++) {
      uint64_t bkt = __umul64hi(mix(t * per_thread + i + d), n_buckets);
      uint32_t s0 = (uint32_t)__cvta_generic_to_shared(&buf[d][0][threadIdx.x]), s1 = (uint32_t)__cvta_generic_to_shared(&buf[d][1][threadIdx.x]);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s0), "l"(table + bkt * 2));
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s1), "l"(table + bkt * 2 + 1));
    }
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
#pragma unroll
    for (int d = 0; d < DEPTH; d++) { ulonglong2 a = buf[d][0][threadIdx.x], b = buf[d][1][threadIdx.x]; acc += a.x ^ a.y ^ b.x ^ b.y; }
  }
  if (acc == 0x1234567) out[0] = acc;
}
// TMA 1-D bulk copy of 32 bytes per gather (global -> shared, mbarrier completion); every thread issues its own copies
template <int DEPTH>
__global__ void gather_bulk(const ulonglong2* table, uint64_t n_buckets, uint64_t per_thread, uint64_t* out) {
  __shared__ __align__(128) ulonglong2 buf[DEPTH][128][2];
  __shared__ __align__(8) uint64_t bar;
  uint32_t bar_s = (uint32_t)__cvta_generic_to_shared(&bar);
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_s), "r"(128)); }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, acc = 0;
  uint32_t phase = 0;
  for (uint64_t i = 0; i < per_thread; i += DEPTH) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_s), "r"(32 * DEPTH) : "memory");
#pragma unroll
    for (int d = 0; d < DEPTH; d++) {
      uint64_t bkt = __umul64hi(mix(t * per_thread + i + d), n_buckets);
      uint32_t s0 = (uint32_t)__cvta_generic_to_shared(&buf[d][threadIdx.x][0]);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 32, [%2];" ::"r"(s0), "l"(table + bkt * 2), "r"(bar_s) : "memory");
    }
    uint32_t done = 0;
    while (!done) {
      asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(bar_s), "r"(phase) : "memory");
    }
    phase ^= 1;
#pragma unroll
    for (int d = 0; d < DEPTH; d++) { ulonglong2 a = buf[d][threadIdx.x][0], b = buf[d][threadIdx.x][1]; acc += a.x ^ a.y ^ b.x ^ b.y; }
    __syncthreads();   // nobody re-arms the barrier before everybody has read
  }
  if (acc == 0x1234567) out[0] = acc;
}
// prefetch.global.L2 of DEPTH buckets, then the loads (the shape of the classify kernel's drain)
template <int DEPTH>
__global__ void gather_prefetch(const ulonglong2* table, uint64_t n_buckets, uint64_t per_thread, uint64_t* out) {
  uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, acc = 0;
  for (uint64_t i = 0; i < per_thread; i += DEPTH) {
#pragma unroll
    for (int d = 0; d < DEPTH; d++) {
      uint64_t bkt = __umul64hi(mix(t * per_thread + i + d), n_buckets);
      asm volatile("prefetch.global.L2 [%0];" ::"l"(table + bkt * 2));
    }
#pragma unroll
    for (int d = 0; d < DEPTH; d++) {
      uint64_t bkt = __umul64hi(mix(t * per_thread + i + d), n_buckets);
      ulonglong2 a = __ldg(table + bkt * 2), b = __ldg(table + bkt * 2 + 1);
      acc += a.x ^ a.y ^ b.x ^ b.y;
    }
  }
  if (acc == 0x1234567) out[0] = acc;
}
// ILP: DEPTH independent gathers in flight per thread
template <int DEPTH>
__global__ void gather_ilp(const ulonglong2* table, uint64_t n_buckets, uint64_t per_thread, uint64_t* out) {
  uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, acc = 0;
  for (uint64_t i = 0; i < per_thread; i += DEPTH) {
    ulonglong2 a[DEPTH], b[DEPTH];
#pragma unroll
    for (int d = 0; d < DEPTH; d++) {
      uint64_t bkt = __umul64hi(mix(t * per_thread + i + d), n_buckets);
      a[d] = __ldg(table + bkt * 2); b[d] = __ldg(table + bkt * 2 + 1);
    }
#pragma unroll
    for (int d = 0; d < DEPTH; d++) acc += a[d].x ^ a[d].y ^ b[d].x ^ b[d].y;
  }
  if (acc == 0x1234567) out[0] = acc;
}

static int g_sms = 148;
template <class F> static void timeit(const char* name, int threads_per_sm, double gathers_per_thread_scale, uint64_t per, F launch) {
  int grid = g_sms * threads_per_sm / 128;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  launch(grid, (uint64_t)32);
  cudaEventRecord(e0); launch(grid, per); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  double n = (double)grid * 128 * per * gathers_per_thread_scale;
  printf("%-52s tps %4d : %7.2f G gathers/s (%.2f ms, %.0f M gathers) %s\n", name, threads_per_sm, n / ms / 1e6, ms, n / 1e6,
         cudaGetErrorString(cudaGetLastError()));
}
#define RUN_MODE(M) timeit(mode_name[M], 512, 1.0, per, [&](int g, uint64_t p) { gather<M><<<g, 128>>>(table, nb, p, out); })

int main(int argc, char** argv) {
  double gb = argc > 1 ? atof(argv[1]) : 16.0;
  uint64_t per = argc > 2 ? strtoull(argv[2], 0, 10) : 1024;
  cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, 0);
  uint64_t nb = (uint64_t)(gb * 1e9 / 32) & ~3ull;
  ulonglong2* table; uint64_t* out;
  cudaMalloc(&table, nb * 32); cudaMalloc(&out, 8); cudaMemset(table, 1, nb * 32);
  printf("== cudaMalloc table %.1f GB, %d SMs\n", gb, g_sms);
  RUN_MODE(LDG); RUN_MODE(L2_64); RUN_MODE(L2_128); RUN_MODE(L2_256); RUN_MODE(EVICT_FIRST); RUN_MODE(NO_ALLOC_POLICY); RUN_MODE(LD8); RUN_MODE(V4);
  timeit("64-byte gathers, 4 lanes x 16 B", 512, 1.0 / 4, per * 4, [&](int g, uint64_t p) { gather_wide<64><<<g, 128>>>(table, nb, p, out); });
  timeit("128-byte gathers, 8 lanes x 16 B", 512, 1.0 / 8, per * 8, [&](int g, uint64_t p) { gather_wide<128><<<g, 128>>>(table, nb, p, out); });
  timeit("128-byte gathers, 8 lanes x 16 B", 1024, 1.0 / 8, per * 8, [&](int g, uint64_t p) { gather_wide<128><<<g, 128>>>(table, nb, p, out); });
  timeit("256-byte gathers, 16 lanes x 16 B", 1024, 1.0 / 16, per * 16, [&](int g, uint64_t p) { gather_wide<256><<<g, 128>>>(table, nb, p, out); });
  {
    cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeLinear; rd.res.linear.devPtr = table;
    rd.res.linear.desc = cudaCreateChannelDesc<uint4>(); rd.res.linear.sizeInBytes = (size_t)((nb * 32 < (1ull << 31) * 16) ? nb * 32 : (1ull << 31) * 16 - 16);
    uint64_t nbt = rd.res.linear.sizeInBytes / 32; if (nbt > (1ull << 26)) nbt = nbt;  // tex1Dfetch index is int: <= 2^27 texels of 16 B = 2 GB
    if (nbt * 2 > (1ull << 27)) { nbt = (1ull << 26); rd.res.linear.sizeInBytes = nbt * 32; }
    cudaTextureDesc td = {}; td.readMode = cudaReadModeElementType;
    cudaTextureObject_t tex = 0; cudaError_t e = cudaCreateTextureObject(&tex, &rd, &td, nullptr);
    if (e == cudaSuccess) timeit("texture tex1Dfetch<uint4> x 2 (2 GB window)", 512, 1.0, per, [&](int g, uint64_t p) { gather_tex<<<g, 128>>>(tex, nbt, p, out); });
    else printf("texture object: %s\n", cudaGetErrorString(e));
    // same window with plain loads for comparison
    timeit("2 x ld.global.nc.v2.u64 (2 GB window)", 512, 1.0, per, [&](int g, uint64_t p) { gather<LDG><<<g, 128>>>(table, nbt, p, out); });
  }
  timeit("cp.async.cg 2 x 16 B, depth 1", 512, 1.0, per, [&](int g, uint64_t p) { gather_cpasync<1><<<g, 128>>>(table, nb, p, out); });
  timeit("cp.async.cg 2 x 16 B, depth 4", 512, 1.0, per, [&](int g, uint64_t p) { gather_cpasync<4><<<g, 128>>>(table, nb, p, out); });
  timeit("cp.async.bulk 32 B (TMA), depth 1", 512, 1.0, per, [&](int g, uint64_t p) { gather_bulk<1><<<g, 128>>>(table, nb, p, out); });
  timeit("cp.async.bulk 32 B (TMA), depth 4", 512, 1.0, per, [&](int g, uint64_t p) { gather_bulk<4><<<g, 128>>>(table, nb, p, out); });
  timeit("prefetch.global.L2 x 8 then loads", 512, 1.0, per, [&](int g, uint64_t p) { gather_prefetch<8><<<g, 128>>>(table, nb, p, out); });
  timeit("prefetch.global.L2 x 16 then loads", 768, 1.0, per, [&](int g, uint64_t p) { gather_prefetch<16><<<g, 128>>>(table, nb, p, out); });
  timeit("ILP 4", 512, 1.0, per, [&](int g, uint64_t p) { gather_ilp<4><<<g, 128>>>(table, nb, p, out); });
  timeit("ILP 8", 512, 1.0, per, [&](int g, uint64_t p) { gather_ilp<8><<<g, 128>>>(table, nb, p, out); });
  timeit("ILP 8", 1024, 1.0, per, [&](int g, uint64_t p) { gather_ilp<8><<<g, 128>>>(table, nb, p, out); });
  timeit("ILP 1", 2048, 1.0, per, [&](int g, uint64_t p) { gather_ilp<1><<<g, 128>>>(table, nb, p, out); });
  cudaFree(table);
  // other allocation kinds
  {
    ulonglong2* m = nullptr;
    if (cudaMallocManaged(&m, nb * 32) == cudaSuccess) {
      cudaMemLocation loc = {}; loc.type = cudaMemLocationTypeDevice; loc.id = 0;
      cudaMemAdvise(m, nb * 32, cudaMemAdviseSetPreferredLocation, loc);
      cudaMemPrefetchAsync(m, nb * 32, loc, 0, 0);
      cudaMemset(m, 1, nb * 32); cudaDeviceSynchronize();
      table = m;
      printf("== cudaMallocManaged (prefetched to the device)\n");
      RUN_MODE(LDG);
      cudaFree(m);
    }
  }
  {
    cuInit(0);
    CUmemAllocationProp prop = {}; prop.type = CU_MEM_ALLOCATION_TYPE_PINNED; prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE; prop.location.id = 0;
    size_t gran = 0; cuMemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED);
    size_t sz = ((nb * 32 + gran - 1) / gran) * gran;
    CUmemGenericAllocationHandle h; CUdeviceptr va = 0;
    if (cuMemCreate(&h, sz, &prop, 0) == CUDA_SUCCESS && cuMemAddressReserve(&va, sz, 0, 0, 0) == CUDA_SUCCESS && cuMemMap(va, sz, 0, h, 0) == CUDA_SUCCESS) {
      CUmemAccessDesc ad = {}; ad.location = prop.location; ad.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
      cuMemSetAccess(va, sz, &ad, 1);
      table = (ulonglong2*)va; cudaMemset(table, 1, nb * 32);
      printf("== cuMemCreate (VMM, granularity %zu)\n", gran);
      RUN_MODE(LDG);
    } else printf("VMM allocation failed\n");
  }
  return cudaDeviceSynchronize() != cudaSuccess;
}
