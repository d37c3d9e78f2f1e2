// slk_sort.cu -- K3a: hand-written LSD radix sort of the (compressed minimizer << 16 | dense taxon) cells.
//
// A (minimizer, taxon) pair is ONE 64-bit word, so a pass moves 8 bytes per pair. Per 8-bit digit pass:
//   1. tile_hist_kernel      per-tile digit counts, written digit-major: cnt[digit][tile]            (reads 8 B/key)
//   2. exclusive scan        over the flattened digit-major array = the output position of every (digit, tile) run
//   3. scatter_kernel        stable ranking inside the tile (warp match_any + per-warp digit counters), keys staged
//                            through shared memory in digit order, then written as contiguous runs     (8 B + 8 B/key)
// = 24 B/key/pass of HBM traffic plus ~1.5 B/key of counters (a single-pass "onesweep" would need 16).
// Only the bits that distinguish minimizers have to be sorted for the LCA reduce that follows (equal keys become
// adjacent; the order of taxa inside a run is irrelevant because LCA is associative and commutative).
#include "slk_sort.h"
#include "slk_core.h"

#include <stdint.h>

#define SORT_THREADS 256
#define SORT_KPT 16                                // keys per thread
#define SORT_TILE (SORT_THREADS * SORT_KPT)        // 4096 keys = 32 KB of shared memory staging
#define SORT_WARPS (SORT_THREADS / 32)

// MIX: the sort key of a build cell is the table-line mix of its minimizer, not the cell itself
template <bool MIX> static __device__ __forceinline__ uint32_t digit_of(uint64_t k, int shift) {
  return MIX ? (slk_key_mix(k >> 16) >> shift) & 0xffu : (uint32_t)(k >> shift) & 0xffu;
}

template <bool MIX>
__global__ void __launch_bounds__(SORT_THREADS) tile_hist_kernel(const uint64_t* __restrict__ keys, uint64_t n, int shift,
                                                                uint64_t n_tiles, uint64_t* __restrict__ cnt) {
  __shared__ uint32_t h[256];
  const uint64_t tile = blockIdx.x;
  h[threadIdx.x] = 0;
  __syncthreads();
  const uint64_t base = tile * SORT_TILE;
#pragma unroll 4
  for (int i = 0; i < SORT_KPT; i++) {
    uint64_t idx = base + (uint64_t)i * SORT_THREADS + threadIdx.x;   // coalesced; counting needs no order
    if (idx < n) atomicAdd(&h[digit_of<MIX>(keys[idx], shift)], 1u);
  }
  __syncthreads();
  cnt[(uint64_t)threadIdx.x * n_tiles + tile] = h[threadIdx.x];
}

// ---- exclusive scan of a uint64 array (three-phase: block sums, recursive scan of the sums, local scan + offset)
#define SCAN_THREADS 256
#define SCAN_PER_BLOCK 2048   // 8 per thread

__global__ void __launch_bounds__(SCAN_THREADS) scan_block_sums(const uint64_t* __restrict__ a, uint64_t n, uint64_t* __restrict__ sums) {
  __shared__ uint64_t w[SCAN_THREADS / 32];
  const uint64_t base = (uint64_t)blockIdx.x * SCAN_PER_BLOCK;
  uint64_t s = 0;
  for (int i = 0; i < SCAN_PER_BLOCK / SCAN_THREADS; i++) {
    uint64_t idx = base + (uint64_t)i * SCAN_THREADS + threadIdx.x;
    if (idx < n) s += a[idx];
  }
  for (int d = 16; d > 0; d >>= 1) s += __shfl_down_sync(0xffffffffu, s, d);
  if ((threadIdx.x & 31) == 0) w[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint64_t t = 0;
    for (int i = 0; i < SCAN_THREADS / 32; i++) t += w[i];
    sums[blockIdx.x] = t;
  }
}

// in place: a[i] <- offset[block] + exclusive prefix inside the block
__global__ void __launch_bounds__(SCAN_THREADS) scan_blocks(uint64_t* __restrict__ a, uint64_t n, const uint64_t* __restrict__ offsets) {
  __shared__ uint64_t wsum[SCAN_THREADS / 32];
  const int per = SCAN_PER_BLOCK / SCAN_THREADS;
  const uint64_t base = (uint64_t)blockIdx.x * SCAN_PER_BLOCK + (uint64_t)threadIdx.x * per;   // blocked: thread owns `per` consecutive
  uint64_t v[SCAN_PER_BLOCK / SCAN_THREADS], s = 0;
#pragma unroll
  for (int i = 0; i < per; i++) { v[i] = base + i < n ? a[base + i] : 0; s += v[i]; }
  // exclusive scan of the per-thread sums across the block
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint64_t incl = s;
  for (int d = 1; d < 32; d <<= 1) {
    uint64_t t = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= (uint32_t)d) incl += t;
  }
  if (lane == 31) wsum[warp] = incl;
  __syncthreads();
  uint64_t woff = 0;
  for (uint32_t i = 0; i < warp; i++) woff += wsum[i];
  uint64_t run = (offsets ? offsets[blockIdx.x] : 0) + woff + incl - s;
#pragma unroll
  for (int i = 0; i < per; i++) {
    if (base + i < n) a[base + i] = run;
    run += v[i];
  }
}

static cudaError_t exclusive_scan_u64(uint64_t* d, uint64_t n, uint64_t* scratch, cudaStream_t stream) {
  // scratch holds the block sums of every level: ceil(n/2048) + ceil(.../2048) + ... + 1 entries
  if (n == 0) return cudaSuccess;
  uint64_t nb = (n + SCAN_PER_BLOCK - 1) / SCAN_PER_BLOCK;
  if (nb == 1) {
    scan_blocks<<<1, SCAN_THREADS, 0, stream>>>(d, n, nullptr);
    return cudaGetLastError();
  }
  scan_block_sums<<<(unsigned)nb, SCAN_THREADS, 0, stream>>>(d, n, scratch);
  cudaError_t e = exclusive_scan_u64(scratch, nb, scratch + nb, stream);
  if (e != cudaSuccess) return e;
  scan_blocks<<<(unsigned)nb, SCAN_THREADS, 0, stream>>>(d, n, scratch);
  return cudaGetLastError();
}

// ---- stable scatter of one tile
template <bool MIX>
__global__ void __launch_bounds__(SORT_THREADS) scatter_kernel(const uint64_t* __restrict__ in, uint64_t* __restrict__ out, uint64_t n,
                                                              int shift, uint64_t n_tiles, const uint64_t* __restrict__ pos) {
  __shared__ uint64_t stage[SORT_TILE];
  __shared__ uint32_t wcnt[SORT_WARPS][256];   // per-warp digit counts -> exclusive prefix over warps
  __shared__ uint32_t dbase[256];              // first slot of every digit inside the tile (exclusive scan over digits)
  __shared__ uint64_t gpos[256];               // first output position of the tile's run of every digit
  const uint64_t tile = blockIdx.x;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < SORT_WARPS * 256; i += SORT_THREADS) (&wcnt[0][0])[i] = 0;
  __syncthreads();

  // Phase 1: warp w owns the contiguous keys [w*KPW, (w+1)*KPW) of the tile, lane l takes key r*32+l of round r, so
  // (warp, round, lane) is the input order. rank = keys of the same digit seen earlier in the warp.
  const uint64_t wbase = tile * SORT_TILE + (uint64_t)warp * (32 * SORT_KPT);
  uint64_t key[SORT_KPT];
  uint32_t rank[SORT_KPT];
#pragma unroll
  for (int r = 0; r < SORT_KPT; r++) {
    const uint64_t idx = wbase + (uint64_t)r * 32 + lane;
    const bool valid = idx < n;
    key[r] = valid ? in[idx] : ~0ull;
    const uint32_t d = valid ? digit_of<MIX>(key[r], shift) : 256u;       // 256: not a digit, groups the tail lanes
    const uint32_t peers = __match_any_sync(0xffffffffu, d);
    const uint32_t leader = __ffs(peers) - 1;
    uint32_t before = 0;
    if (valid && lane == leader) { before = wcnt[warp][d]; wcnt[warp][d] = before + __popc(peers); }
    before = __shfl_sync(0xffffffffu, before, leader);
    rank[r] = before + __popc(peers & ((1u << lane) - 1u));
    __syncwarp();
  }
  __syncthreads();

  // Phase 2: thread d owns digit d: exclusive prefix of the warps' counts, tile total, output position of the run
  {
    const uint32_t d = threadIdx.x;
    uint32_t run = 0;
#pragma unroll
    for (int w = 0; w < SORT_WARPS; w++) { uint32_t c = wcnt[w][d]; wcnt[w][d] = run; run += c; }
    dbase[d] = run;                       // tile total for now
    gpos[d] = pos[(uint64_t)d * n_tiles + tile];
  }
  __syncthreads();
  if (warp == 0) {                        // exclusive scan of the 256 totals by one warp, 8 digits per lane
    uint32_t v[8], s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { v[i] = dbase[lane * 8 + i]; s += v[i]; }
    uint32_t incl = s;
    for (int dd = 1; dd < 32; dd <<= 1) {
      uint32_t t = __shfl_up_sync(0xffffffffu, incl, dd);
      if (lane >= (uint32_t)dd) incl += t;
    }
    uint32_t run = incl - s;
#pragma unroll
    for (int i = 0; i < 8; i++) { dbase[lane * 8 + i] = run; run += v[i]; }
  }
  __syncthreads();

  // Phase 3: keys into shared memory in digit order, then out as runs
#pragma unroll
  for (int r = 0; r < SORT_KPT; r++) {
    const uint64_t idx = wbase + (uint64_t)r * 32 + lane;
    if (idx < n) {
      const uint32_t d = digit_of<MIX>(key[r], shift);
      stage[dbase[d] + wcnt[warp][d] + rank[r]] = key[r];
    }
  }
  __syncthreads();
  const uint64_t tile_n = (tile + 1) * SORT_TILE <= n ? SORT_TILE : n - tile * SORT_TILE;
#pragma unroll 4
  for (int i = 0; i < SORT_KPT; i++) {
    const uint32_t s = i * SORT_THREADS + threadIdx.x;
    if (s < tile_n) {
      const uint64_t k = stage[s];
      const uint32_t d = digit_of<MIX>(k, shift);
      out[gpos[d] + (s - dbase[d])] = k;
    }
  }
}

template <bool MIX>
static int sort_impl(uint64_t* keys, uint64_t* tmp, uint64_t n, int begin_bit, int end_bit, cudaStream_t stream,
                     uint64_t** sorted) {
  *sorted = keys;
  if (n <= 1 || end_bit <= begin_bit) return 0;
  const uint64_t n_tiles = (n + SORT_TILE - 1) / SORT_TILE;
  const uint64_t m = 256 * n_tiles;                       // digit-major counters
  uint64_t scratch_len = 0;
  for (uint64_t x = (m + SCAN_PER_BLOCK - 1) / SCAN_PER_BLOCK; ; x = (x + SCAN_PER_BLOCK - 1) / SCAN_PER_BLOCK) {
    scratch_len += x;
    if (x <= 1) break;
  }
  uint64_t* cnt = nullptr;
  cudaError_t e = cudaMalloc(&cnt, (m + scratch_len + 8) * sizeof(uint64_t));
  if (e != cudaSuccess) return (int)e;
  uint64_t* src = keys;
  uint64_t* dst = tmp;
  for (int shift = begin_bit; shift < end_bit; shift += 8) {
    tile_hist_kernel<MIX><<<(unsigned)n_tiles, SORT_THREADS, 0, stream>>>(src, n, shift, n_tiles, cnt);
    e = cudaGetLastError();
    if (e == cudaSuccess) e = exclusive_scan_u64(cnt, m, cnt + m, stream);
    if (e != cudaSuccess) break;
    scatter_kernel<MIX><<<(unsigned)n_tiles, SORT_THREADS, 0, stream>>>(src, dst, n, shift, n_tiles, cnt);
    e = cudaGetLastError();
    if (e != cudaSuccess) break;
    uint64_t* t = src; src = dst; dst = t;
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
  cudaFree(cnt);
  if (e != cudaSuccess) return (int)e;
  *sorted = src;
  return 0;
}
int slk_sort_u64(uint64_t* keys, uint64_t* tmp, uint64_t n, int begin_bit, int end_bit, cudaStream_t stream,
                 uint64_t** sorted) {
  return sort_impl<false>(keys, tmp, n, begin_bit, end_bit, stream, sorted);
}
int slk_sort_cells_by_line(uint64_t* cells, uint64_t* tmp, uint64_t n, cudaStream_t stream, uint64_t** sorted) {
  return sort_impl<true>(cells, tmp, n, 0, 32, stream, sorted);
}

// the same without allocation or synchronisation: scratch holds slk_scan_scratch_words(n) words
uint64_t slk_scan_scratch_words(uint64_t n) {
  uint64_t w = 8;
  for (uint64_t x = (n + SCAN_PER_BLOCK - 1) / SCAN_PER_BLOCK; ; x = (x + SCAN_PER_BLOCK - 1) / SCAN_PER_BLOCK) {
    w += x;
    if (x <= 1) break;
  }
  return w;
}
int slk_exclusive_scan_u64_async(uint64_t* d, uint64_t n, uint64_t* scratch, cudaStream_t stream) {
  return (int)exclusive_scan_u64(d, n, scratch, stream);
}
int slk_exclusive_scan_u64(uint64_t* d, uint64_t n, cudaStream_t stream) {
  if (n == 0) return 0;
  uint64_t scratch_len = 0;
  for (uint64_t x = (n + SCAN_PER_BLOCK - 1) / SCAN_PER_BLOCK; ; x = (x + SCAN_PER_BLOCK - 1) / SCAN_PER_BLOCK) {
    scratch_len += x;
    if (x <= 1) break;
  }
  uint64_t* scratch = nullptr;
  cudaError_t e = cudaMalloc(&scratch, (scratch_len + 8) * sizeof(uint64_t));
  if (e != cudaSuccess) return (int)e;
  e = exclusive_scan_u64(d, n, scratch, stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
  cudaFree(scratch);
  return (int)e;
}
