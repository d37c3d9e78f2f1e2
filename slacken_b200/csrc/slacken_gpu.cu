// slacken_gpu.cu -- sm_100a kernels and the C ABI (include/slacken_gpu.h) of the Slacken classify/build hot path.
// Kernel bodies live in slk_core.h; this file holds the __global__ wrappers, the host-side orchestration
// (streams, pinned staging, chunked pipelines, dense taxonomy) and the exported entry points.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <new>
#include <string>
#include <type_traits>
#include <unordered_map>
#include <vector>

#include "../../include/slacken_gpu.h"
#include "slk_host.h"
#include "slk_kernels.cuh"
#include "slk_sort.h"

static_assert(sizeof(slk_hit) == 8, "slk_hit layout");
static_assert(sizeof(slk_read_detail) == 24, "slk_read_detail layout");

// ---------------------------------------------------------------------------------------------- errors
static thread_local std::string g_err;
static int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}
#define CU(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
      return fail(e_ == cudaErrorMemoryAllocation ? SLK_E_NOMEM : SLK_E_CUDA, "%s failed: %s (%s:%d)", #call, \
                  cudaGetErrorString(e_), __FILE__, __LINE__);                                     \
  } while (0)
#define TRY(call)            \
  do {                       \
    int r_ = (call);         \
    if (r_ != SLK_OK) return r_; \
  } while (0)

extern "C" const char* slk_last_error(void) { return g_err.c_str(); }
int slk_fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

// ---------------------------------------------------------------------------------------------- handles
struct slk_builder {
  slk_ctx* ctx;
  slk_tax* tax;
  slk_params params;
  slk_scan_params sp;
  dense_tax dt;
  uint64_t* cells = nullptr;  // emitted (compressed key << 16 | dense taxon)
  uint64_t cap = 0;
  unsigned long long* d_count = nullptr;
  uint64_t count = 0;
  uint64_t launches = 0;
  // slk_build_reduce: the reduced cells, owner after owner, seg_count[r] of them for owner r
  uint64_t* red = nullptr;
  uint64_t* red_aux = nullptr;
  std::vector<uint64_t> seg_count;
  // per-batch staging, grow-only (a cudaMalloc / cudaFree pair per batch costs more than the batch's copies)
  uint8_t* st_frag = nullptr; size_t st_frag_bytes = 0;     // dense taxon + item prefix of the batch's fragments
  uint8_t* st_bases = nullptr; size_t st_bases_bytes = 0;   // slk_build_add: the batch's bases + fragment offsets
};
struct slk_counts {
  slk_ctx* ctx;
  int32_t n_samples, n_taxa;
  unsigned long long* d;  // [n_samples x n_taxa]
};

// ---------------------------------------------------------------------------------------------- params
static int make_scan_params(const slk_params* p, slk_scan_params* sp);
int slk_make_scan_params_checked(const slk_params* p, slk_scan_params* sp) { return make_scan_params(p, sp); }
static int make_scan_params(const slk_params* p, slk_scan_params* sp) {
  memset(sp, 0, sizeof(*sp));
  switch (slk_make_scan_params(p->k, p->m, p->spaces, p->toggle_mask, p->canonical, sp)) {
    case 0: return SLK_OK;
    case 1: return fail(SLK_E_UNSUPPORTED, "minimizer width m=%d outside 1..31", p->m);
    case 2: return fail(SLK_E_UNSUPPORTED, "k=%d, m=%d: need m <= k and k-m+1 <= %d", p->k, p->m, SLK_MAX_W);
    case 3: return fail(SLK_E_INVALID, "spaces=%d outside 0..m/2", p->spaces);
    default:
      return fail(SLK_E_UNSUPPORTED, "minimizers with %d significant bits do not fit the 48-bit key of a compact cell",
                  sp->key_bits);
  }
}

extern "C" int slk_params_init(int k, int m, int spaces, uint64_t toggle_mask, int canonical, slk_params* out) {
  if (!out) return fail(SLK_E_INVALID, "null out");
  slk_params p;
  p.k = k; p.m = m; p.spaces = spaces; p.canonical = canonical ? 1 : 0; p.toggle_mask = toggle_mask;
  slk_scan_params sp;
  TRY(make_scan_params(&p, &sp));
  *out = p;
  return SLK_OK;
}

// ---------------------------------------------------------------------------------------------- context
extern "C" int slk_ctx_create(int device, slk_ctx** out) {
  if (!out) return fail(SLK_E_INVALID, "null out");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return fail(SLK_E_CUDA, "no CUDA device available (%s); libslacken_gpu has no CPU fallback",
                e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
  if (device < 0 || device >= n) return fail(SLK_E_INVALID, "device %d out of range (0..%d)", device, n - 1);
  CU(cudaSetDevice(device));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  slk_ctx* c = new (std::nothrow) slk_ctx;
  if (!c) return fail(SLK_E_NOMEM, "host allocation failed");
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  CU(cudaStreamCreateWithFlags(&c->scan_stream, cudaStreamNonBlocking));
  // The probes are random 32-byte sectors: ask L2 to fetch single sectors from HBM instead of the default 64 bytes.
  if (!getenv("SLK_L2_FETCH_DEFAULT")) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, 32);
  cudaGetLastError();
  *out = c;
  return SLK_OK;
}
extern "C" void slk_ctx_destroy(slk_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamDestroy(c->stream);
  cudaStreamDestroy(c->scan_stream);
  cudaFree(c->span_scratch); cudaFree(c->d_maxlen);
  delete c;
}
extern "C" int slk_ctx_device(const slk_ctx* c) { return c ? c->device : -1; }
extern "C" int slk_ctx_sync(slk_ctx* c) {
  CU(cudaSetDevice(c->device));
  CU(cudaDeviceSynchronize());
  return SLK_OK;
}
extern "C" int slk_host_alloc(size_t bytes, void** out) {
  CU(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault));
  return SLK_OK;
}
extern "C" void slk_host_free(void* p) { if (p) cudaFreeHost(p); }
extern "C" int slk_host_register(void* p, size_t bytes) {
  CU(cudaHostRegister(p, bytes, cudaHostRegisterDefault));
  return SLK_OK;
}
extern "C" int slk_host_unregister(void* p) {
  CU(cudaHostUnregister(p));
  return SLK_OK;
}
extern "C" int slk_dev_alloc(slk_ctx* c, size_t bytes, void** out) {
  CU(cudaSetDevice(c->device));
  CU(cudaMalloc(out, bytes + 16));
  return SLK_OK;
}
extern "C" void slk_dev_free(slk_ctx* c, void* p) {
  if (!p) return;
  cudaSetDevice(c->device);
  cudaFree(p);
}
extern "C" int slk_memcpy_h2d(slk_ctx* c, void* dst, const void* src, size_t bytes) {
  CU(cudaSetDevice(c->device));
  CU(cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice));
  return SLK_OK;
}
extern "C" int slk_memcpy_d2h(slk_ctx* c, void* dst, const void* src, size_t bytes) {
  CU(cudaSetDevice(c->device));
  CU(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
  return SLK_OK;
}

// ---------------------------------------------------------------------------------------------- taxonomy
extern "C" int slk_taxonomy_create(slk_ctx* ctx, const int32_t* parents, int32_t n, slk_tax** out) {
  if (!ctx || !parents || n < 2 || !out) return fail(SLK_E_INVALID, "bad taxonomy arguments");
  for (int32_t i = 0; i < n; i++)
    if (parents[i] < 0 || parents[i] >= n) return fail(SLK_E_INVALID, "parents[%d]=%d out of range", i, parents[i]);
  slk_tax* t = new (std::nothrow) slk_tax;
  if (!t) return fail(SLK_E_NOMEM, "host allocation failed");
  t->ctx = ctx;
  t->parents.assign(parents, parents + n);
  t->parents[1] = 0;  // parents(ROOT) = NONE (slacken/Taxonomy.scala:105)
  t->parents[0] = 0;
  *out = t;
  return SLK_OK;
}
extern "C" void slk_taxonomy_destroy(slk_tax* t) { delete t; }

// dense, ancestor-closed numbering of the taxa a library touches; dense 0 = NONE, ROOT is always present
static void dense_init(dense_tax& dt) {
  dt.raw.assign(1, 0); dt.parent.assign(1, 0); dt.depth.assign(1, 0);
  dt.to_dense.clear(); dt.to_dense[0] = 0;
}
static int dense_add(dense_tax& dt, const slk_tax* tax, int32_t raw, uint32_t* out) {
  auto it = dt.to_dense.find(raw);
  if (it != dt.to_dense.end()) { if (out) *out = it->second; return SLK_OK; }
  // walk up until a known node, then number the path top-down
  std::vector<int32_t> path;
  int32_t t = raw;
  while (dt.to_dense.find(t) == dt.to_dense.end()) {
    path.push_back(t);
    if (path.size() > 250) return fail(SLK_E_UNSUPPORTED, "taxon %d is deeper than 250 levels (cycle in parents?)", raw);
    t = tax->parents[t];
  }
  uint32_t pd = dt.to_dense[t];
  for (size_t i = path.size(); i-- > 0;) {
    if (dt.raw.size() >= 65535) return fail(SLK_E_UNSUPPORTED, "library touches more than 65534 taxonomy nodes (compact 16-bit cells)");
    if (dt.depth[pd] >= 254) return fail(SLK_E_UNSUPPORTED, "taxonomy deeper than 254 levels");
    uint32_t id = (uint32_t)dt.raw.size();
    dt.raw.push_back(path[i]); dt.parent.push_back((uint16_t)pd); dt.depth.push_back((uint8_t)(dt.depth[pd] + 1));
    dt.to_dense[path[i]] = id;
    pd = id;
  }
  if (out) *out = pd;
  return SLK_OK;
}
static int dense_upload(dense_tax& dt) {
  size_t n = dt.raw.size();
  cudaFree(dt.d_parent); cudaFree(dt.d_depth); cudaFree(dt.d_raw);
  dt.d_parent = nullptr; dt.d_depth = nullptr; dt.d_raw = nullptr;
  CU(cudaMalloc(&dt.d_parent, n * sizeof(uint16_t)));
  CU(cudaMalloc(&dt.d_depth, n));
  CU(cudaMalloc(&dt.d_raw, n * sizeof(int32_t)));
  CU(cudaMemcpy(dt.d_parent, dt.parent.data(), n * sizeof(uint16_t), cudaMemcpyHostToDevice));
  CU(cudaMemcpy(dt.d_depth, dt.depth.data(), n, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(dt.d_raw, dt.raw.data(), n * sizeof(int32_t), cudaMemcpyHostToDevice));
  return SLK_OK;
}
static void dense_free(dense_tax& dt) {
  cudaFree(dt.d_parent); cudaFree(dt.d_depth); cudaFree(dt.d_raw);
  dt.d_parent = nullptr; dt.d_depth = nullptr; dt.d_raw = nullptr;
}
void slk_dense_init(dense_tax& dt) { dense_init(dt); }
int slk_dense_add(dense_tax& dt, const slk_tax* tax, int32_t raw, uint32_t* out) { return dense_add(dt, tax, raw, out); }
int slk_dense_upload(dense_tax& dt) { return dense_upload(dt); }
void slk_dense_free(dense_tax& dt) { dense_free(dt); }

// ---------------------------------------------------------------------------------------------- table kernels
// K4: insert (compressed key << 16 | dense taxon) cells. Equal keys merge by LCA (TaxonLCA.merge,
// slacken/LowestCommonAncestor.scala:152-170), so the kernel also serves incremental builds.
// one thread's insert; returns true when the key was new
// Returns true when the key is new. A table without a free cell on the whole probe sequence (an undersized table: cannot
// happen with table_alloc's sizing, but would silently lose the record) raises bit 63 of *full.
#define SLK_TABLE_FULL_BIT (1ull << 63)
__device__ __forceinline__ bool insert_cell(uint64_t cell, const slk_table_view& tb, const slk_tax_view& tx, unsigned long long* full) {
  uint64_t ckey = cell >> 16;
  uint32_t taxon = (uint32_t)(cell & 0xffffu);
  if (taxon == 0) return false;  // a record whose taxon is NONE behaves exactly like a missing record
  uint64_t b = slk_bucket_of(ckey, tb);
  for (uint64_t tries = 1; tries <= tb.n_buckets; tries++) {
    unsigned long long* slot = reinterpret_cast<unsigned long long*>(tb.cells + b * 4);
    for (int j = 0; j < 4; j++) {
      unsigned long long cur = slot[j];
      if (cur == 0) {
        unsigned long long old = atomicCAS(&slot[j], 0ull, (unsigned long long)cell);
        if (old == 0) return true;
        cur = old;
      }
      if ((cur >> 16) == ckey) {
        for (;;) {
          uint32_t t_old = (uint32_t)(cur & 0xffffu);
          uint32_t t_new = slk_lca(tx, t_old, taxon);
          if (t_new == t_old) return false;
          unsigned long long want = (ckey << 16) | t_new;
          unsigned long long old = atomicCAS(&slot[j], cur, want);
          if (old == cur) return false;
          cur = old;
        }
      }
    }
    b = slk_next_bucket(b, tries, tb.n_buckets);   // the lookup's probe sequence: own 128-byte line first
  }
  atomicOr(full, SLK_TABLE_FULL_BIT);
  return false;
}
// map != NULL: the cells carry the dense taxa of another rank's builder; map[] translates them to this index's
__global__ void __launch_bounds__(256) insert_cells_kernel(const uint64_t* __restrict__ in, uint64_t n,
                                                           slk_table_view tb, slk_tax_view tx,
                                                           unsigned long long* n_new, const uint16_t* __restrict__ map) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t cell = i < n ? in[i] : 0;
  if (map) cell = (cell & ~0xffffull) | map[cell & 0xffffu];
  const bool fresh = i < n && insert_cell(cell, tb, tx, n_new);
  const uint32_t cnt = (uint32_t)__syncthreads_count(fresh);   // one atomic per block on the record counter
  if (threadIdx.x == 0 && cnt) atomicAdd(n_new, (unsigned long long)cnt);
}

// records (id1, raw taxon) -> cells, via the raw->dense lookup table
__global__ void __launch_bounds__(256) records_to_cells_kernel(const int64_t* __restrict__ id1,
                                                               const int32_t* __restrict__ taxon, uint64_t n,
                                                               const uint16_t* __restrict__ raw2dense,
                                                               slk_scan_params sp, uint64_t* __restrict__ out) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t key = (uint64_t)id1[i];
  out[i] = (slk_compress(sp, key) << 16) | raw2dense[taxon[i]];
}
__global__ void __launch_bounds__(256) mark_taxa_kernel(const int32_t* __restrict__ taxon, uint64_t n, int32_t n_tax,
                                                        uint32_t* __restrict__ bitmap, uint32_t* __restrict__ bad) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int32_t t = taxon[i];
  if (t < 0 || t >= n_tax) { atomicExch(bad, 1u); return; }
  uint32_t bit = 1u << (t & 31);
  if (!(bitmap[t >> 5] & bit)) atomicOr(&bitmap[t >> 5], bit);
}
// table -> records
__global__ void __launch_bounds__(256) dump_table_kernel(slk_table_view tb, slk_scan_params sp, const int32_t* __restrict__ raw,
                                                         int64_t* __restrict__ id1, int32_t* __restrict__ taxon,
                                                         uint64_t cap, unsigned long long* cursor) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t ncell = tb.n_buckets * 4;
  uint64_t cell = i < ncell ? tb.cells[i] : 0;
  uint64_t o = warp_agg_alloc(cursor, cell != 0 ? 1u : 0u);
  if (cell != 0 && o < cap) {
    id1[o] = (int64_t)slk_expand(sp, cell >> 16);
    taxon[o] = raw[cell & 0xffffu];
  }
}

// K3b: segmented LCA reduce over the sorted cells. The head of every run of equal keys folds the run's taxa (in
// whatever order the stable sort left them: LCA is associative and commutative) and appends one cell.
__global__ void __launch_bounds__(256) reduce_cells_kernel(const uint64_t* __restrict__ in, uint64_t n, slk_tax_view tx,
                                                           uint64_t* __restrict__ out, unsigned long long* cursor) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t cell = 0;
  bool head = false;
  if (i < n) {
    cell = in[i];
    head = (i == 0) || ((in[i - 1] >> 16) != (cell >> 16));
  }
  uint64_t res = 0;
  if (head) {
    uint64_t key = cell >> 16;
    uint32_t acc = (uint32_t)(cell & 0xffffu), last = acc;
    for (uint64_t j = i + 1; j < n; j++) {
      uint64_t c = in[j];
      if ((c >> 16) != key) break;
      uint32_t t = (uint32_t)(c & 0xffffu);
      if (t != last) { acc = slk_lca(tx, acc, t); last = t; }
    }
    res = (key << 16) | acc;
  }
  // one atomic per BLOCK on the output cursor (41 M same-address atomics, one per warp, were a third of this kernel)
  __shared__ uint32_t wtot[8];
  __shared__ unsigned long long bbase;
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const uint32_t bal = __ballot_sync(0xffffffffu, head);
  if (lane == 0) wtot[warp] = __popc(bal);
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
    for (int w = 0; w < 8; w++) { uint32_t c = wtot[w]; wtot[w] = t; t += c; }
    bbase = t ? atomicAdd(cursor, (unsigned long long)t) : 0ull;
  }
  __syncthreads();
  if (head) out[bbase + wtot[warp] + __popc(bal & ((1u << lane) - 1u))] = res;
}

// K1: 2-bit encode + ambiguity mask, one read per thread, into the packed block layout (slk_read_src)
// K1 as a stand-alone step: ASCII reads -> 2-bit code blocks + ambiguity masks. One 32-base block per lane, so that the
// lanes of a warp read consecutive 32-byte pieces of the input and write consecutive output words; a warp owns 32
// consecutive reads and finds the read of a block by a shuffle search over the reads' block offsets.
__global__ void __launch_bounds__(256) pack_reads_kernel(const uint8_t* __restrict__ bases, const uint64_t* __restrict__ off,
                                                         uint32_t n_reads, const uint64_t* __restrict__ boff,
                                                         uint64_t* __restrict__ codes, uint32_t* __restrict__ mask,
                                                         uint32_t* __restrict__ len_out) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t r0 = warp * 32;
  if (r0 >= n_reads) return;
  const uint32_t nvalid = (uint32_t)(n_reads - r0 < 32 ? n_reads - r0 : 32);
  const uint64_t r = r0 + (lane < nvalid ? lane : nvalid - 1);
  const uint64_t my_off = off[r], my_end = off[r + 1], my_boff = boff[r];
  if (lane < nvalid) len_out[r] = (uint32_t)(my_end - my_off);
  const uint64_t b_first = __shfl_sync(0xffffffffu, my_boff, 0);
  const uint64_t b_last = boff[r0 + nvalid];
  for (uint64_t b0 = b_first; b0 < b_last; b0 += 32) {
    const uint64_t b = b0 + lane;
    uint32_t lo = 0;   // the last read of the warp whose first block is <= b
#pragma unroll
    for (uint32_t step = 16; step >= 1; step >>= 1) {
      const uint32_t cand = lo + step;
      const uint64_t v = __shfl_sync(0xffffffffu, my_boff, cand & 31u);
      if (cand < nvalid && v <= b) lo = cand;
    }
    const uint64_t roff = __shfl_sync(0xffffffffu, my_off, lo), rend = __shfl_sync(0xffffffffu, my_end, lo);
    const uint64_t rb = __shfl_sync(0xffffffffu, my_boff, lo);
    if (b >= b_last) continue;
    const uint64_t start = roff + 32 * (b - rb);
    const uint32_t n = (uint32_t)(rend - start < 32 ? rend - start : 32);
    // 32 bytes from an arbitrary address: three aligned 16-byte words (only those that hold bytes of the read: the buffer
    // is readable up to the next 16-byte boundary, as everywhere in this library), moved into place by a word select in
    // two levels and a funnel shift, then coded four bytes at a time (slk_code4)
    const uintptr_t a0 = reinterpret_cast<uintptr_t>(bases + start);
    const uint4* vp = reinterpret_cast<const uint4*>(a0 & ~(uintptr_t)15);
    const uint32_t mis = (uint32_t)(a0 & 15);
    const uint32_t nv = (n + mis + 15) >> 4;   // aligned 16-byte words that hold the n bytes
    uint32_t w[12];
#pragma unroll
    for (int i = 0; i < 3; i++) {
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if ((uint32_t)i < nv) v = __ldg(vp + i);
      w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
    }
    const bool q1 = (mis & 4u) != 0, q2 = (mis & 8u) != 0;
    const uint32_t sh = (mis & 3u) * 8u;
    uint32_t u[11], t[9];
#pragma unroll
    for (int i = 0; i < 11; i++) u[i] = q1 ? w[i + 1] : w[i];
#pragma unroll
    for (int i = 0; i < 9; i++) t[i] = q2 ? u[i + 2] : u[i];
    uint64_t cw = 0;
    uint32_t mw = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      uint32_t c8, i4;
      slk_code4(__funnelshift_r(t[i], t[i + 1], sh), &c8, &i4);
      cw |= (uint64_t)c8 << (8 * i);
      mw |= i4 << (4 * i);
    }
    if (n < 32) { cw &= (1ull << (2 * n)) - 1ull; mw &= (1u << n) - 1u; }
    codes[b] = cw;
    mask[b] = mw;
  }
}

__global__ void snapshot_kernel(const unsigned long long* src, unsigned long long* dst) { *dst = *src; }

__global__ void __launch_bounds__(256) counts_add_kernel(const int32_t* __restrict__ taxon, const uint8_t* __restrict__ flags,
                                                         const int32_t* __restrict__ sample, uint32_t n, int32_t n_samples,
                                                         int32_t n_taxa, unsigned long long* counts, uint32_t* bad) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (!(flags[i] & SLK_READ_HAS_SPAN)) return;
  int32_t s = sample ? sample[i] : 0, t = taxon[i];
  if (s < 0 || s >= n_samples || t < 0 || t >= n_taxa) { atomicExch(bad, 1u); return; }
  atomicAdd(&counts[(uint64_t)s * n_taxa + t], 1ull);
}

// ---------------------------------------------------------------------------------------------- synthetic data kernels
__global__ void __launch_bounds__(256) synth_genome_kernel(uint64_t seed, uint64_t start, uint64_t n, uint8_t* out) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = slk_synth_genome_base(seed, start + i);
}
__global__ void __launch_bounds__(256) synth_reads_kernel(uint64_t gseed, uint64_t rseed, uint64_t n_genomes,
                                                          uint64_t genome_len, uint64_t first, uint64_t n_reads,
                                                          uint32_t L, uint32_t mate, uint8_t* out) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_reads * L) return;
  uint64_t r = i / L;
  uint32_t j = (uint32_t)(i - r * L);
  out[i] = slk_synth_read_base(gseed, rseed, n_genomes, genome_len, first + r, L, j, mate);
}

extern "C" int slk_synth_genome_dev(slk_ctx* ctx, uint64_t seed, uint64_t start, uint64_t n, uint8_t* out) {
  CU(cudaSetDevice(ctx->device));
  if (n == 0) return SLK_OK;
  synth_genome_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(seed, start, n, out);
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(ctx->stream));
  return SLK_OK;
}
extern "C" int slk_synth_mates_dev(slk_ctx* ctx, uint64_t gseed, uint64_t rseed, uint64_t n_genomes, uint64_t genome_len,
                                   uint64_t first_read, uint64_t n_reads, uint32_t read_len, uint32_t mate, uint8_t* out) {
  if (!ctx || !out) return fail(SLK_E_INVALID, "bad arguments");
  CU(cudaSetDevice(ctx->device));
  if (n_reads == 0) return SLK_OK;
  if (read_len == 0 || genome_len < read_len || mate > 1) return fail(SLK_E_INVALID, "bad read/genome length or mate");
  uint64_t n = n_reads * read_len;
  synth_reads_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(gseed, rseed, n_genomes, genome_len, first_read,
                                                                            n_reads, read_len, mate, out);
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(ctx->stream));
  return SLK_OK;
}
extern "C" int slk_synth_reads_dev(slk_ctx* ctx, uint64_t gseed, uint64_t rseed, uint64_t n_genomes, uint64_t genome_len,
                                   uint64_t first_read, uint64_t n_reads, uint32_t read_len, uint8_t* out) {
  return slk_synth_mates_dev(ctx, gseed, rseed, n_genomes, genome_len, first_read, n_reads, read_len, 0, out);
}

// ---------------------------------------------------------------------------------------------- index construction
// Load factor of the table. A lookup whose home bucket is full without a match costs a second request, and the
// B200 serves only ~25 G random line requests per second to a kernel of this kind, so emptier is faster (measured,
// 10 M reads vs the 4 Gbp library: 0.3 -> 615, 0.4 -> 604, 0.5 -> 586, 0.7 -> 477 M reads/s). Default: 0.4 while the
// table stays below a quarter of the free device memory, 0.5 below half of it, otherwise 0.7.
// SLK_TABLE_LOAD_FACTOR (0.2 .. 0.9) overrides.
static double table_load_factor(uint64_t n_keys) {
  const char* e = getenv("SLK_TABLE_LOAD_FACTOR");
  if (e) { double lf = atof(e); return lf < 0.2 ? 0.2 : lf > 0.9 ? 0.9 : lf; }
  size_t free_b = 0, total_b = 0;
  if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) return 0.5;
  const double need = 8.0 * (double)n_keys;
  if (need / 0.4 <= 0.25 * (double)free_b) return 0.4;
  if (need / 0.5 <= 0.5 * (double)free_b) return 0.5;
  return 0.7;
}
static uint64_t buckets_for(uint64_t n_keys) {
  uint64_t cells = (uint64_t)((double)n_keys / table_load_factor(n_keys)) + 64;
  return ((cells + 15) / 16) * 4;   // whole 128-byte lines of four buckets
}
static int table_alloc(slk_table_view* tb, uint64_t n_keys, uint32_t world = 1) {
  tb->n_buckets = buckets_for(n_keys);
  tb->mix_mul = world;
  tb->pad_ = 0;
  CU(cudaMalloc(&tb->cells, tb->n_buckets * 32));
  CU(cudaMemset(tb->cells, 0, tb->n_buckets * 32));
  return SLK_OK;
}
static int insert_cells(slk_ctx* ctx, slk_index* idx, const uint64_t* d_cells, uint64_t n, unsigned long long* d_new,
                        const uint16_t* d_map = nullptr) {
  if (n == 0) return SLK_OK;
  insert_cells_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(d_cells, n, idx->table, idx->dt.view(), d_new, d_map);
  CU(cudaGetLastError());
  return SLK_OK;
}

extern "C" int slk_index_from_records(slk_ctx* ctx, slk_tax* tax, const slk_params* params, const int64_t* id1,
                                      const int32_t* taxon, uint64_t n, slk_index** out) {
  return slk_index_from_records_shard(ctx, tax, params, id1, taxon, n, 1, out);
}
extern "C" int slk_index_from_records_shard(slk_ctx* ctx, slk_tax* tax, const slk_params* params, const int64_t* id1,
                                            const int32_t* taxon, uint64_t n, uint32_t world, slk_index** out) {
  if (!ctx || !tax || !params || !out || world < 1 || (n && (!id1 || !taxon))) return fail(SLK_E_INVALID, "bad arguments");
  CU(cudaSetDevice(ctx->device));
  slk_index* idx = new (std::nothrow) slk_index;
  if (!idx) return fail(SLK_E_NOMEM, "host allocation failed");
  idx->ctx = ctx; idx->tax = tax; idx->params = *params;
  int rc = make_scan_params(params, &idx->sp);
  if (rc != SLK_OK) { delete idx; return rc; }
  dense_init(idx->dt);
  int32_t n_tax = (int32_t)tax->parents.size();
  // 1) which taxa occur: device bitmap over raw ids, records streamed in chunks
  // records already in device memory (the distributed build hands over what it received from its peers) are read in
  // place, in larger pieces; host records are staged piece by piece
  cudaPointerAttributes pa_id{}, pa_tx{};
  const bool on_device = n > 0 && cudaPointerGetAttributes(&pa_id, id1) == cudaSuccess && pa_id.type == cudaMemoryTypeDevice &&
                         cudaPointerGetAttributes(&pa_tx, taxon) == cudaSuccess && pa_tx.type == cudaMemoryTypeDevice;
  cudaGetLastError();
  const uint64_t CH = on_device ? 1ull << 28 : 1ull << 26;
  uint64_t chunk = std::min<uint64_t>(n ? n : 1, CH);
  int64_t* d_id = nullptr; int32_t* d_tx = nullptr; uint64_t* d_cells = nullptr; uint64_t* d_tmp = nullptr;
  uint32_t *d_bitmap = nullptr, *d_bad = nullptr; uint16_t* d_r2d = nullptr; unsigned long long* d_new = nullptr;
  size_t words = ((size_t)n_tax + 31) / 32;
  auto cleanup = [&]() {
    if (!on_device) { cudaFree(d_id); cudaFree(d_tx); }
    cudaFree(d_cells); cudaFree(d_tmp); cudaFree(d_bitmap); cudaFree(d_bad); cudaFree(d_r2d); cudaFree(d_new);
  };
#define CUX(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cleanup(); slk_index_destroy(idx); \
    return fail(e_ == cudaErrorMemoryAllocation ? SLK_E_NOMEM : SLK_E_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); } } while (0)
  if (!on_device) { CUX(cudaMalloc(&d_id, chunk * 8)); CUX(cudaMalloc(&d_tx, chunk * 4)); }
  CUX(cudaMalloc(&d_cells, chunk * 8)); CUX(cudaMalloc(&d_tmp, chunk * 8));
  CUX(cudaMalloc(&d_bitmap, words * 4)); CUX(cudaMalloc(&d_bad, 4)); CUX(cudaMalloc(&d_new, 8));
  CUX(cudaMemset(d_bitmap, 0, words * 4)); CUX(cudaMemset(d_bad, 0, 4)); CUX(cudaMemset(d_new, 0, 8));
  for (uint64_t s = 0; s < n; s += chunk) {
    uint64_t c = std::min(chunk, n - s);
    if (on_device) d_tx = const_cast<int32_t*>(taxon + s);
    else CUX(cudaMemcpyAsync(d_tx, taxon + s, c * 4, cudaMemcpyDefault, ctx->stream));
    mark_taxa_kernel<<<(unsigned)((c + 255) / 256), 256, 0, ctx->stream>>>(d_tx, c, n_tax, d_bitmap, d_bad);
    CUX(cudaGetLastError());
    CUX(cudaStreamSynchronize(ctx->stream));
  }
  uint32_t bad = 0;
  CUX(cudaMemcpy(&bad, d_bad, 4, cudaMemcpyDeviceToHost));
  if (bad) { cleanup(); slk_index_destroy(idx); return fail(SLK_E_INVALID, "a record's taxon is outside the taxonomy (0..%d)", n_tax - 1); }
  std::vector<uint32_t> bitmap(words);
  CUX(cudaMemcpy(bitmap.data(), d_bitmap, words * 4, cudaMemcpyDeviceToHost));
  rc = dense_add(idx->dt, tax, 1, &idx->dt.root);
  for (int32_t t = 1; t < n_tax && rc == SLK_OK; t++)
    if (bitmap[t >> 5] & (1u << (t & 31))) rc = dense_add(idx->dt, tax, t, nullptr);
  if (rc != SLK_OK) { cleanup(); slk_index_destroy(idx); return rc; }
  rc = dense_upload(idx->dt);
  if (rc != SLK_OK) { cleanup(); slk_index_destroy(idx); return rc; }
  std::vector<uint16_t> r2d((size_t)n_tax, 0);
  for (auto& kv : idx->dt.to_dense) r2d[kv.first] = (uint16_t)kv.second;
  CUX(cudaMalloc(&d_r2d, (size_t)n_tax * 2));
  CUX(cudaMemcpy(d_r2d, r2d.data(), (size_t)n_tax * 2, cudaMemcpyHostToDevice));
  // 2) table + insert
  rc = table_alloc(&idx->table, n, world);
  if (rc != SLK_OK) { cleanup(); slk_index_destroy(idx); return rc; }
  for (uint64_t s = 0; s < n; s += chunk) {
    uint64_t c = std::min(chunk, n - s);
    if (on_device) { d_id = const_cast<int64_t*>(id1 + s); d_tx = const_cast<int32_t*>(taxon + s); }
    else {
      CUX(cudaMemcpyAsync(d_id, id1 + s, c * 8, cudaMemcpyDefault, ctx->stream));
      CUX(cudaMemcpyAsync(d_tx, taxon + s, c * 4, cudaMemcpyDefault, ctx->stream));
    }
    records_to_cells_kernel<<<(unsigned)((c + 255) / 256), 256, 0, ctx->stream>>>(d_id, d_tx, c, d_r2d, idx->sp, d_cells);
    CUX(cudaGetLastError());
    // in the order of the table's lines, as in slk_build_finish: each piece walks the table front to back
    uint64_t* sorted_ptr = nullptr;
    if (slk_sort_cells_by_line(d_cells, d_tmp, c, ctx->stream, &sorted_ptr) != 0) {
      cleanup(); slk_index_destroy(idx); return fail(SLK_E_CUDA, "radix sort failed");
    }
    rc = insert_cells(ctx, idx, sorted_ptr, c, d_new);
    if (rc != SLK_OK) { cleanup(); slk_index_destroy(idx); return rc; }
    CUX(cudaStreamSynchronize(ctx->stream));
  }
  unsigned long long nn = 0;
  CUX(cudaMemcpy(&nn, d_new, 8, cudaMemcpyDeviceToHost));
  if (nn & SLK_TABLE_FULL_BIT) { cleanup(); slk_index_destroy(idx); return fail(SLK_E_NOSPACE, "the minimizer table is full: records were not stored"); }
  idx->n_records = nn;
  cleanup();
#undef CUX
  *out = idx;
  return SLK_OK;
}
extern "C" void slk_index_destroy(slk_index* idx) {
  if (!idx) return;
  cudaSetDevice(idx->ctx->device);
  cudaFree(idx->table.cells);
  dense_free(idx->dt);
  delete idx;
}
extern "C" uint64_t slk_index_size(const slk_index* idx) { return idx ? idx->n_records : 0; }

extern "C" int slk_index_records(slk_index* idx, int64_t* id1_out, int32_t* taxon_out, uint64_t cap, uint64_t* n_out) {
  if (!idx || !n_out) return fail(SLK_E_INVALID, "bad arguments");
  slk_ctx* ctx = idx->ctx;
  CU(cudaSetDevice(ctx->device));
  *n_out = idx->n_records;
  if (cap < idx->n_records) return fail(SLK_E_NOSPACE, "records need room for %llu rows", (unsigned long long)idx->n_records);
  if (idx->n_records == 0) return SLK_OK;
  int64_t* d_id = nullptr; int32_t* d_tx = nullptr; unsigned long long* d_cur = nullptr;
  CU(cudaMalloc(&d_id, idx->n_records * 8));
  CU(cudaMalloc(&d_tx, idx->n_records * 4));
  CU(cudaMalloc(&d_cur, 8));
  CU(cudaMemset(d_cur, 0, 8));
  uint64_t ncell = idx->table.n_buckets * 4;
  dump_table_kernel<<<(unsigned)((ncell + 255) / 256), 256, 0, ctx->stream>>>(idx->table, idx->sp, idx->dt.d_raw, d_id, d_tx,
                                                                              idx->n_records, d_cur);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(id1_out, d_id, idx->n_records * 8, cudaMemcpyDefault, ctx->stream));
  CU(cudaMemcpyAsync(taxon_out, d_tx, idx->n_records * 4, cudaMemcpyDefault, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  cudaFree(d_id); cudaFree(d_tx); cudaFree(d_cur);
  return SLK_OK;
}

// ---------------------------------------------------------------------------------------------- builder
extern "C" int slk_build_begin(slk_ctx* ctx, slk_tax* tax, const slk_params* params, uint64_t expected_bases,
                               slk_builder** out) {
  if (!ctx || !tax || !params || !out) return fail(SLK_E_INVALID, "bad arguments");
  CU(cudaSetDevice(ctx->device));
  slk_builder* b = new (std::nothrow) slk_builder;
  if (!b) return fail(SLK_E_NOMEM, "host allocation failed");
  b->ctx = ctx; b->tax = tax; b->params = *params;
  int rc = make_scan_params(params, &b->sp);
  if (rc != SLK_OK) { delete b; return rc; }
  dense_init(b->dt);
  rc = dense_add(b->dt, tax, 1, &b->dt.root);
  if (rc != SLK_OK) { delete b; return rc; }
  b->cap = expected_bases / 2 + (1u << 20);
  cudaError_t e = cudaMalloc(&b->cells, b->cap * 8);
  if (e == cudaSuccess) e = cudaMalloc(&b->d_count, 8);
  if (e == cudaSuccess) e = cudaMemset(b->d_count, 0, 8);
  if (e != cudaSuccess) { slk_build_destroy(b); return fail(SLK_E_NOMEM, "builder allocation failed: %s", cudaGetErrorString(e)); }
  *out = b;
  return SLK_OK;
}
extern "C" void slk_build_destroy(slk_builder* b) {
  if (!b) return;
  cudaSetDevice(b->ctx->device);
  cudaFree(b->cells); cudaFree(b->d_count);
  if (b->red != b->cells) cudaFree(b->red);
  if (b->red_aux != b->cells) cudaFree(b->red_aux);
  cudaFree(b->st_frag); cudaFree(b->st_bases);
  dense_free(b->dt);
  delete b;
}
static int builder_reserve(slk_builder* b, uint64_t extra) {
  if (b->count + extra <= b->cap) return SLK_OK;
  uint64_t ncap = std::max(b->count + extra, b->cap + b->cap / 2);
  uint64_t* n = nullptr;
  CU(cudaMalloc(&n, ncap * 8));
  CU(cudaMemcpyAsync(n, b->cells, b->count * 8, cudaMemcpyDeviceToDevice, b->ctx->stream));
  CU(cudaStreamSynchronize(b->ctx->stream));
  cudaFree(b->cells);
  b->cells = n; b->cap = ncap;
  return SLK_OK;
}

#define DISPATCH_W(w, FN, ARGS)          \
  switch (w) {                           \
    case 1: FN##1(ARGS); break;          \
    case 2: FN##2(ARGS); break;          \
    case 3: FN##3(ARGS); break;          \
    case 4: FN##4(ARGS); break;          \
    case 5: FN##5(ARGS); break;          \
    case 6: FN##6(ARGS); break;          \
    case 7: FN##7(ARGS); break;          \
    default: FN##8(ARGS); break;         \
  }

// shared by the host- and device-buffer entry points; off_host: fragment offsets on the host (always needed)
static int build_add_impl(slk_builder* b, const uint8_t* d_bases, const uint64_t* d_off, uint64_t shift,
                          const uint64_t* off_host, const int32_t* taxon_host, uint32_t n_frag) {
  slk_ctx* ctx = b->ctx;
  int32_t n_tax = (int32_t)b->tax->parents.size();
  std::vector<uint32_t> dense(n_frag, 0);
  std::vector<uint64_t> prefix((size_t)n_frag + 1, 0);
  uint64_t windows = 0;
  for (uint32_t f = 0; f < n_frag; f++) {
    int32_t t = taxon_host[f];
    uint64_t len = off_host[f + 1] - off_host[f];
    uint64_t nw = len >= (uint64_t)b->sp.k ? len - b->sp.k + 1 : 0;
    // Taxonomy.isDefined (slacken/Taxonomy.scala:175-176): undefined labels are dropped (KeyValueIndex.scala:118-120)
    bool defined = t > 0 && t < n_tax && (b->tax->parents[t] != 0 || t == 1);
    if (!defined) nw = 0;
    else TRY(dense_add(b->dt, b->tax, t, &dense[f]));
    prefix[f + 1] = prefix[f] + (nw + BUILD_WPT - 1) / BUILD_WPT;
    windows += nw;
  }
  uint64_t n_items = prefix[n_frag];
  if (n_items == 0) return SLK_OK;
  TRY(builder_reserve(b, windows));
  const size_t prefix_bytes = ((size_t)n_frag + 1) * 8, need = prefix_bytes + (size_t)n_frag * 4;
  if (need > b->st_frag_bytes) {
    cudaFree(b->st_frag); b->st_frag = nullptr; b->st_frag_bytes = 0;
    CU(cudaMalloc(&b->st_frag, need + need / 2));
    b->st_frag_bytes = need + need / 2;
  }
  uint64_t* d_prefix = reinterpret_cast<uint64_t*>(b->st_frag);
  uint32_t* d_dense = reinterpret_cast<uint32_t*>(b->st_frag + prefix_bytes);
  CU(cudaMemcpyAsync(d_dense, dense.data(), (size_t)n_frag * 4, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(d_prefix, prefix.data(), ((size_t)n_frag + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
  slk_emit_args ea;
  ea.sp = b->sp; ea.bases = d_bases; ea.frag_off = d_off; ea.off_shift = shift; ea.frag_dense = d_dense;
  ea.item_prefix = d_prefix; ea.n_frag = n_frag; ea.n_items = n_items; ea.out = b->cells; ea.cap = b->cap;
  ea.cursor = b->d_count; ea.stream = ctx->stream;
  DISPATCH_W(b->sp.w, slk_launch_emit_w, ea);
  b->launches++;
  CU(cudaGetLastError());
  unsigned long long cnt = 0;
  CU(cudaMemcpyAsync(&cnt, b->d_count, 8, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  b->count = cnt;
  return SLK_OK;
}

extern "C" int slk_build_add(slk_builder* b, const uint8_t* bases, const uint64_t* frag_off, const int32_t* frag_taxon,
                             uint32_t n_frag) {
  if (!b || !bases || !frag_off || !frag_taxon) return fail(SLK_E_INVALID, "bad arguments");
  if (n_frag == 0) return SLK_OK;
  slk_ctx* ctx = b->ctx;
  CU(cudaSetDevice(ctx->device));
  uint64_t s = frag_off[0], total = frag_off[n_frag] - s;
  const size_t off_bytes = ((size_t)n_frag + 1) * 8, need = off_bytes + total + 16;
  if (need > b->st_bases_bytes) {
    cudaFree(b->st_bases); b->st_bases = nullptr; b->st_bases_bytes = 0;
    CU(cudaMalloc(&b->st_bases, need + need / 4));
    b->st_bases_bytes = need + need / 4;
  }
  uint64_t* d_off = reinterpret_cast<uint64_t*>(b->st_bases);
  uint8_t* d_bases = b->st_bases + off_bytes;
  CU(cudaMemcpyAsync(d_bases, bases + s, total, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(d_off, frag_off, ((size_t)n_frag + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
  int rc = build_add_impl(b, d_bases, d_off, s, frag_off, frag_taxon, n_frag);
  cudaStreamSynchronize(ctx->stream);
  return rc;
}
extern "C" int slk_build_add_dev(slk_builder* b, const uint8_t* bases, const uint64_t* frag_off, const int32_t* frag_taxon,
                                 uint32_t n_frag, uint64_t total_bases) {
  if (!b || !bases || !frag_off || !frag_taxon) return fail(SLK_E_INVALID, "bad arguments");
  (void)total_bases;
  if (n_frag == 0) return SLK_OK;
  CU(cudaSetDevice(b->ctx->device));
  std::vector<uint64_t> off((size_t)n_frag + 1);
  std::vector<int32_t> tx(n_frag);
  CU(cudaMemcpy(off.data(), frag_off, ((size_t)n_frag + 1) * 8, cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(tx.data(), frag_taxon, (size_t)n_frag * 4, cudaMemcpyDeviceToHost));
  // device offsets index `bases` directly (no shift)
  return build_add_impl(b, bases, frag_off, 0, off.data(), tx.data(), n_frag);
}

extern "C" int slk_build_finish(slk_builder* b, slk_index** out) {
  if (!b || !out) return fail(SLK_E_INVALID, "bad arguments");
  slk_ctx* ctx = b->ctx;
  CU(cudaSetDevice(ctx->device));
  slk_index* idx = new (std::nothrow) slk_index;
  if (!idx) return fail(SLK_E_NOMEM, "host allocation failed");
  idx->ctx = ctx; idx->tax = b->tax; idx->params = b->params; idx->sp = b->sp;
  idx->dt = b->dt;  // host vectors copied; device arrays created below
  idx->dt.d_parent = nullptr; idx->dt.d_depth = nullptr; idx->dt.d_raw = nullptr;
  int rc = dense_upload(idx->dt);
  if (rc != SLK_OK) { slk_index_destroy(idx); return rc; }
  uint64_t n = b->count;
  uint64_t* d_sorted = nullptr; uint64_t* d_unique = nullptr; unsigned long long* d_cur = nullptr;
  auto cleanup = [&]() { cudaFree(d_sorted); cudaFree(d_unique); cudaFree(d_cur); };
#define CUX(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cleanup(); slk_index_destroy(idx); \
    return fail(e_ == cudaErrorMemoryAllocation ? SLK_E_NOMEM : SLK_E_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); } } while (0)
  unsigned long long n_unique = 0;
  if (n > 0) {
    // K3a: radix sort of the cells
    CUX(cudaMalloc(&d_sorted, n * 8));
    uint64_t* sorted_ptr = nullptr;
    // Ordered by the 32-bit line mix of the minimizer (four passes, whatever the key width): the cells of a key meet,
    // so the reduce below folds (nearly) all duplicates, and the unique cells come out in the order of the table's
    // lines, so the insert walks the table front to back instead of hitting random lines. Two different keys with the
    // same mix may interleave the cells of one of them; what the reduce then leaves double, the insert merges by LCA.
    rc = slk_sort_cells_by_line(b->cells, d_sorted, n, ctx->stream, &sorted_ptr);
    if (rc != 0) { cleanup(); slk_index_destroy(idx); return fail(SLK_E_CUDA, "radix sort failed (%d)", rc); }
    b->launches += 3 * 4;
    // K3b: segmented LCA reduce; the other buffer receives the unique cells
    uint64_t* other = sorted_ptr == d_sorted ? b->cells : d_sorted;
    CUX(cudaMalloc(&d_cur, 8));
    CUX(cudaMemsetAsync(d_cur, 0, 8, ctx->stream));
    reduce_cells_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(sorted_ptr, n, idx->dt.view(), other, d_cur);
    CUX(cudaGetLastError());
    CUX(cudaMemcpyAsync(&n_unique, d_cur, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CUX(cudaStreamSynchronize(ctx->stream));
    b->launches++;
    // K4: hash table
    rc = table_alloc(&idx->table, n_unique);
    if (rc != SLK_OK) { cleanup(); slk_index_destroy(idx); return rc; }
    CUX(cudaMemsetAsync(d_cur, 0, 8, ctx->stream));
    rc = insert_cells(ctx, idx, other, n_unique, d_cur);
    if (rc != SLK_OK) { cleanup(); slk_index_destroy(idx); return rc; }
    CUX(cudaMemcpyAsync(&n_unique, d_cur, 8, cudaMemcpyDeviceToHost, ctx->stream));   // distinct keys actually stored
    CUX(cudaStreamSynchronize(ctx->stream));
    if (n_unique & SLK_TABLE_FULL_BIT) { cleanup(); slk_index_destroy(idx); return fail(SLK_E_NOSPACE, "the minimizer table is full: records were not stored"); }
    b->launches++;
  } else {
    rc = table_alloc(&idx->table, 0);
    if (rc != SLK_OK) { cleanup(); slk_index_destroy(idx); return rc; }
  }
#undef CUX
  idx->n_records = n_unique;
  cleanup();
  cudaFree(b->cells); b->cells = nullptr; b->cap = 0; b->count = 0;
  *out = idx;
  return SLK_OK;
}

// ---------------------------------------------------------------------------------------------- distributed build
// The build of a library that is range-partitioned over `world` GPUs (slk_shard_of): every rank scans its own genomes
// (slk_build_add*), slk_build_reduce sorts and LCA-reduces them, the reduced cells -- ordered by the table-line mix and
// therefore already grouped by owner -- travel to their owners (the caller's all-to-all), and slk_index_from_cell_runs
// inserts the `world` ordered runs an owner has received front to back. The dense taxon ids inside the cells are the
// sender's: its dense -> raw list (slk_build_dense_taxa) travels with them.
// Replaces the shuffle of groupBy(idColumns).agg(udafLca), slacken/KeyValueIndex.scala:85-93; the merge of the runs on
// the owner is TaxonLCA.merge, slacken/LowestCommonAncestor.scala:152-170.
__global__ void owner_bounds_kernel(const uint64_t* __restrict__ sorted, uint64_t n, uint32_t world, uint64_t* __restrict__ first) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r > world) return;
  uint64_t lo = 0, hi = n;                       // first index whose owner is >= r
  while (r < world && lo < hi) {
    const uint64_t mid = lo + (hi - lo) / 2;
    if (slk_shard_of(sorted[mid] >> 16, world) < r) lo = mid + 1; else hi = mid;
  }
  first[r] = r < world ? lo : n;
}
extern "C" int slk_build_reduce(slk_builder* b, uint32_t world, uint64_t* counts_out) {
  if (!b || world < 1 || world > 255 || !counts_out) return fail(SLK_E_INVALID, "bad arguments");
  if (b->red || !b->seg_count.empty()) return fail(SLK_E_INVALID, "slk_build_reduce was already called on this builder");
  slk_ctx* ctx = b->ctx;
  CU(cudaSetDevice(ctx->device));
  const uint64_t n = b->count;
  b->seg_count.assign(world, 0);
  if (n == 0) { for (uint32_t r = 0; r < world; r++) counts_out[r] = 0; return SLK_OK; }
  dense_tax dt = b->dt;                           // device copy of the dense taxonomy for the reduce
  dt.d_parent = nullptr; dt.d_depth = nullptr; dt.d_raw = nullptr;
  int rc = dense_upload(dt);
  if (rc != SLK_OK) return rc;
  uint64_t* d_tmp = nullptr; uint64_t* d_first = nullptr;   // d_first: [world + 1] range starts | cursor | [world] cursor snapshots
  auto cleanup = [&]() { cudaFree(d_tmp); cudaFree(d_first); dense_free(dt); };
#define CUX(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cleanup(); \
    return fail(e_ == cudaErrorMemoryAllocation ? SLK_E_NOMEM : SLK_E_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); } } while (0)
  CUX(cudaMalloc(&d_tmp, n * 8));
  CUX(cudaMalloc(&d_first, (2 * (size_t)world + 2) * 8));
  unsigned long long* d_cur = reinterpret_cast<unsigned long long*>(d_first + world + 1);
  unsigned long long* d_snap = d_cur + 1;
  CUX(cudaMemsetAsync(d_cur, 0, 8, ctx->stream));
  uint64_t* sorted_ptr = nullptr;
  if (slk_sort_cells_by_line(b->cells, d_tmp, n, ctx->stream, &sorted_ptr) != 0) { cleanup(); return fail(SLK_E_CUDA, "radix sort failed"); }
  uint64_t* other = sorted_ptr == d_tmp ? b->cells : d_tmp;
  std::vector<uint64_t> first((size_t)world + 1), snap(world);
  owner_bounds_kernel<<<(world + 1 + 63) / 64, 64, 0, ctx->stream>>>(sorted_ptr, n, world, d_first);
  CUX(cudaGetLastError());
  CUX(cudaMemcpyAsync(first.data(), d_first, ((size_t)world + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CUX(cudaStreamSynchronize(ctx->stream));
  // One reduce per owner's range (a range never splits a key), all appending through the same cursor: the launches run
  // one after the other, so the output is grouped by owner without gaps -- the send buffer of the all-to-all.
  for (uint32_t r = 0; r < world; r++) {
    const uint64_t s0 = first[r], sn = first[r + 1] - s0;
    if (sn) reduce_cells_kernel<<<(unsigned)((sn + 255) / 256), 256, 0, ctx->stream>>>(sorted_ptr + s0, sn, dt.view(), other, d_cur);
    snapshot_kernel<<<1, 1, 0, ctx->stream>>>(d_cur, d_snap + r);
    CUX(cudaGetLastError());
  }
  CUX(cudaMemcpyAsync(snap.data(), d_snap, (size_t)world * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CUX(cudaStreamSynchronize(ctx->stream));
#undef CUX
  b->launches += 3 * 4 + 1 + 2 * world;
  b->red = other; b->red_aux = sorted_ptr;
  d_tmp = nullptr;                                // owned by red / red_aux now (the other of the two is b->cells)
  cleanup();
  for (uint32_t r = 0; r < world; r++) counts_out[r] = b->seg_count[r] = snap[r] - (r ? snap[r - 1] : 0);
  return SLK_OK;
}

// The reduced cells (device memory owned by the builder until slk_build_destroy): owner after owner without gaps.
extern "C" int slk_build_cells_dev(slk_builder* b, const uint64_t** cells_dev, uint64_t* n_out) {
  if (!b || !cells_dev || !n_out || b->seg_count.empty()) return fail(SLK_E_INVALID, "slk_build_reduce has not been called");
  uint64_t total = 0;
  for (uint64_t c : b->seg_count) total += c;
  *cells_dev = b->red; *n_out = total;
  return SLK_OK;
}

// dense -> raw taxon ids of the cells this builder emits (entry 0 = NONE); *n_out = their number, also when cap is too small
extern "C" int slk_build_dense_taxa(slk_builder* b, int32_t* raw_out, uint32_t cap, uint32_t* n_out) {
  if (!b || !n_out) return fail(SLK_E_INVALID, "bad arguments");
  *n_out = (uint32_t)b->dt.raw.size();
  if (*n_out > cap) return raw_out ? fail(SLK_E_NOSPACE, "%u dense taxa, room for %u", *n_out, cap) : SLK_OK;
  memcpy(raw_out, b->dt.raw.data(), (size_t)*n_out * 4);
  return SLK_OK;
}

// The runs an owner receives are each ordered by table line. Inserted one after the other they would sweep the whole
// table once per run; cut at the same line boundaries instead (regions of ~64 MB of table, which the L2 holds), region g
// of every run is inserted before region g + 1 of any: one sweep of the table whatever the number of runs.
// bounds[r * (R + 1) + g] = first cell of run r whose line is >= g * lines_per_region (the runs are ordered up to the
// block shuffling of the reduce, which only blurs the cuts).
__global__ void run_region_bounds_kernel(const uint64_t* __restrict__ cells, const uint64_t* __restrict__ run_first, uint32_t n_runs,
                                         uint32_t n_regions, uint64_t lines_per_region, slk_table_view tb, uint64_t* __restrict__ bounds) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_runs * (n_regions + 1)) return;
  const uint32_t r = i / (n_regions + 1), g = i % (n_regions + 1);
  const uint64_t* run = cells + run_first[r];
  uint64_t lo = 0, hi = run_first[r + 1] - run_first[r];
  const uint64_t want = (uint64_t)g * lines_per_region;
  while (g < n_regions && lo < hi) {
    const uint64_t mid = lo + (hi - lo) / 2;
    if ((slk_bucket_of(run[mid] >> 16, tb) >> 2) < want) lo = mid + 1; else hi = mid;
  }
  bounds[i] = g < n_regions ? lo : hi;
}
// thread i inserts cell number i of the region-major order: seg_first[s] <= i < seg_first[s + 1] names its segment
__global__ void __launch_bounds__(256) insert_segments_kernel(const uint64_t* __restrict__ cells, uint64_t n, const uint64_t* __restrict__ seg_first,
                                                              const uint64_t* __restrict__ seg_src, const uint32_t* __restrict__ seg_map,
                                                              uint32_t n_seg, const uint16_t* __restrict__ maps, slk_table_view tb,
                                                              slk_tax_view tx, unsigned long long* n_new) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool fresh = false;
  if (i < n) {
    uint32_t lo = 0, hi = n_seg;            // last segment with seg_first <= i
    while (hi - lo > 1) { const uint32_t mid = (lo + hi) / 2; if (seg_first[mid] <= i) lo = mid; else hi = mid; }
    uint64_t cell = cells[seg_src[lo] + (i - seg_first[lo])];
    cell = (cell & ~0xffffull) | maps[seg_map[lo] + (cell & 0xffffu)];
    fresh = insert_cell(cell, tb, tx, n_new);
  }
  const uint32_t cnt = (uint32_t)__syncthreads_count(fresh);
  if (threadIdx.x == 0 && cnt) atomicAdd(n_new, (unsigned long long)cnt);
}

extern "C" int slk_index_from_cell_runs(slk_ctx* ctx, slk_tax* tax, const slk_params* params, uint32_t world, uint32_t n_runs,
                                        const uint64_t* cells_dev, const uint64_t* run_cells, const int32_t* dense_raw,
                                        const uint32_t* run_dense, slk_index** out) {
  if (!ctx || !tax || !params || !out || world < 1 || !run_cells || !run_dense || (n_runs && !dense_raw))
    return fail(SLK_E_INVALID, "bad arguments");
  CU(cudaSetDevice(ctx->device));
  slk_index* idx = new (std::nothrow) slk_index;
  if (!idx) return fail(SLK_E_NOMEM, "host allocation failed");
  idx->ctx = ctx; idx->tax = tax; idx->params = *params;
  int rc = make_scan_params(params, &idx->sp);
  if (rc != SLK_OK) { delete idx; return rc; }
  dense_init(idx->dt);
  const int32_t n_tax = (int32_t)tax->parents.size();
  // this owner's dense taxonomy = the union of the senders' (each is ancestor-closed); one 16-bit map per run
  uint64_t total = 0, n_dense = 0;
  for (uint32_t r = 0; r < n_runs; r++) { total += run_cells[r]; n_dense += run_dense[r]; }
  if (total && !cells_dev) { delete idx; return fail(SLK_E_INVALID, "bad arguments"); }
  std::vector<uint16_t> maps(n_dense ? n_dense : 1, 0);
  rc = dense_add(idx->dt, tax, 1, &idx->dt.root);
  for (uint64_t i = 0, r = 0, left = n_runs ? run_dense[0] : 0; i < n_dense && rc == SLK_OK; i++, left--) {
    while (left == 0) left = run_dense[++r];
    const int32_t t = dense_raw[i];
    if (t < 0 || t >= n_tax) rc = fail(SLK_E_INVALID, "a sender's taxon is outside the taxonomy (0..%d)", n_tax - 1);
    else if (t != 0) { uint32_t d = 0; rc = dense_add(idx->dt, tax, t, &d); maps[i] = (uint16_t)d; }
  }
  if (rc == SLK_OK) rc = dense_upload(idx->dt);
  if (rc != SLK_OK) { slk_index_destroy(idx); return rc; }
  uint16_t* d_maps = nullptr; unsigned long long* d_new = nullptr;
  auto cleanup = [&]() { cudaFree(d_maps); cudaFree(d_new); };
#define CUX(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cleanup(); slk_index_destroy(idx); \
    return fail(e_ == cudaErrorMemoryAllocation ? SLK_E_NOMEM : SLK_E_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); } } while (0)
  CUX(cudaMalloc(&d_maps, maps.size() * 2)); CUX(cudaMalloc(&d_new, 8));
  CUX(cudaMemcpyAsync(d_maps, maps.data(), maps.size() * 2, cudaMemcpyHostToDevice, ctx->stream));
  CUX(cudaMemsetAsync(d_new, 0, 8, ctx->stream));
  rc = table_alloc(&idx->table, total, world);
  if (rc != SLK_OK) { cleanup(); slk_index_destroy(idx); return rc; }
  const uint64_t n_lines = idx->table.n_buckets >> 2;
  // SLK_REGION_SHIFT (tests): log2 of the region size in bytes, so that small tables take the interleaved path as well
  const char* rs_env = getenv("SLK_REGION_SHIFT");
  const int region_shift = rs_env ? std::min(30, std::max(8, atoi(rs_env))) : 26;
  const uint32_t n_regions = (uint32_t)std::min<uint64_t>(1024, std::max<uint64_t>(1, (n_lines * 128) >> region_shift));
  if (n_runs <= 1 || n_regions <= 1 || total == 0) {   // nothing to interleave
    uint64_t co = 0, mo = 0;
    for (uint32_t r = 0; r < n_runs; r++) {
      const uint64_t c = run_cells[r];
      if (c) {
        rc = insert_cells(ctx, idx, cells_dev + co, c, d_new, d_maps + mo);
        if (rc != SLK_OK) { cleanup(); slk_index_destroy(idx); return rc; }
      }
      co += c; mo += run_dense[r];
    }
  } else {
    const uint64_t lines_per_region = (n_lines + n_regions - 1) / n_regions;
    const uint32_t nb = n_runs * (n_regions + 1), n_seg = n_runs * n_regions;
    std::vector<uint64_t> run_first(n_runs + 1, 0), bounds(nb), seg_first((size_t)n_seg + 1), seg_src(n_seg);
    std::vector<uint32_t> seg_map(n_seg), map_off(n_runs, 0);
    for (uint32_t r = 0; r < n_runs; r++) { run_first[r + 1] = run_first[r] + run_cells[r]; if (r) map_off[r] = map_off[r - 1] + run_dense[r - 1]; }
    uint64_t* d_seg = nullptr;   // run_first | bounds, then seg_first | seg_src | seg_map
    const size_t words = std::max<size_t>((size_t)n_runs + 1 + nb, 2 * (size_t)n_seg + 1 + (n_seg + 1) / 2 + 1);
    auto cleanup2 = [&]() { cudaFree(d_seg); cleanup(); };
#define CUY(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cleanup2(); slk_index_destroy(idx); \
    return fail(e_ == cudaErrorMemoryAllocation ? SLK_E_NOMEM : SLK_E_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); } } while (0)
    CUY(cudaMalloc(&d_seg, words * 8));
    CUY(cudaMemcpyAsync(d_seg, run_first.data(), ((size_t)n_runs + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    run_region_bounds_kernel<<<(nb + 127) / 128, 128, 0, ctx->stream>>>(cells_dev, d_seg, n_runs, n_regions, lines_per_region, idx->table,
                                                                          d_seg + n_runs + 1);
    CUY(cudaGetLastError());
    CUY(cudaMemcpyAsync(bounds.data(), d_seg + n_runs + 1, (size_t)nb * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CUY(cudaStreamSynchronize(ctx->stream));
    // the runs are ordered only up to the block shuffling of the reduce, so neighbouring searches may cross: make the
    // cuts of a run monotone (they then partition the run exactly, whatever the searches found)
    for (uint32_t r = 0; r < n_runs; r++) {
      uint64_t* b = &bounds[(size_t)r * (n_regions + 1)];
      b[0] = 0; b[n_regions] = run_cells[r];
      for (uint32_t g = 1; g < n_regions; g++) b[g] = std::min<uint64_t>(std::max(b[g], b[g - 1]), run_cells[r]);
    }
    uint64_t at = 0;
    for (uint32_t g = 0; g < n_regions; g++)
      for (uint32_t r = 0; r < n_runs; r++) {
        const uint32_t sgi = g * n_runs + r;
        const uint64_t b0 = bounds[(size_t)r * (n_regions + 1) + g], b1 = bounds[(size_t)r * (n_regions + 1) + g + 1];
        seg_first[sgi] = at; seg_src[sgi] = run_first[r] + b0; seg_map[sgi] = map_off[r];
        at += b1 > b0 ? b1 - b0 : 0;
      }
    seg_first[n_seg] = at;
    if (at != total) { cleanup2(); slk_index_destroy(idx); return fail(SLK_E_CUDA, "internal: the region cuts lose cells (%llu of %llu)", (unsigned long long)at, (unsigned long long)total); }
    uint64_t* d_seg_first = d_seg; uint64_t* d_seg_src = d_seg + n_seg + 1;
    uint32_t* d_seg_map = reinterpret_cast<uint32_t*>(d_seg + 2 * (size_t)n_seg + 1);
    CUY(cudaMemcpyAsync(d_seg_first, seg_first.data(), ((size_t)n_seg + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    CUY(cudaMemcpyAsync(d_seg_src, seg_src.data(), (size_t)n_seg * 8, cudaMemcpyHostToDevice, ctx->stream));
    CUY(cudaMemcpyAsync(d_seg_map, seg_map.data(), (size_t)n_seg * 4, cudaMemcpyHostToDevice, ctx->stream));
    insert_segments_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(cells_dev, total, d_seg_first, d_seg_src, d_seg_map, n_seg,
                                                                                      d_maps, idx->table, idx->dt.view(), d_new);
    CUY(cudaGetLastError());
    CUY(cudaStreamSynchronize(ctx->stream));
#undef CUY
    cudaFree(d_seg);
  }
  unsigned long long nn = 0;
  CUX(cudaMemcpyAsync(&nn, d_new, 8, cudaMemcpyDeviceToHost, ctx->stream));
  CUX(cudaStreamSynchronize(ctx->stream));
#undef CUX
  cleanup();
  if (nn & SLK_TABLE_FULL_BIT) { slk_index_destroy(idx); return fail(SLK_E_NOSPACE, "the minimizer table is full: records were not stored"); }
  idx->n_records = nn;
  *out = idx;
  return SLK_OK;
}

// ---------------------------------------------------------------------------------------------- counts
extern "C" int slk_counts_create(slk_ctx* ctx, slk_tax* tax, int32_t n_samples, slk_counts** out) {
  if (!ctx || !tax || n_samples < 1 || !out) return fail(SLK_E_INVALID, "bad arguments");
  CU(cudaSetDevice(ctx->device));
  slk_counts* c = new (std::nothrow) slk_counts;
  if (!c) return fail(SLK_E_NOMEM, "host allocation failed");
  c->ctx = ctx; c->n_samples = n_samples; c->n_taxa = (int32_t)tax->parents.size(); c->d = nullptr;
  size_t bytes = (size_t)n_samples * c->n_taxa * 8;
  cudaError_t e = cudaMalloc(&c->d, bytes);
  if (e == cudaSuccess) e = cudaMemset(c->d, 0, bytes);
  if (e != cudaSuccess) { delete c; return fail(SLK_E_NOMEM, "counts allocation failed: %s", cudaGetErrorString(e)); }
  *out = c;
  return SLK_OK;
}
extern "C" void slk_counts_destroy(slk_counts* c) {
  if (!c) return;
  cudaSetDevice(c->ctx->device);
  cudaFree(c->d);
  delete c;
}
extern "C" int slk_counts_reset(slk_counts* c) {
  CU(cudaSetDevice(c->ctx->device));
  CU(cudaMemset(c->d, 0, (size_t)c->n_samples * c->n_taxa * 8));
  return SLK_OK;
}
extern "C" void* slk_counts_device_ptr(slk_counts* c) { return c ? c->d : nullptr; }
extern "C" int slk_counts_add(slk_counts* c, const int32_t* taxon, const uint8_t* flags, const int32_t* sample_id, uint32_t n) {
  if (!c || !taxon || !flags) return fail(SLK_E_INVALID, "bad arguments");
  if (n == 0) return SLK_OK;
  slk_ctx* ctx = c->ctx;
  CU(cudaSetDevice(ctx->device));
  int32_t *d_t = nullptr, *d_s = nullptr; uint8_t* d_f = nullptr; uint32_t* d_bad = nullptr;
  CU(cudaMalloc(&d_t, (size_t)n * 4)); CU(cudaMalloc(&d_f, n)); CU(cudaMalloc(&d_bad, 4));
  CU(cudaMemsetAsync(d_bad, 0, 4, ctx->stream));
  CU(cudaMemcpyAsync(d_t, taxon, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(d_f, flags, n, cudaMemcpyHostToDevice, ctx->stream));
  if (sample_id) {
    CU(cudaMalloc(&d_s, (size_t)n * 4));
    CU(cudaMemcpyAsync(d_s, sample_id, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
  }
  counts_add_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(d_t, d_f, d_s, n, c->n_samples, c->n_taxa, c->d, d_bad);
  CU(cudaGetLastError());
  uint32_t bad = 0;
  CU(cudaMemcpyAsync(&bad, d_bad, 4, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  cudaFree(d_t); cudaFree(d_f); cudaFree(d_s); cudaFree(d_bad);
  if (bad) return fail(SLK_E_INVALID, "a sample id or taxon is out of range");
  return SLK_OK;
}
extern "C" int slk_counts_fetch(slk_counts* c, int32_t sample, int64_t* per_taxon_out, int32_t n_taxa) {
  if (!c || !per_taxon_out || sample < 0 || sample >= c->n_samples) return fail(SLK_E_INVALID, "bad arguments");
  if (n_taxa < c->n_taxa) return fail(SLK_E_NOSPACE, "per_taxon_out needs %d entries", c->n_taxa);
  CU(cudaSetDevice(c->ctx->device));
  CU(cudaDeviceSynchronize());
  CU(cudaMemcpy(per_taxon_out, c->d + (size_t)sample * c->n_taxa, (size_t)c->n_taxa * 8, cudaMemcpyDeviceToHost));
  return SLK_OK;
}

// ---------------------------------------------------------------------------------------------- classifier
#define NSLOT 6
struct cls_slot {
  uint8_t *bases1 = nullptr, *bases2 = nullptr;   // ASCII bases, or the uint64 code blocks of packed input
  uint64_t *off1 = nullptr, *off2 = nullptr;
  uint32_t *mask1 = nullptr, *mask2 = nullptr, *len1 = nullptr, *len2 = nullptr;   // packed input only
  // ASCII input: stage 1 (pack_reads_kernel) writes the chunk's packed form here before the classify kernel runs
  uint64_t *pk_codes[2] = {nullptr, nullptr}, *pk_boff[2] = {nullptr, nullptr};
  uint32_t *pk_mask[2] = {nullptr, nullptr}, *pk_len[2] = {nullptr, nullptr};
  uint64_t* h_boff[2] = {nullptr, nullptr};   // pinned: the chunk's block offsets, computed from the host's offsets
  // compact boundary (slk_classify_batch_compact): 16-byte results, hits in read order, the chunk's ambiguity entries
  slk_read_result* res16 = nullptr; slk_hit* hits_ord = nullptr; uint64_t* hoff = nullptr; uint64_t* scan_scr = nullptr;
  uint64_t* amb = nullptr;
  uint64_t* h_amb = nullptr;   // pinned staging of the chunk's ambiguity entries (the caller's list need not be pinned)
  bool staged = false;
  int32_t* taxon = nullptr; uint8_t* flags = nullptr; slk_read_detail* detail = nullptr;
  slk_hit* hits = nullptr; uint64_t hits_cap = 0;
  unsigned long long* d_range = nullptr;   // [0] = cursor value before the kernel, [1] = after, [2] = hits in read order (compact)
  unsigned long long* h_range = nullptr;   // pinned copy
  cudaEvent_t h2d_done, k_done, d2h_done;
  cudaEvent_t copied_in, post_done;   // compact boundary: copies landed (the prep kernels may run), results ready (the copies back may run)
  bool busy = false;
  uint32_t r0 = 0, n = 0;
};
struct slk_classifier {
  slk_index* idx;
  slk_ctx* ctx;
  cudaStream_t s_h2d, s_k, s_d2h;
  cudaStream_t s_prep, s_post;   // compact boundary: the small kernels in front of / behind a chunk's classify kernel
  cudaStream_t s_k2;             // compact boundary: odd chunks classify here, so that a kernel's last wave overlaps its successor's first
  cls_slot slot[NSLOT];
  size_t cap_reads = 0, cap_bases = 0;
  bool cap_paired = false, cap_hits = false, cap_packed = false, cap_ascii = false, cap_compact = false;
  // device-resident ASCII input (slk_classify_batch_dev): its packed form, grow-only
  uint64_t *dv_codes[2] = {nullptr, nullptr}, *dv_boff[2] = {nullptr, nullptr};
  uint32_t *dv_mask[2] = {nullptr, nullptr}, *dv_len[2] = {nullptr, nullptr};
  uint64_t dv_blocks[2] = {0, 0}, dv_reads[2] = {0, 0};
  unsigned long long* d_cursor = nullptr;
  uint32_t* d_err = nullptr;
  uint16_t* d_r2d = nullptr;   // raw -> dense taxon, for the 4-byte hits of slk_classify_batch_compact_short (made on first use)
  unsigned long long* d_stats = nullptr;  // [0] probes, [1] merged hits, accumulated over launches
  slk_counts* counts = nullptr;
  int32_t counts_sample = 0;
  uint64_t launches = 0;
};
#ifndef SLK_CH_READS
#define SLK_CH_READS (1u << 19)
#endif
static const uint32_t CH_READS = SLK_CH_READS;
static const uint64_t CH_BASES = 96ull << 20;

extern "C" int slk_classifier_create(slk_index* idx, slk_classifier** out) {
  if (!idx || !out) return fail(SLK_E_INVALID, "bad arguments");
  slk_ctx* ctx = idx->ctx;
  CU(cudaSetDevice(ctx->device));
  slk_classifier* c = new (std::nothrow) slk_classifier;
  if (!c) return fail(SLK_E_NOMEM, "host allocation failed");
  c->idx = idx; c->ctx = ctx;
  // The small kernels around a chunk's classify kernel (block offsets, ambiguity bits, results in read order) run on two
  // streams of their own with the higher priority: their few blocks are placed ahead of the thousands of blocks the
  // neighbouring chunk's classify kernel still has to place. (The copy streams keep the default priority: pending work of
  // a higher-priority stream, copies included, holds back kernel launches of the lower ones.)
  int prio_lo = 0, prio_hi = 0;
  CU(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
  CU(cudaStreamCreateWithFlags(&c->s_h2d, cudaStreamNonBlocking));
  CU(cudaStreamCreateWithFlags(&c->s_k, cudaStreamNonBlocking));
  CU(cudaStreamCreateWithFlags(&c->s_d2h, cudaStreamNonBlocking));
  CU(cudaStreamCreateWithPriority(&c->s_prep, cudaStreamNonBlocking, prio_hi));
  CU(cudaStreamCreateWithPriority(&c->s_post, cudaStreamNonBlocking, prio_hi));
  CU(cudaStreamCreateWithFlags(&c->s_k2, cudaStreamNonBlocking));
  for (int i = 0; i < NSLOT; i++) {
    CU(cudaEventCreateWithFlags(&c->slot[i].h2d_done, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c->slot[i].k_done, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c->slot[i].d2h_done, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c->slot[i].copied_in, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c->slot[i].post_done, cudaEventDisableTiming));
    CU(cudaMalloc(&c->slot[i].d_range, 32));
    CU(cudaHostAlloc(&c->slot[i].h_range, 32, cudaHostAllocDefault));
  }
  CU(cudaMalloc(&c->d_cursor, 8));
  CU(cudaMalloc(&c->d_err, 4));
  CU(cudaMemset(c->d_err, 0, 4));
  CU(cudaMalloc(&c->d_stats, 16));
  CU(cudaMemset(c->d_stats, 0, 16));
  *out = c;
  return SLK_OK;
}
static void slot_free(cls_slot& s) {
  cudaFree(s.bases1); cudaFree(s.bases2); cudaFree(s.off1); cudaFree(s.off2); cudaFree(s.taxon); cudaFree(s.flags);
  cudaFree(s.detail); cudaFree(s.hits); cudaFree(s.mask1); cudaFree(s.mask2); cudaFree(s.len1); cudaFree(s.len2);
  cudaFree(s.res16); cudaFree(s.hits_ord); cudaFree(s.hoff); cudaFree(s.scan_scr); cudaFree(s.amb); cudaFreeHost(s.h_amb);
  s.h_amb = nullptr;
  s.res16 = nullptr; s.hits_ord = nullptr; s.hoff = nullptr; s.scan_scr = nullptr; s.amb = nullptr;
  for (int m = 0; m < 2; m++) {
    cudaFree(s.pk_codes[m]); cudaFree(s.pk_boff[m]); cudaFree(s.pk_mask[m]); cudaFree(s.pk_len[m]); cudaFreeHost(s.h_boff[m]);
    s.pk_codes[m] = s.pk_boff[m] = nullptr; s.pk_mask[m] = s.pk_len[m] = nullptr; s.h_boff[m] = nullptr;
  }
  s.bases1 = s.bases2 = nullptr; s.off1 = s.off2 = nullptr; s.taxon = nullptr; s.flags = nullptr; s.detail = nullptr; s.hits = nullptr;
  s.mask1 = s.mask2 = s.len1 = s.len2 = nullptr;
}
extern "C" void slk_classifier_destroy(slk_classifier* c) {
  if (!c) return;
  cudaSetDevice(c->ctx->device);
  cudaDeviceSynchronize();
  for (int i = 0; i < NSLOT; i++) {
    slot_free(c->slot[i]);
    cudaFree(c->slot[i].d_range); cudaFreeHost(c->slot[i].h_range);
    cudaEventDestroy(c->slot[i].h2d_done); cudaEventDestroy(c->slot[i].k_done); cudaEventDestroy(c->slot[i].d2h_done);
    cudaEventDestroy(c->slot[i].copied_in); cudaEventDestroy(c->slot[i].post_done);
  }
  cudaFree(c->d_cursor); cudaFree(c->d_err); cudaFree(c->d_stats); cudaFree(c->d_r2d);
  for (int m = 0; m < 2; m++) { cudaFree(c->dv_codes[m]); cudaFree(c->dv_boff[m]); cudaFree(c->dv_mask[m]); cudaFree(c->dv_len[m]); }
  cudaStreamDestroy(c->s_h2d); cudaStreamDestroy(c->s_k); cudaStreamDestroy(c->s_d2h);
  cudaStreamDestroy(c->s_prep); cudaStreamDestroy(c->s_post); cudaStreamDestroy(c->s_k2);
  delete c;
}
extern "C" int slk_classifier_sync(slk_classifier* c) {
  CU(cudaSetDevice(c->ctx->device));
  CU(cudaStreamSynchronize(c->s_k));
  return SLK_OK;
}
extern "C" void* slk_classifier_stream(slk_classifier* c) { return c ? (void*)c->s_k : nullptr; }
extern "C" uint64_t slk_classifier_launches(const slk_classifier* c) { return c ? c->launches : 0; }
extern "C" int slk_classifier_stats(slk_classifier* c, uint64_t* probes, uint64_t* merged_hits) {
  if (!c) return fail(SLK_E_INVALID, "bad arguments");
  CU(cudaSetDevice(c->ctx->device));
  CU(cudaStreamSynchronize(c->s_k));
  unsigned long long v[2];
  CU(cudaMemcpy(v, c->d_stats, 16, cudaMemcpyDeviceToHost));
  if (probes) *probes = v[0];
  if (merged_hits) *merged_hits = v[1];
  return SLK_OK;
}
// CUDA-event timing on the classifier's launch stream (bench / profiling helpers)
struct slk_event { cudaEvent_t ev; int device; };
extern "C" int slk_event_create(slk_ctx* ctx, slk_event** out) {
  if (!ctx || !out) return fail(SLK_E_INVALID, "bad arguments");
  CU(cudaSetDevice(ctx->device));
  slk_event* e = new (std::nothrow) slk_event;
  if (!e) return fail(SLK_E_NOMEM, "host allocation failed");
  e->device = ctx->device;
  CU(cudaEventCreate(&e->ev));
  *out = e;
  return SLK_OK;
}
extern "C" void slk_event_destroy(slk_event* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  cudaEventDestroy(e->ev);
  delete e;
}
extern "C" int slk_event_record(slk_event* e, slk_classifier* c) {
  if (!e || !c) return fail(SLK_E_INVALID, "bad arguments");
  CU(cudaSetDevice(e->device));
  CU(cudaEventRecord(e->ev, c->s_k));
  return SLK_OK;
}
extern "C" int slk_event_record_ctx(slk_event* e, slk_ctx* ctx) {
  if (!e || !ctx) return fail(SLK_E_INVALID, "bad arguments");
  CU(cudaSetDevice(e->device));
  CU(cudaEventRecord(e->ev, ctx->stream));
  return SLK_OK;
}
extern "C" int slk_event_elapsed_ms(slk_event* start, slk_event* end, float* ms) {
  if (!start || !end || !ms) return fail(SLK_E_INVALID, "bad arguments");
  CU(cudaSetDevice(start->device));
  CU(cudaEventSynchronize(end->ev));
  CU(cudaEventElapsedTime(ms, start->ev, end->ev));
  return SLK_OK;
}
// test hook: sort a host array of 64-bit keys on bits [begin_bit, end_bit) with the library's radix sort (stable)
extern "C" int slk_debug_sort_u64(slk_ctx* ctx, uint64_t* keys, uint64_t n, int begin_bit, int end_bit) {
  if (!ctx || (!keys && n)) return fail(SLK_E_INVALID, "bad arguments");
  CU(cudaSetDevice(ctx->device));
  if (n == 0) return SLK_OK;
  uint64_t *a = nullptr, *b = nullptr, *res = nullptr;
  CU(cudaMalloc(&a, n * 8));
  CU(cudaMalloc(&b, n * 8));
  CU(cudaMemcpy(a, keys, n * 8, cudaMemcpyHostToDevice));
  int rc = slk_sort_u64(a, b, n, begin_bit, end_bit, ctx->stream, &res);
  if (rc == 0) rc = (int)cudaMemcpy(keys, res, n * 8, cudaMemcpyDeviceToHost);
  cudaFree(a); cudaFree(b);
  if (rc != 0) return fail(SLK_E_CUDA, "radix sort failed (%d)", rc);
  return SLK_OK;
}
// test hook: the scan's 64-bit minimum (one FP64-pipe compare on the device, slk_min62) on host arrays of values < 2^62
__global__ void __launch_bounds__(256) debug_min62_kernel(const uint64_t* __restrict__ a, const uint64_t* __restrict__ b, uint64_t n,
                                                          uint64_t* __restrict__ out) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = slk_min62(a[i], b[i]);
}
extern "C" int slk_debug_min62(slk_ctx* ctx, const uint64_t* a, const uint64_t* b, uint64_t n, uint64_t* out) {
  if (!ctx || (n && (!a || !b || !out))) return fail(SLK_E_INVALID, "bad arguments");
  CU(cudaSetDevice(ctx->device));
  if (n == 0) return SLK_OK;
  uint64_t *da = nullptr, *db = nullptr, *dout = nullptr;
  CU(cudaMalloc(&da, n * 8)); CU(cudaMalloc(&db, n * 8)); CU(cudaMalloc(&dout, n * 8));
  cudaError_t e = cudaMemcpy(da, a, n * 8, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(db, b, n * 8, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    debug_min62_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(da, db, n, dout);
    e = cudaStreamSynchronize(ctx->stream);
  }
  if (e == cudaSuccess) e = cudaMemcpy(out, dout, n * 8, cudaMemcpyDeviceToHost);
  cudaFree(da); cudaFree(db); cudaFree(dout);
  if (e != cudaSuccess) return fail(SLK_E_CUDA, "slk_debug_min62 failed: %s", cudaGetErrorString(e));
  return SLK_OK;
}
extern "C" int slk_memcpy_d2d(slk_ctx* c, void* dst, const void* src, size_t bytes) {
  CU(cudaSetDevice(c->device));
  CU(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToDevice));
  return SLK_OK;
}
extern "C" int slk_classifier_attach_counts(slk_classifier* c, slk_counts* cn, int32_t sample) {
  if (!c) return fail(SLK_E_INVALID, "bad arguments");
  if (cn && (sample < 0 || sample >= cn->n_samples)) return fail(SLK_E_INVALID, "sample %d out of range", sample);
  c->counts = cn; c->counts_sample = sample;
  return SLK_OK;
}
extern "C" uint64_t slk_classify_hits_bound(const slk_params* p, uint32_t n_reads, uint64_t total_bases, int paired) {
  (void)p;
  return total_bases + (uint64_t)n_reads * (paired ? 5ull : 3ull);
}

// one mate of a batch as the kernel sees it (device pointers)
struct mate_dev {
  const uint8_t* bases = nullptr;    // ASCII bases or code blocks
  const uint64_t* off = nullptr;     // byte offsets (ASCII) or block offsets (packed)
  uint64_t shift = 0;                // value of off[] that corresponds to the start of `bases`
  const uint32_t* mask = nullptr;    // packed only
  const uint32_t* len = nullptr;     // packed only
};
// SLK_KERNEL=1 selects the first-generation fused kernel (one fragment per thread, cp.async tiles) for A/B runs;
// the default is the warp-cooperative kernel of slk_group.h (packed input only: ASCII input is packed first, stage 1)
static int kernel_generation() {
  static const int g = [] { const char* e = getenv("SLK_KERNEL"); return (e && atoi(e) == 1) ? 1 : 2; }();
  return g;
}
// ClassifyParams of one call: the thresholds (row t of taxon / flags starts `stride` elements after row t - 1)
struct cls_opts {
  double conf[SLK_MAX_THRESHOLDS];
  uint32_t n = 1;
  int32_t mhg = 2;
  uint64_t stride = 0;
};
static int make_opts(const slk_classify_opts* o, uint64_t stride, cls_opts* out) {
  if (!o) return fail(SLK_E_INVALID, "bad arguments");
  out->n = 1; out->conf[0] = o->confidence; out->mhg = o->min_hit_groups; out->stride = stride;
  return SLK_OK;
}
static int make_opts(const slk_classify_multi_opts* o, uint64_t stride, cls_opts* out) {
  if (!o || o->n_thresholds < 1 || o->n_thresholds > SLK_MAX_THRESHOLDS)
    return fail(SLK_E_INVALID, "1 to %d confidence thresholds per call", SLK_MAX_THRESHOLDS);
  if (o->n_thresholds > 1 && kernel_generation() != 2) return fail(SLK_E_UNSUPPORTED, "several thresholds need the second-generation kernel");
  out->n = o->n_thresholds; out->mhg = o->min_hit_groups; out->stride = stride;
  for (uint32_t t = 0; t < out->n; t++) out->conf[t] = o->confidence[t];
  return SLK_OK;
}
static void launch_classify(slk_classifier* c, bool hits, bool packed, const cls_opts& o, const mate_dev& m1,
                            const mate_dev& m2, uint32_t n, int32_t* taxon, uint8_t* flags, slk_read_detail* detail,
                            slk_hit* hbase, const unsigned long long* hshift, uint64_t hcap, unsigned long long* cursor,
                            unsigned long long* hits_over = nullptr, cudaStream_t st = nullptr) {
  slk_index* idx = c->idx;
  if (st == nullptr) st = c->s_k;
  if (packed && kernel_generation() == 2) {
    slk_classify2_args a;
    a.sp = idx->sp; a.tb = idx->table; a.tx = idx->dt.view();
    a.in1 = slk_group_in{reinterpret_cast<const uint64_t*>(m1.bases), m1.mask, m1.off, m1.len, m1.shift};
    a.in2 = slk_group_in{reinterpret_cast<const uint64_t*>(m2.bases), m2.mask, m2.off, m2.len, m2.shift};
    a.paired = m2.bases != nullptr; a.n_reads = n;
    a.mt.n = o.n; a.mt.stride = o.stride; a.mt.taxon_out = taxon; a.mt.flags_out = flags;
    for (uint32_t t = 0; t < o.n; t++) a.mt.confidence[t] = o.conf[t];
    a.min_hit_groups = o.mhg; a.hits = hits;
    a.taxon_out = taxon; a.flags_out = flags; a.detail_out = detail;
    a.hits_base = hbase; a.hits_shift_ptr = hshift; a.hits_cap = hcap; a.hits_cursor = cursor; a.hits_over = hits_over;
    a.counts = c->counts ? c->counts->d + (size_t)c->counts_sample * c->counts->n_taxa : nullptr;
    a.error_flag = c->d_err; a.stats = c->d_stats;
    switch (idx->sp.w) {
      case 1: slk_launch_classify2_w1(a, st); break; case 2: slk_launch_classify2_w2(a, st); break;
      case 3: slk_launch_classify2_w3(a, st); break; case 4: slk_launch_classify2_w4(a, st); break;
      case 5: slk_launch_classify2_w5(a, st); break; case 6: slk_launch_classify2_w6(a, st); break;
      case 7: slk_launch_classify2_w7(a, st); break; default: slk_launch_classify2_w8(a, st); break;
    }
    c->launches++;
    return;
  }
  slk_classify_args a;
  a.sp = idx->sp; a.tb = idx->table; a.tx = idx->dt.view();
  a.bases1 = m1.bases; a.off1 = m1.off; a.shift1 = m1.shift; a.mask1 = m1.mask; a.len1 = m1.len;
  a.bases2 = m2.bases; a.off2 = m2.off; a.shift2 = m2.shift; a.mask2 = m2.mask; a.len2 = m2.len;
  a.packed = packed; a.n_reads = n;
  a.confidence = o.conf[0]; a.min_hit_groups = o.mhg;
  a.taxon_out = taxon; a.flags_out = flags; a.detail_out = detail;
  a.hits_base = hbase; a.hits_shift_ptr = hshift; a.hits_cap = hcap; a.hits_cursor = cursor;
  a.counts = c->counts ? c->counts->d + (size_t)c->counts_sample * c->counts->n_taxa : nullptr;
  a.error_flag = c->d_err; a.stats = c->d_stats; a.hits = hits; a.stream = st;
  DISPATCH_W(idx->sp.w, slk_launch_classify_w, a);
  c->launches++;
}

static int check_error_flag(slk_classifier* c) {
  uint32_t err = 0;
  CU(cudaMemcpy(&err, c->d_err, 4, cudaMemcpyDeviceToHost));
  if (err) {
    CU(cudaMemset(c->d_err, 0, 4));
    return fail(SLK_E_UNSUPPORTED, "a fragment hit more than %d distinct taxa (per-read histogram limit)", SLK_KMAX);
  }
  return SLK_OK;
}

// block counts of a batch of ASCII reads: boff[i] = ceil(len_i / 32), boff[n] = 0 (the exclusive scan follows)
__global__ void __launch_bounds__(256) block_counts_kernel(const uint64_t* __restrict__ off, uint32_t n, uint64_t* __restrict__ boff) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) boff[i] = (off[i + 1] - off[i] + 31) >> 5;
  else if (i == n) boff[i] = 0;
}
static void launch_pack(const uint8_t* bases, const uint64_t* off, uint32_t n, const uint64_t* boff, uint64_t* codes, uint32_t* mask,
                        uint32_t* len, cudaStream_t st) {
  const uint64_t warps = ((uint64_t)n + 31) / 32;   // one warp per 32 reads, 8 warps per block
  pack_reads_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, st>>>(bases, off, n, boff, codes, mask, len);
}
// Stage 1 for device-resident ASCII input: the batch's packed form in scratch owned by the classifier (grow-only).
static int pack_device_mate(slk_classifier* c, int m, const mate_dev& in, uint32_t n, mate_dev* out) {
  uint64_t ends[2];
  CU(cudaMemcpyAsync(&ends[0], in.off, 8, cudaMemcpyDeviceToHost, c->s_k));
  CU(cudaMemcpyAsync(&ends[1], in.off + n, 8, cudaMemcpyDeviceToHost, c->s_k));
  CU(cudaStreamSynchronize(c->s_k));
  const uint64_t blocks = (ends[1] - ends[0]) / 32 + n + 1;
  if (blocks > c->dv_blocks[m]) {
    CU(cudaStreamSynchronize(c->s_k));
    cudaFree(c->dv_codes[m]); cudaFree(c->dv_mask[m]); c->dv_codes[m] = nullptr; c->dv_mask[m] = nullptr; c->dv_blocks[m] = 0;
    CU(cudaMalloc(&c->dv_codes[m], blocks * 8));
    CU(cudaMalloc(&c->dv_mask[m], blocks * 4));
    c->dv_blocks[m] = blocks;
  }
  if (n > c->dv_reads[m]) {
    CU(cudaStreamSynchronize(c->s_k));
    cudaFree(c->dv_boff[m]); cudaFree(c->dv_len[m]); c->dv_boff[m] = nullptr; c->dv_len[m] = nullptr; c->dv_reads[m] = 0;
    CU(cudaMalloc(&c->dv_boff[m], ((size_t)n + 1) * 8));
    CU(cudaMalloc(&c->dv_len[m], (size_t)n * 4));
    c->dv_reads[m] = n;
  }
  block_counts_kernel<<<(n + 256) / 256, 256, 0, c->s_k>>>(in.off, n, c->dv_boff[m]);
  if (slk_exclusive_scan_u64(c->dv_boff[m], (uint64_t)n + 1, c->s_k) != 0) return fail(SLK_E_CUDA, "prefix sum of the block counts failed");
  launch_pack(in.bases - in.shift, in.off, n, c->dv_boff[m], c->dv_codes[m], c->dv_mask[m], c->dv_len[m], c->s_k);
  c->launches += 2;
  out->bases = reinterpret_cast<const uint8_t*>(c->dv_codes[m]); out->off = c->dv_boff[m]; out->shift = 0;
  out->mask = c->dv_mask[m]; out->len = c->dv_len[m];
  return SLK_OK;
}

static int classify_dev_common(slk_classifier* c, const slk_classify_opts* opts_in, bool packed, const mate_dev& m1_in,
                               const mate_dev& m2_in, uint32_t n_reads, int32_t* taxon_out, uint8_t* flags_out,
                               slk_read_detail* detail_out, slk_hit* hits_out, uint64_t hits_cap, uint64_t* hits_used_dev) {
  mate_dev m1 = m1_in, m2 = m2_in;
  cls_opts opts;
  if (!c || !opts_in || !m1.bases || !m1.off || !taxon_out || !flags_out) return fail(SLK_E_INVALID, "bad arguments");
  if ((m2.bases == nullptr) != (m2.off == nullptr)) return fail(SLK_E_INVALID, "mate 2 needs both its data and its offsets");
  if (packed && (!m1.mask || !m1.len || (m2.bases && (!m2.mask || !m2.len))))
    return fail(SLK_E_INVALID, "packed input needs mask and len arrays");
  bool hits = hits_out != nullptr;
  if (hits && (!detail_out || !hits_used_dev)) return fail(SLK_E_INVALID, "hits_out needs detail_out and hits_used_dev");
  TRY(make_opts(opts_in, n_reads, &opts));
  CU(cudaSetDevice(c->ctx->device));
  if (n_reads == 0) return SLK_OK;
  if (!packed && kernel_generation() == 2) {   // stage 1 as a kernel of its own, then the packed-input classify kernel
    TRY(pack_device_mate(c, 0, m1_in, n_reads, &m1));
    if (m2_in.bases) TRY(pack_device_mate(c, 1, m2_in, n_reads, &m2));
    packed = true;
  }
  if (hits) CU(cudaMemsetAsync(hits_used_dev, 0, 8, c->s_k));
  launch_classify(c, hits, packed, opts, m1, m2, n_reads, taxon_out, flags_out, detail_out, hits_out, nullptr, hits_cap,
                  reinterpret_cast<unsigned long long*>(hits_used_dev));
  CU(cudaGetLastError());
  return SLK_OK;
}
extern "C" int slk_classify_batch_dev(slk_classifier* c, const slk_classify_opts* opts, const uint8_t* bases1,
                                      const uint64_t* off1, const uint8_t* bases2, const uint64_t* off2, uint32_t n_reads,
                                      int32_t* taxon_out, uint8_t* flags_out, slk_read_detail* detail_out, slk_hit* hits_out,
                                      uint64_t hits_cap, uint64_t* hits_used_dev) {
  mate_dev m1, m2;
  m1.bases = bases1; m1.off = off1; m2.bases = bases2; m2.off = off2;
  return classify_dev_common(c, opts, false, m1, m2, n_reads, taxon_out, flags_out, detail_out, hits_out, hits_cap, hits_used_dev);
}
extern "C" int slk_classify_packed_dev(slk_classifier* c, const slk_classify_opts* opts, const uint64_t* codes1,
                                       const uint32_t* mask1, const uint64_t* boff1, const uint32_t* len1,
                                       const uint64_t* codes2, const uint32_t* mask2, const uint64_t* boff2,
                                       const uint32_t* len2, uint32_t n_reads, int32_t* taxon_out, uint8_t* flags_out,
                                       slk_read_detail* detail_out, slk_hit* hits_out, uint64_t hits_cap,
                                       uint64_t* hits_used_dev) {
  mate_dev m1, m2;
  m1.bases = reinterpret_cast<const uint8_t*>(codes1); m1.off = boff1; m1.mask = mask1; m1.len = len1;
  m2.bases = reinterpret_cast<const uint8_t*>(codes2); m2.off = boff2; m2.mask = mask2; m2.len = len2;
  return classify_dev_common(c, opts, true, m1, m2, n_reads, taxon_out, flags_out, detail_out, hits_out, hits_cap, hits_used_dev);
}
// K1 as an entry point: ASCII reads already in HBM -> packed blocks (boff_dev = exclusive prefix of ceil(len/32))
extern "C" int slk_pack_reads_dev(slk_ctx* ctx, const uint8_t* bases_dev, const uint64_t* off_dev, uint32_t n_reads,
                                  const uint64_t* boff_dev, uint64_t* codes_dev, uint32_t* mask_dev, uint32_t* len_dev) {
  if (!ctx || !bases_dev || !off_dev || !boff_dev || !codes_dev || !mask_dev || !len_dev) return fail(SLK_E_INVALID, "bad arguments");
  CU(cudaSetDevice(ctx->device));
  if (n_reads == 0) return SLK_OK;
  const uint64_t warps = ((uint64_t)n_reads + 31) / 32;   // one warp per 32 reads, 8 warps per block
  pack_reads_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, ctx->stream>>>(bases_dev, off_dev, n_reads, boff_dev, codes_dev, mask_dev, len_dev);
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(ctx->stream));
  return SLK_OK;
}

static const uint64_t CH_BLOCKS = CH_BASES / 8;   // packed input: the same buffer holds CH_BLOCKS code blocks

static const uint64_t CH_PK_BLOCKS = CH_BASES / 32 + SLK_CH_READS + 1;   // most 32-base blocks a chunk of ASCII reads can have
static int ensure_slots(slk_classifier* c, bool paired, bool hits, bool packed) {
  const bool ascii = !packed && kernel_generation() == 2;
  bool need = c->cap_reads == 0 || (paired && !c->cap_paired) || (hits && !c->cap_hits) || (packed && !c->cap_packed) ||
              (ascii && !c->cap_ascii);
  paired = paired || c->cap_paired; hits = hits || c->cap_hits; packed = packed || c->cap_packed;
  const bool want_ascii = ascii || c->cap_ascii;
  if (!need) return SLK_OK;
  CU(cudaDeviceSynchronize());
  for (int i = 0; i < NSLOT; i++) {
    cls_slot& s = c->slot[i];
    slot_free(s);
    CU(cudaMalloc(&s.bases1, CH_BASES + 64));
    CU(cudaMalloc(&s.off1, ((size_t)CH_READS + 1) * 8));
    if (paired) {
      CU(cudaMalloc(&s.bases2, CH_BASES + 64));
      CU(cudaMalloc(&s.off2, ((size_t)CH_READS + 1) * 8));
    }
    if (packed) {
      CU(cudaMalloc(&s.mask1, CH_BLOCKS * 4)); CU(cudaMalloc(&s.len1, (size_t)CH_READS * 4));
      if (paired) { CU(cudaMalloc(&s.mask2, CH_BLOCKS * 4)); CU(cudaMalloc(&s.len2, (size_t)CH_READS * 4)); }
    }
    if (want_ascii) {
      for (int m = 0; m < (paired ? 2 : 1); m++) {
        CU(cudaMalloc(&s.pk_codes[m], CH_PK_BLOCKS * 8)); CU(cudaMalloc(&s.pk_mask[m], CH_PK_BLOCKS * 4));
        CU(cudaMalloc(&s.pk_boff[m], ((size_t)CH_READS + 1) * 8)); CU(cudaMalloc(&s.pk_len[m], (size_t)CH_READS * 4));
        CU(cudaHostAlloc(&s.h_boff[m], ((size_t)CH_READS + 1) * 8, cudaHostAllocDefault));
      }
    }
    CU(cudaMalloc(&s.taxon, (size_t)CH_READS * 4 * SLK_MAX_THRESHOLDS));   // one row per confidence threshold
    CU(cudaMalloc(&s.flags, (size_t)CH_READS * SLK_MAX_THRESHOLDS));
    CU(cudaMalloc(&s.detail, (size_t)CH_READS * sizeof(slk_read_detail)));
    if (hits) {
      s.hits_cap = (paired ? 2 : 1) * CH_BASES + 5ull * CH_READS;
      CU(cudaMalloc(&s.hits, s.hits_cap * sizeof(slk_hit)));
    }
  }
  c->cap_reads = CH_READS; c->cap_bases = CH_BASES; c->cap_paired = paired; c->cap_hits = hits; c->cap_packed = packed;
  c->cap_ascii = want_ascii;
  c->cap_compact = false;   // (slot_free released the compact-boundary buffers as well)
  return SLK_OK;
}

// finish one chunk: wait for its kernel, then copy its results into the caller's arrays
static int finalize_slot(slk_classifier* c, cls_slot& s, bool hits, uint32_t n_thr, uint32_t n_reads, int32_t* taxon_out, uint8_t* flags_out,
                         slk_read_detail* detail_out, slk_hit* hits_out, uint64_t hits_cap, uint64_t* hits_total, bool* nospace) {
  CU(cudaEventSynchronize(s.k_done));
  for (uint32_t t = 0; t < n_thr; t++) {   // row t of the caller's [n_thr][n_reads] arrays
    CU(cudaMemcpyAsync(taxon_out + (size_t)t * n_reads + s.r0, s.taxon + (size_t)t * CH_READS, (size_t)s.n * 4, cudaMemcpyDeviceToHost, c->s_d2h));
    CU(cudaMemcpyAsync(flags_out + (size_t)t * n_reads + s.r0, s.flags + (size_t)t * CH_READS, s.n, cudaMemcpyDeviceToHost, c->s_d2h));
  }
  if (detail_out)
    CU(cudaMemcpyAsync(detail_out + s.r0, s.detail, (size_t)s.n * sizeof(slk_read_detail), cudaMemcpyDeviceToHost, c->s_d2h));
  if (hits) {
    uint64_t lo = s.h_range[0], hi = s.h_range[1];
    *hits_total = hi;
    if (hi > hits_cap) *nospace = true;
    uint64_t end = std::min<uint64_t>(hi, hits_cap);
    if (end > lo)
      CU(cudaMemcpyAsync(hits_out + lo, s.hits, (size_t)(end - lo) * sizeof(slk_hit), cudaMemcpyDeviceToHost, c->s_d2h));
  }
  CU(cudaEventRecord(s.d2h_done, c->s_d2h));
  return SLK_OK;
}

// one mate of a batch in HOST memory, either input form
struct mate_host {
  const uint8_t* bases = nullptr;    // ASCII
  const uint64_t* codes = nullptr;   // packed
  const uint32_t* mask = nullptr;
  const uint32_t* len = nullptr;
  const uint64_t* off = nullptr;     // byte offsets (ASCII) or block offsets (packed), n+1 entries
};

// The chunked, triple-buffered H2D -> kernel -> D2H pipeline behind both host-buffer entry points.
static int classify_host_common(slk_classifier* c, const cls_opts& opts, bool packed, const mate_host& h1,
                                const mate_host& h2, uint32_t n_reads, int32_t* taxon_out, uint8_t* flags_out,
                                slk_read_detail* detail_out, slk_hit* hits_out, uint64_t hits_cap, uint64_t* hits_used) {
  if (!c || !h1.off || !taxon_out || !flags_out) return fail(SLK_E_INVALID, "bad arguments");
  const bool paired = h2.off != nullptr, hits = hits_out != nullptr;
  if (packed ? (!h1.codes || !h1.mask || !h1.len || (paired && (!h2.codes || !h2.mask || !h2.len))) : (!h1.bases || (paired && !h2.bases)))
    return fail(SLK_E_INVALID, "missing input arrays");
  if (hits && !detail_out) return fail(SLK_E_INVALID, "hits_out needs detail_out");
  if (hits_used) *hits_used = 0;
  CU(cudaSetDevice(c->ctx->device));
  if (n_reads == 0) return SLK_OK;
  TRY(ensure_slots(c, paired, hits, packed));
  const bool gen2_ascii = !packed && kernel_generation() == 2;
  CU(cudaMemsetAsync(c->d_cursor, 0, 8, c->s_k));
  const uint64_t unit_cap = packed ? CH_BLOCKS : CH_BASES;   // offsets count blocks or bytes
  uint64_t hits_total = 0;
  bool nospace = false;
  int prev = -1, ci = 0;
  uint32_t r0 = 0;
  while (r0 < n_reads) {
    // chunk [r0, r1): bounded by reads and by the data volume of either mate
    // chunk sizes ramp up at the start and down at the end of a batch, so that the first kernel does not wait for a
    // full-size copy and the last copy back is short
    uint32_t want = CH_READS;
    if (ci < 3) want = std::max<uint32_t>(CH_READS >> (3 - ci), 32768u);
    const uint32_t left = n_reads - r0;
    if (left < 2 * (uint64_t)want) want = std::max<uint32_t>(left / 2, 32768u);
    uint32_t r1 = (uint32_t)std::min<uint64_t>(n_reads, (uint64_t)r0 + want);
    auto fits = [&](uint32_t e) {
      if (h1.off[e] - h1.off[r0] > unit_cap) return false;
      if (paired && h2.off[e] - h2.off[r0] > unit_cap) return false;
      return true;
    };
    if (!fits(r1)) {
      uint32_t lo = r0, hi = r1;  // fits(lo) holds
      while (hi - lo > 1) { uint32_t mid = lo + (hi - lo) / 2; if (fits(mid)) lo = mid; else hi = mid; }
      r1 = lo;
      if (r1 == r0) return fail(SLK_E_UNSUPPORTED, "read %u is longer than %llu bases", r0, (unsigned long long)CH_BASES);
    }
    cls_slot& s = c->slot[ci % NSLOT];
    if (s.busy) { CU(cudaEventSynchronize(s.d2h_done)); s.busy = false; }
    s.r0 = r0; s.n = r1 - r0;
    mate_dev d1, d2;
    for (int mt = 0; mt < (paired ? 2 : 1); mt++) {
      const mate_host& h = mt ? h2 : h1;
      mate_dev& d = mt ? d2 : d1;
      uint8_t* dbases = mt ? s.bases2 : s.bases1;
      uint64_t* doff = mt ? s.off2 : s.off1;
      const uint64_t u0 = h.off[r0], units = h.off[r1] - u0;
      if (packed) {
        uint32_t* dmask = mt ? s.mask2 : s.mask1;
        uint32_t* dlen = mt ? s.len2 : s.len1;
        CU(cudaMemcpyAsync(dbases, h.codes + u0, units * 8, cudaMemcpyHostToDevice, c->s_h2d));
        CU(cudaMemcpyAsync(dmask, h.mask + u0, units * 4, cudaMemcpyHostToDevice, c->s_h2d));
        CU(cudaMemcpyAsync(dlen, h.len + r0, (size_t)s.n * 4, cudaMemcpyHostToDevice, c->s_h2d));
        d.mask = dmask; d.len = dlen;
      } else {
        CU(cudaMemcpyAsync(dbases, h.bases + u0, units, cudaMemcpyHostToDevice, c->s_h2d));
        if (gen2_ascii) {   // the chunk's block offsets, from the host's offsets (the slot's pinned buffer is free: its last use
          uint64_t* hb = s.h_boff[mt];   // was copied before the slot's d2h_done, which was awaited above)
          uint64_t acc = 0;
          for (uint32_t i = 0; i < s.n; i++) { hb[i] = acc; acc += (h.off[r0 + i + 1] - h.off[r0 + i] + 31) >> 5; }
          hb[s.n] = acc;
          CU(cudaMemcpyAsync(s.pk_boff[mt], hb, ((size_t)s.n + 1) * 8, cudaMemcpyHostToDevice, c->s_h2d));
        }
      }
      CU(cudaMemcpyAsync(doff, h.off + r0, ((size_t)s.n + 1) * 8, cudaMemcpyHostToDevice, c->s_h2d));
      d.bases = dbases; d.off = doff; d.shift = u0;
    }
    CU(cudaEventRecord(s.h2d_done, c->s_h2d));
    CU(cudaStreamWaitEvent(c->s_k, s.h2d_done, 0));
    bool chunk_packed = packed;
    if (gen2_ascii) {   // stage 1 as a kernel of its own: ASCII chunk -> packed chunk
      for (int mt = 0; mt < (paired ? 2 : 1); mt++) {
        mate_dev& d = mt ? d2 : d1;
        launch_pack(d.bases - d.shift, d.off, s.n, s.pk_boff[mt], s.pk_codes[mt], s.pk_mask[mt], s.pk_len[mt], c->s_k);
        c->launches++;
        d.bases = reinterpret_cast<const uint8_t*>(s.pk_codes[mt]); d.off = s.pk_boff[mt]; d.shift = 0;
        d.mask = s.pk_mask[mt]; d.len = s.pk_len[mt];
      }
      chunk_packed = true;
    }
    if (hits) snapshot_kernel<<<1, 1, 0, c->s_k>>>(c->d_cursor, s.d_range);
    launch_classify(c, hits, chunk_packed, opts, d1, d2, s.n, s.taxon, s.flags, s.detail, s.hits, s.d_range, s.hits_cap, c->d_cursor);
    CU(cudaGetLastError());
    if (hits) {
      snapshot_kernel<<<1, 1, 0, c->s_k>>>(c->d_cursor, s.d_range + 1);
      CU(cudaMemcpyAsync(s.h_range, s.d_range, 16, cudaMemcpyDeviceToHost, c->s_k));
      c->launches += 2;
    }
    CU(cudaEventRecord(s.k_done, c->s_k));
    s.busy = true;
    if (prev >= 0)
      TRY(finalize_slot(c, c->slot[prev], hits, opts.n, n_reads, taxon_out, flags_out, detail_out, hits_out, hits_cap, &hits_total, &nospace));
    prev = ci % NSLOT;
    ci++;
    r0 = r1;
  }
  if (prev >= 0)
    TRY(finalize_slot(c, c->slot[prev], hits, opts.n, n_reads, taxon_out, flags_out, detail_out, hits_out, hits_cap, &hits_total, &nospace));
  CU(cudaStreamSynchronize(c->s_d2h));
  for (int i = 0; i < NSLOT; i++) c->slot[i].busy = false;
  if (hits_used) *hits_used = hits_total;
  TRY(check_error_flag(c));
  if (nospace) return fail(SLK_E_NOSPACE, "hits_out needs room for %llu hits", (unsigned long long)hits_total);
  return SLK_OK;
}

// ---------------------------------------------------------------------------------------------- compact boundary
// The same classification with fewer bytes across PCIe ("boundary"): in, 2-bit codes + read lengths + a sparse
// list of ambiguous positions (44 instead of 72 bytes per 150-base read: the block offsets are a prefix sum the device
// does itself, and nearly all ambiguity mask words are zero); out, one 16-byte slk_read_result per read and the merged
// hits in READ ORDER (a read's hits follow its predecessor's, so no offsets travel).
#define SLK_AMB_CAP (4u << 20)   // ambiguity entries per chunk
#define SCAN_SCR_WORDS 4096      // >= slk_scan_scratch_words(CH_READS + 1)
__global__ void __launch_bounds__(256) block_counts_len_kernel(const uint32_t* __restrict__ len, uint32_t n, uint64_t* __restrict__ boff) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) boff[i] = ((uint64_t)len[i] + 31) >> 5;
  else if (i == n) boff[i] = 0;
}
// entry = read << 32 | mate << 31 | position: sets the position's bit in the read's mask words
__global__ void __launch_bounds__(256) amb_scatter_kernel(const uint64_t* __restrict__ amb, uint32_t n_amb, uint32_t r0, uint32_t n,
                                                          const uint64_t* __restrict__ boff1, const uint32_t* __restrict__ len1, uint32_t* mask1,
                                                          const uint64_t* __restrict__ boff2, const uint32_t* __restrict__ len2, uint32_t* mask2,
                                                          uint32_t* error_flag) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_amb) return;
  const uint64_t e = amb[i];
  const uint32_t r = (uint32_t)(e >> 32) - r0, mate = (uint32_t)(e >> 31) & 1u, pos = (uint32_t)e & 0x7fffffffu;
  const uint64_t* boff = mate ? boff2 : boff1;
  const uint32_t* len = mate ? len2 : len1;
  uint32_t* mask = mate ? mask2 : mask1;
  if (r >= n || boff == nullptr || pos >= len[r]) { atomicExch(error_flag, 2u); return; }
  atomicOr(&mask[boff[r] + (pos >> 5)], 1u << (pos & 31u));
}
__global__ void __launch_bounds__(256) hit_counts_kernel(const slk_read_detail* __restrict__ detail, uint32_t n, uint64_t* __restrict__ hoff) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) hoff[i] = detail[i].hit_cnt;
  else if (i == n) hoff[i] = 0;
}
// results + hits in read order. hits: the chunk's hit block (index = detail.hit_off - *lo), ord: the ordered copy
__global__ void __launch_bounds__(256) compact_results_kernel(const int32_t* __restrict__ taxon, const uint8_t* __restrict__ flags,
                                                              const slk_read_detail* __restrict__ detail, uint32_t n,
                                                              const uint64_t* __restrict__ hoff, const slk_hit* __restrict__ hits,
                                                              const unsigned long long* __restrict__ lo, uint64_t cap, slk_hit* __restrict__ ord,
                                                              slk_read_result* __restrict__ res, const uint16_t* __restrict__ r2d,
                                                              uint32_t* __restrict__ err) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const slk_read_detail d = detail[i];
  slk_read_result r;
  r.taxon = taxon[i]; r.len1 = d.len1; r.len2 = d.len2; r.hits_flags = (d.hit_cnt << 2) | (flags[i] & 3u);
  if (r2d == nullptr) res[i] = r;
  else reinterpret_cast<uint2*>(res)[i] = make_uint2((uint32_t)r.taxon, r.hits_flags);   // slk_read_result_short
  if (hits != nullptr) {
    const uint64_t src = d.hit_off - *lo, dst = hoff[i];
    if (r2d == nullptr) {
      for (uint32_t j = 0; j < d.hit_cnt; j++)
        if (src + j < cap && dst + j < cap) ord[dst + j] = hits[src + j];
    } else {   // 4-byte hits (slk_classify_batch_compact_short): dense taxon << 16 | k-mers
      uint32_t* ord4 = reinterpret_cast<uint32_t*>(ord);
      for (uint32_t j = 0; j < d.hit_cnt; j++) {
        if (src + j >= cap || dst + j >= cap) continue;
        const slk_hit h = hits[src + j];
        uint32_t wd;
        if (h.taxon == SLK_MATE_PAIR_BORDER) wd = 0xFFFFFFFFu;
        else {
          if ((uint32_t)h.count >= 0xFFFFu) atomicOr(err, 4u);
          wd = ((h.taxon == SLK_AMBIGUOUS_SPAN ? 0xFFFFu : (uint32_t)r2d[h.taxon]) << 16) | ((uint32_t)h.count & 0xFFFFu);
        }
        ord4[dst + j] = wd;
      }
    }
  }
}

static int classify_compact_impl(slk_classifier* c, const slk_classify_multi_opts* opts_in, const uint64_t* codes1,
                                 const uint32_t* len1, const uint64_t* codes2, const uint32_t* len2, const uint64_t* ambiguous,
                                 uint64_t n_ambiguous, uint32_t n_reads, void* results_out, int32_t* taxon_more,
                                 uint8_t* flags_more, void* hits_out, bool short_hits, uint64_t hits_cap, uint64_t* hits_used) {
  if (!c || !opts_in || !codes1 || !len1 || !results_out || (codes2 != nullptr) != (len2 != nullptr) || (n_ambiguous && !ambiguous))
    return fail(SLK_E_INVALID, "bad arguments");
  const size_t hit_bytes = short_hits ? 4 : sizeof(slk_hit);
  const size_t res_bytes = short_hits ? sizeof(slk_read_result_short) : sizeof(slk_read_result);
  if (short_hits && !c->d_r2d) {
    const slk_index* idx = c->idx;
    std::vector<uint16_t> r2d(idx->tax->parents.size(), 0);
    for (auto& kv : idx->dt.to_dense) r2d[kv.first] = (uint16_t)kv.second;
    CU(cudaSetDevice(c->ctx->device));
    CU(cudaMalloc(&c->d_r2d, r2d.size() * 2));
    CU(cudaMemcpy(c->d_r2d, r2d.data(), r2d.size() * 2, cudaMemcpyHostToDevice));
  }
  if (kernel_generation() != 2) return fail(SLK_E_UNSUPPORTED, "the compact entry point needs the second-generation kernel");
  cls_opts opts;
  TRY(make_opts(opts_in, CH_READS, &opts));
  if (opts.n > 1 && (!taxon_more || !flags_more)) return fail(SLK_E_INVALID, "several thresholds need taxon_more and flags_more");
  const bool paired = codes2 != nullptr, hits = hits_out != nullptr;
  if (hits_used) *hits_used = 0;
  CU(cudaSetDevice(c->ctx->device));
  if (n_reads == 0) return SLK_OK;
  TRY(ensure_slots(c, paired, true, true));
  if (!c->cap_compact) {
    CU(cudaDeviceSynchronize());
    for (int i = 0; i < NSLOT; i++) {
      cls_slot& s = c->slot[i];
      CU(cudaMalloc(&s.res16, (size_t)CH_READS * sizeof(slk_read_result)));
      CU(cudaMalloc(&s.hits_ord, s.hits_cap * sizeof(slk_hit)));
      CU(cudaMalloc(&s.hoff, ((size_t)CH_READS + 1) * 8));
      CU(cudaMalloc(&s.scan_scr, 2 * SCAN_SCR_WORDS * 8));   // one half per stream that scans
      CU(cudaMalloc(&s.amb, (size_t)SLK_AMB_CAP * 8));
      CU(cudaHostAlloc(&s.h_amb, (size_t)SLK_AMB_CAP * 8, cudaHostAllocDefault));
    }
    c->cap_compact = true;
  }
  CU(cudaMemsetAsync(c->d_cursor, 0, 8, c->s_k));
  // SLK_TRACE=1: the device timeline of every chunk (copy in / kernel / copy back) on stderr, for tuning the pipeline
  static const bool trace = getenv("SLK_TRACE") != nullptr;
  struct tr_ev { cudaEvent_t h0, h1, k0, k1, d0, d1; uint32_t n; };
  std::vector<tr_ev> tr;
  cudaEvent_t tr0 = nullptr;
  auto tmark = [&](cudaEvent_t* e, cudaStream_t st) { if (trace) { cudaEventCreate(e); cudaEventRecord(*e, st); } };
  if (trace) { cudaEventCreate(&tr0); cudaEventRecord(tr0, c->s_h2d); }
  uint64_t hits_total = 0, blk0[2] = {0, 0}, a0 = 0;
  bool nospace = false;
  int ci = 0;
  uint32_t r0 = 0;
  mate_dev sd1[NSLOT], sd2[NSLOT];
  // Results and hits in read order are made on the COPY-BACK stream in front of the copies: that work overlaps the classify
  // kernel of the next chunk. The number of hits is known on the host without another synchronisation: the cursor's advance
  // minus what spilled fragments reserved in excess.
  auto finalize = [&](cls_slot& s, int idx) -> int {
    CU(cudaEventSynchronize(s.k_done));
    if (trace) tmark(&tr[idx].d0, c->s_post);
    const uint64_t cnt = hits ? (s.h_range[1] - s.h_range[0]) - s.h_range[2] : 0;
    if (hits) {
      hit_counts_kernel<<<(s.n + 256) / 256, 256, 0, c->s_post>>>(s.detail, s.n, s.hoff);
      if (slk_exclusive_scan_u64_async(s.hoff, (uint64_t)s.n + 1, s.scan_scr + SCAN_SCR_WORDS, c->s_post) != 0) return fail(SLK_E_CUDA, "prefix sum of the hit counts failed");
    }
    compact_results_kernel<<<(s.n + 255) / 256, 256, 0, c->s_post>>>(s.taxon, s.flags, s.detail, s.n, s.hoff, hits ? s.hits : nullptr, s.d_range,
                                                                      s.hits_cap, s.hits_ord, s.res16, short_hits ? c->d_r2d : nullptr, c->d_err);
    CU(cudaGetLastError());
    CU(cudaEventRecord(s.post_done, c->s_post));
    CU(cudaStreamWaitEvent(c->s_d2h, s.post_done, 0));
    CU(cudaMemcpyAsync(static_cast<uint8_t*>(results_out) + (size_t)s.r0 * res_bytes, s.res16, (size_t)s.n * res_bytes, cudaMemcpyDeviceToHost, c->s_d2h));
    for (uint32_t t = 1; t < opts.n; t++) {
      CU(cudaMemcpyAsync(taxon_more + (size_t)(t - 1) * n_reads + s.r0, s.taxon + (size_t)t * CH_READS, (size_t)s.n * 4, cudaMemcpyDeviceToHost, c->s_d2h));
      CU(cudaMemcpyAsync(flags_more + (size_t)(t - 1) * n_reads + s.r0, s.flags + (size_t)t * CH_READS, s.n, cudaMemcpyDeviceToHost, c->s_d2h));
    }
    if (hits) {
      uint64_t used = cnt;
      if (hits_total + used > hits_cap) { nospace = true; used = hits_cap > hits_total ? hits_cap - hits_total : 0; }
      if (used) CU(cudaMemcpyAsync(static_cast<uint8_t*>(hits_out) + hits_total * hit_bytes, s.hits_ord, (size_t)used * hit_bytes, cudaMemcpyDeviceToHost, c->s_d2h));
      hits_total += cnt;
    }
    CU(cudaEventRecord(s.d2h_done, c->s_d2h));
    if (trace) tmark(&tr[idx].d1, c->s_d2h);
    return SLK_OK;
  };
  // Three chunks are in flight: chunk i is copied in and prepared while chunk i - 1 is classified and chunk i - 2 is copied
  // back, so that a classify kernel always finds its input on the device when its predecessor ends.
  auto stage_in = [&]() -> int {
    uint32_t want = CH_READS;
    if (ci < 3) want = std::max<uint32_t>(CH_READS >> (3 - ci), 32768u);
    const uint32_t left = n_reads - r0;
    if (left < 2 * (uint64_t)want) want = std::max<uint32_t>(left / 2, 32768u);
    uint32_t r1 = (uint32_t)std::min<uint64_t>(n_reads, (uint64_t)r0 + want);
    uint64_t blocks[2] = {0, 0};
    auto count = [&](const uint32_t* len, uint32_t a, uint32_t b) { uint64_t t = 0; for (uint32_t i = a; i < b; i++) t += ((uint64_t)len[i] + 31) >> 5; return t; };
    blocks[0] = count(len1, r0, r1);
    if (paired) blocks[1] = count(len2, r0, r1);
    if (blocks[0] > CH_BLOCKS || blocks[1] > CH_BLOCKS) {   // long reads: as many as fit
      uint64_t b1 = 0, b2 = 0;
      uint32_t e = r0;
      for (; e < r1; e++) {
        const uint64_t n1 = ((uint64_t)len1[e] + 31) >> 5, n2 = paired ? ((uint64_t)len2[e] + 31) >> 5 : 0;
        if (b1 + n1 > CH_BLOCKS || b2 + n2 > CH_BLOCKS) break;
        b1 += n1; b2 += n2;
      }
      if (e == r0) return fail(SLK_E_UNSUPPORTED, "read %u is longer than %llu bases", r0, (unsigned long long)CH_BASES);
      r1 = e; blocks[0] = b1; blocks[1] = b2;
    }
    uint64_t a1 = a0;   // the chunk's ambiguity entries (the list is sorted by read)
    while (a1 < n_ambiguous && (ambiguous[a1] >> 32) < r1) a1++;
    if (a1 - a0 > SLK_AMB_CAP) return fail(SLK_E_UNSUPPORTED, "more than %u ambiguous positions in one chunk: use the mask entry point", SLK_AMB_CAP);
    if (a1 > a0 && (ambiguous[a0] >> 32) < r0) return fail(SLK_E_INVALID, "the ambiguity list must be sorted by read");
    cls_slot& s = c->slot[ci % NSLOT];
    if (s.busy) { CU(cudaEventSynchronize(s.d2h_done)); s.busy = false; }
    s.r0 = r0; s.n = r1 - r0;
    if (trace) { tr.push_back(tr_ev{}); tr[ci].n = s.n; tmark(&tr[ci].h0, c->s_h2d); }
    mate_dev d1, d2;
    for (int mt = 0; mt < (paired ? 2 : 1); mt++) {
      mate_dev& d = mt ? d2 : d1;
      uint8_t* dcodes = mt ? s.bases2 : s.bases1;
      uint32_t* dlen = mt ? s.len2 : s.len1;
      CU(cudaMemcpyAsync(dcodes, (mt ? codes2 : codes1) + blk0[mt], blocks[mt] * 8, cudaMemcpyHostToDevice, c->s_h2d));
      CU(cudaMemcpyAsync(dlen, (mt ? len2 : len1) + r0, (size_t)s.n * 4, cudaMemcpyHostToDevice, c->s_h2d));
      d.bases = dcodes; d.off = mt ? s.off2 : s.off1; d.shift = 0; d.mask = mt ? s.mask2 : s.mask1; d.len = dlen;
    }
    if (a1 > a0) {
      memcpy(s.h_amb, ambiguous + a0, (size_t)(a1 - a0) * 8);
      CU(cudaMemcpyAsync(s.amb, s.h_amb, (size_t)(a1 - a0) * 8, cudaMemcpyHostToDevice, c->s_h2d));
    }
    // Block offsets, (empty) masks and the ambiguity bits are made on the device, behind the copies on a stream of their
    // own: that work overlaps the classify kernel of the previous chunk instead of standing in front of this chunk's.
    CU(cudaEventRecord(s.copied_in, c->s_h2d));
    CU(cudaStreamWaitEvent(c->s_prep, s.copied_in, 0));
    for (int mt = 0; mt < (paired ? 2 : 1); mt++) {
      mate_dev& d = mt ? d2 : d1;
      CU(cudaMemsetAsync(const_cast<uint32_t*>(d.mask), 0, (blocks[mt] + 1) * 4, c->s_prep));
      block_counts_len_kernel<<<(s.n + 256) / 256, 256, 0, c->s_prep>>>(d.len, s.n, const_cast<uint64_t*>(d.off));
      if (slk_exclusive_scan_u64_async(const_cast<uint64_t*>(d.off), (uint64_t)s.n + 1, s.scan_scr, c->s_prep) != 0)
        return fail(SLK_E_CUDA, "prefix sum of the block counts failed");
    }
    if (a1 > a0)
      amb_scatter_kernel<<<(unsigned)((a1 - a0 + 255) / 256), 256, 0, c->s_prep>>>(s.amb, (uint32_t)(a1 - a0), r0, s.n, d1.off, d1.len,
                                                                                  const_cast<uint32_t*>(d1.mask), paired ? d2.off : nullptr,
                                                                                  paired ? d2.len : nullptr, paired ? const_cast<uint32_t*>(d2.mask) : nullptr,
                                                                                  c->d_err);
    CU(cudaMemsetAsync(s.d_range, 0, 32, c->s_prep));   // [0] first hit slot of the chunk (0), [2] slots reserved in excess, [3] the cursor
    CU(cudaEventRecord(s.h2d_done, c->s_prep));
    if (trace) tmark(&tr[ci].h1, c->s_prep);
    sd1[ci % NSLOT] = d1; sd2[ci % NSLOT] = d2;
    ci++;
    r0 = r1; a0 = a1; blk0[0] += blocks[0]; blk0[1] += blocks[1];
    return SLK_OK;
  };
  // Every chunk allocates its hits from a cursor of its own (d_range[3]) into its slot's hit block, and odd and even chunks
  // classify on two streams: the first blocks of a chunk's kernel fill the SMs that the last wave of its predecessor frees.
  auto launch = [&](int idx) -> int {
    cls_slot& s = c->slot[idx % NSLOT];
    const mate_dev &d1 = sd1[idx % NSLOT], &d2 = sd2[idx % NSLOT];
    cudaStream_t st = (idx & 1) ? c->s_k2 : c->s_k;
    CU(cudaStreamWaitEvent(st, s.h2d_done, 0));
    if (trace) tmark(&tr[idx].k0, st);
    launch_classify(c, hits, true, opts, d1, d2, s.n, s.taxon, s.flags, s.detail, s.hits, nullptr, s.hits_cap, s.d_range + 3, s.d_range + 2, st);
    snapshot_kernel<<<1, 1, 0, st>>>(s.d_range + 3, s.d_range + 1);
    CU(cudaMemcpyAsync(s.h_range, s.d_range, 24, cudaMemcpyDeviceToHost, st));
    CU(cudaGetLastError());
    c->launches += 8;
    CU(cudaEventRecord(s.k_done, st));
    if (trace) tmark(&tr[idx].k1, st);
    s.busy = true;
    return SLK_OK;
  };
  int launched = 0, finalized = 0;
  for (;;) {
    const bool more = r0 < n_reads;
    if (more) TRY(stage_in());
    if (launched < ci && (launched + 1 < ci || !more)) { TRY(launch(launched)); launched++; }
    // finalize waits on the host for its chunk's kernel (it needs the chunk's hit count): keep it two kernels behind the
    // launches, so that the wait falls on a kernel that has long finished and the next launch is never held up by it
    // (with one kernel of lag the device idled ~0.3 ms between every pair of chunks: SLK_TRACE=1)
    if (finalized < launched && (finalized + 2 < launched || (launched == ci && !more))) { TRY(finalize(c->slot[finalized % NSLOT], finalized)); finalized++; }
    if (!more && finalized == ci) break;
  }
  CU(cudaStreamSynchronize(c->s_d2h));
  if (trace) {
    for (size_t i = 0; i < tr.size(); i++) {
      float t[6];
      cudaEvent_t ev[6] = {tr[i].h0, tr[i].h1, tr[i].k0, tr[i].k1, tr[i].d0, tr[i].d1};
      for (int j = 0; j < 6; j++) { cudaEventElapsedTime(&t[j], tr0, ev[j]); cudaEventDestroy(ev[j]); }
      fprintf(stderr, "[slk trace] chunk %2zu n=%7u  in %7.3f-%7.3f  kernel %7.3f-%7.3f  out %7.3f-%7.3f ms\n", i, tr[i].n, t[0], t[1], t[2], t[3], t[4], t[5]);
    }
    cudaEventDestroy(tr0);
  }
  for (int i = 0; i < NSLOT; i++) c->slot[i].busy = false;
  if (hits_used) *hits_used = hits_total;
  {
    uint32_t err = 0;
    CU(cudaMemcpy(&err, c->d_err, 4, cudaMemcpyDeviceToHost));
    if (err & 2u) { CU(cudaMemset(c->d_err, 0, 4)); return fail(SLK_E_INVALID, "an ambiguity entry names a read or position outside its chunk"); }
    if (err & 4u) { CU(cudaMemset(c->d_err, 0, 4)); return fail(SLK_E_UNSUPPORTED, "a merged hit of 65 535 or more k-mers does not fit the 4-byte hit format: use slk_classify_batch_compact"); }
  }
  TRY(check_error_flag(c));
  if (nospace) return fail(SLK_E_NOSPACE, "hits_out needs room for %llu hits", (unsigned long long)hits_total);
  return SLK_OK;
}
extern "C" int slk_classify_batch_compact(slk_classifier* c, const slk_classify_multi_opts* opts_in, const uint64_t* codes1,
                                          const uint32_t* len1, const uint64_t* codes2, const uint32_t* len2, const uint64_t* ambiguous,
                                          uint64_t n_ambiguous, uint32_t n_reads, slk_read_result* results_out, int32_t* taxon_more,
                                          uint8_t* flags_more, slk_hit* hits_out, uint64_t hits_cap, uint64_t* hits_used) {
  return classify_compact_impl(c, opts_in, codes1, len1, codes2, len2, ambiguous, n_ambiguous, n_reads, results_out, taxon_more,
                               flags_more, hits_out, false, hits_cap, hits_used);
}
extern "C" int slk_classify_batch_compact_short(slk_classifier* c, const slk_classify_multi_opts* opts_in, const uint64_t* codes1,
                                                const uint32_t* len1, const uint64_t* codes2, const uint32_t* len2,
                                                const uint64_t* ambiguous, uint64_t n_ambiguous, uint32_t n_reads,
                                                slk_read_result_short* results_out, int32_t* taxon_more, uint8_t* flags_more,
                                                uint32_t* hits_out, uint64_t hits_cap, uint64_t* hits_used) {
  return classify_compact_impl(c, opts_in, codes1, len1, codes2, len2, ambiguous, n_ambiguous, n_reads, results_out, taxon_more,
                               flags_more, hits_out, true, hits_cap, hits_used);
}

extern "C" int slk_classify_batch(slk_classifier* c, const slk_classify_opts* opts, const uint8_t* bases1, const uint64_t* off1,
                                  const uint8_t* bases2, const uint64_t* off2, uint32_t n_reads, int32_t* taxon_out,
                                  uint8_t* flags_out, slk_read_detail* detail_out, slk_hit* hits_out, uint64_t hits_cap,
                                  uint64_t* hits_used) {
  if ((bases2 == nullptr) != (off2 == nullptr)) return fail(SLK_E_INVALID, "bases2/off2 must both be given or both be NULL");
  mate_host h1, h2;
  h1.bases = bases1; h1.off = off1; h2.bases = bases2; h2.off = off2;
  cls_opts o;
  TRY(make_opts(opts, CH_READS, &o));
  return classify_host_common(c, o, false, h1, h2, n_reads, taxon_out, flags_out, detail_out, hits_out, hits_cap, hits_used);
}
extern "C" int slk_classify_batch_packed(slk_classifier* c, const slk_classify_opts* opts, const uint64_t* codes1,
                                         const uint32_t* mask1, const uint64_t* boff1, const uint32_t* len1,
                                         const uint64_t* codes2, const uint32_t* mask2, const uint64_t* boff2,
                                         const uint32_t* len2, uint32_t n_reads, int32_t* taxon_out, uint8_t* flags_out,
                                         slk_read_detail* detail_out, slk_hit* hits_out, uint64_t hits_cap,
                                         uint64_t* hits_used) {
  mate_host h1, h2;
  h1.codes = codes1; h1.mask = mask1; h1.off = boff1; h1.len = len1;
  h2.codes = codes2; h2.mask = mask2; h2.off = boff2; h2.len = len2;
  cls_opts o;
  TRY(make_opts(opts, CH_READS, &o));
  return classify_host_common(c, o, true, h1, h2, n_reads, taxon_out, flags_out, detail_out, hits_out, hits_cap, hits_used);
}
// Classifier.classify with several thresholds (slacken/Classifier.scala:156-170): the reads are scanned and looked up ONCE,
// resolveTree runs once per threshold; taxon_out / flags_out are [n_thresholds][n_reads], details and hits are shared.
extern "C" int slk_classify_batch_packed_multi(slk_classifier* c, const slk_classify_multi_opts* opts, const uint64_t* codes1,
                                               const uint32_t* mask1, const uint64_t* boff1, const uint32_t* len1,
                                               const uint64_t* codes2, const uint32_t* mask2, const uint64_t* boff2,
                                               const uint32_t* len2, uint32_t n_reads, int32_t* taxon_out, uint8_t* flags_out,
                                               slk_read_detail* detail_out, slk_hit* hits_out, uint64_t hits_cap,
                                               uint64_t* hits_used) {
  mate_host h1, h2;
  h1.codes = codes1; h1.mask = mask1; h1.off = boff1; h1.len = len1;
  h2.codes = codes2; h2.mask = mask2; h2.off = boff2; h2.len = len2;
  cls_opts o;
  TRY(make_opts(opts, CH_READS, &o));
  return classify_host_common(c, o, true, h1, h2, n_reads, taxon_out, flags_out, detail_out, hits_out, hits_cap, hits_used);
}
