// slk_split.cu -- the split classification path for libraries sharded over several GPUs by minimizer hash range
// (SURVEY section 8e): scan -> span words | route by owner | probe on the owner | merge + resolve on the query side.
// The exchange between the steps (one all-to-all of 8-byte keys, one of 4-byte taxa) is the caller's: the Python
// host uses torch.distributed (NCCL over NVLink), a JVM host would use its own transport. Every pointer here is a
// DEVICE pointer unless it says host; every call is synchronous.
#include <cuda_runtime.h>
#include <stdint.h>

#include <string.h>

#include <algorithm>
#include <new>
#include <vector>

#include "../../include/slacken_gpu.h"
#include "slk_host.h"
#include "slk_kernels.cuh"
#include "slk_sort.h"

// ---------------------------------------------------------------------------------------------- kernels
// keys (compressed minimizers) -> raw taxon of the record, 0 when the shard has none
__global__ void __launch_bounds__(256) probe_keys_kernel(slk_table_view tb, slk_tax_view tx, const uint64_t* __restrict__ keys,
                                                         uint64_t n, int32_t* __restrict__ taxa) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t d = slk_probe(tb, keys[i]);
  taxa[i] = d ? tx.raw[d] : 0;
}

// longest mate 1 / mate 2 of a batch (out[0], out[1])
__global__ void __launch_bounds__(256) max_len_kernel(const uint64_t* __restrict__ off1, const uint64_t* __restrict__ off2, uint32_t n,
                                                      uint32_t* out) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t l1 = 0, l2 = 0;
  if (r < n) {
    const uint64_t a = off1[r + 1] - off1[r];
    l1 = a > 0xffffffffull ? 0xffffffffu : (uint32_t)a;
    if (off2) { const uint64_t b = off2[r + 1] - off2[r]; l2 = b > 0xffffffffull ? 0xffffffffu : (uint32_t)b; }
  }
  l1 = __reduce_max_sync(0xffffffffu, l1); l2 = __reduce_max_sync(0xffffffffu, l2);
  if ((threadIdx.x & 31) == 0) { if (l1) atomicMax(out, l1); if (l2) atomicMax(out + 1, l2); }
}
// rows of `stride` slots -> the packed span array, one warp per fragment
__global__ void __launch_bounds__(256) compact_spans_kernel(const uint64_t* __restrict__ scratch, uint32_t stride,
                                                            const uint64_t* __restrict__ span_off, uint32_t n_reads,
                                                            uint64_t* __restrict__ spans) {
  const uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (r >= n_reads) return;
  const uint64_t s0 = span_off[r];
  const uint32_t n = (uint32_t)(span_off[r + 1] - s0);
  const uint64_t* row = scratch + (uint64_t)r * stride;
  for (uint32_t j = lane; j < n; j += 32) spans[s0 + j] = row[j];
}

// Routing of the SEQ spans of a batch: pass 1 counts per owner, pass 2 writes (key, span index) grouped by owner.
// Counters per owner live in shared memory; a block touches the global counters once per owner.
#define ROUTE_MAX_WORLD 1024
template <bool SCATTER>
__global__ void __launch_bounds__(256) route_kernel(const uint64_t* __restrict__ spans, uint64_t n, uint32_t world,
                                                    unsigned long long* cursors, uint64_t* __restrict__ send_keys,
                                                    uint32_t* __restrict__ send_idx) {
  __shared__ uint32_t s_cnt[ROUTE_MAX_WORLD];
  __shared__ unsigned long long s_base[ROUTE_MAX_WORLD];
  for (uint32_t d = threadIdx.x; d < world; d += blockDim.x) s_cnt[d] = 0;
  __syncthreads();
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t ck = 0;
  uint32_t dest = 0, pos = 0;
  bool seq = false;
  if (i < n) {
    const uint64_t w = spans[i];
    seq = SLK_SPAN_TYPE(w) == SLK_E_SEQ;
    ck = SLK_SPAN_KEY(w);
  }
  if (seq) { dest = slk_shard_of(ck, world); pos = atomicAdd(&s_cnt[dest], 1u); }
  __syncthreads();
  for (uint32_t d = threadIdx.x; d < world; d += blockDim.x)
    if (s_cnt[d]) s_base[d] = atomicAdd(&cursors[d], (unsigned long long)s_cnt[d]);
  if (!SCATTER) return;
  __syncthreads();
  if (seq) { send_keys[s_base[dest] + pos] = ck; send_idx[s_base[dest] + pos] = (uint32_t)i; }
}

// taxa come back in send order: dense label of span send_idx[j] = raw2dense[taxa[j]]
__global__ void __launch_bounds__(256) unroute_kernel(const int32_t* __restrict__ taxa, const uint32_t* __restrict__ send_idx,
                                                      uint64_t n, const uint16_t* __restrict__ raw2dense, int32_t n_tax,
                                                      uint16_t* __restrict__ dense, uint32_t* bad) {
  const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int32_t t = taxa[j];
  uint16_t d = 0;
  if (t != 0) {
    if (t < 0 || t >= n_tax || (d = raw2dense[t]) == 0) { atomicExch(bad, 1u); d = 0; }
  }
  dense[send_idx[j]] = d;
}

__global__ void __launch_bounds__(128) resolve_spans_kernel(slk_tax_view tx, int32_t k, const uint64_t* __restrict__ spans,
                                                            const uint64_t* __restrict__ span_off,
                                                            const uint16_t* __restrict__ dense, uint32_t n_reads, int paired,
                                                            double confidence, int32_t min_hit_groups,
                                                            int32_t* __restrict__ taxon_out, uint8_t* __restrict__ flags_out,
                                                            slk_read_detail* __restrict__ detail_out,
                                                            slk_hit* __restrict__ hits_out, uint32_t* error_flag) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_reads) return;
  const uint64_t s0 = span_off[r], s1 = span_off[r + 1];
  slk_hit* ho = hits_out ? hits_out + s0 : nullptr;   // a fragment has at most as many merged hits as spans
  slk_frag_result res;
  slk_resolve_spans(tx, k, spans + s0, dense + s0, (uint32_t)(s1 - s0), confidence, min_hit_groups,
                    [&](int32_t l, int32_t c) { if (ho) { slk_hit h; h.taxon = l >= 0 ? tx.raw[l] : l; h.count = c; *ho++ = h; } }, res);
  taxon_out[r] = res.taxon;
  flags_out[r] = (uint8_t)(res.flags & 3u);
  if (res.flags & SLK_F_OVERFLOW) atomicExch(error_flag, 1u);
  if (detail_out) {
    slk_read_detail d;
    d.hit_off = s0; d.hit_cnt = hits_out ? res.n_hits : 0u;
    d.len1 = res.kmers1 + (uint32_t)(k - 1);
    d.len2 = paired ? res.kmers2 + (uint32_t)(k - 1) : 0xFFFFFFFFu;
    d.num_distinct = res.num_distinct;
    detail_out[r] = d;
  }
}

// ---------------------------------------------------------------------------------------------- host
#define SLK_SPAN_SCRATCH_MAX (16ull << 30)   // most scratch the one-pass scan may take (it falls back to two passes beyond)
struct slk_resolver {
  slk_ctx* ctx;
  slk_tax* tax;
  slk_scan_params sp;
  dense_tax dt;
  uint16_t* d_r2d = nullptr;
  uint32_t* d_err = nullptr;
  uint16_t* d_dense = nullptr;   // scratch: dense taxon of every span of the batch being resolved (grow-only)
  uint64_t dense_cap = 0;
};

extern "C" int slk_index_taxa(slk_index* idx, int32_t* out, uint32_t cap, uint32_t* n_out) {
  if (!idx || !n_out) return slk_fail(SLK_E_INVALID, "bad arguments");
  const uint32_t n = (uint32_t)idx->dt.raw.size() - 1;   // without dense 0 = NONE
  *n_out = n;
  if (!out) return SLK_OK;
  if (cap < n) return slk_fail(SLK_E_NOSPACE, "taxa buffer too small: %u < %u", cap, n);
  for (uint32_t i = 0; i < n; i++) out[i] = idx->dt.raw[i + 1];
  return SLK_OK;
}

extern "C" int slk_resolver_create(slk_ctx* ctx, slk_tax* tax, const slk_params* params, const int32_t* taxa, uint32_t n,
                                   slk_resolver** out) {
  if (!ctx || !tax || !params || !out || (n && !taxa)) return slk_fail(SLK_E_INVALID, "bad arguments");
  SLK_CU(cudaSetDevice(ctx->device));
  slk_resolver* r = new (std::nothrow) slk_resolver;
  if (!r) return slk_fail(SLK_E_NOMEM, "host allocation failed");
  r->ctx = ctx; r->tax = tax;
  int rc = slk_make_scan_params_checked(params, &r->sp);
  if (rc != SLK_OK) { delete r; return rc; }
  slk_dense_init(r->dt);
  const int32_t n_tax = (int32_t)tax->parents.size();
  // the same numbering on every rank: ROOT first, then the taxa in increasing raw id
  std::vector<int32_t> sorted(taxa, taxa + n);
  std::sort(sorted.begin(), sorted.end());
  rc = slk_dense_add(r->dt, tax, 1, &r->dt.root);
  for (size_t i = 0; i < sorted.size() && rc == SLK_OK; i++) {
    if (sorted[i] <= 0 || sorted[i] >= n_tax) rc = slk_fail(SLK_E_INVALID, "taxon %d is outside the taxonomy", sorted[i]);
    else rc = slk_dense_add(r->dt, tax, sorted[i], nullptr);
  }
  if (rc == SLK_OK) rc = slk_dense_upload(r->dt);
  if (rc != SLK_OK) { slk_dense_free(r->dt); delete r; return rc; }
  std::vector<uint16_t> r2d((size_t)n_tax, 0);
  for (auto& kv : r->dt.to_dense) r2d[kv.first] = (uint16_t)kv.second;
  cudaError_t e = cudaMalloc(&r->d_r2d, (size_t)n_tax * 2);
  if (e == cudaSuccess) e = cudaMemcpy(r->d_r2d, r2d.data(), (size_t)n_tax * 2, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMalloc(&r->d_err, 4);
  if (e == cudaSuccess) e = cudaMemset(r->d_err, 0, 4);
  if (e != cudaSuccess) {
    cudaFree(r->d_r2d); cudaFree(r->d_err); slk_dense_free(r->dt); delete r;
    return slk_fail(SLK_E_CUDA, "resolver allocation failed: %s", cudaGetErrorString(e));
  }
  *out = r;
  return SLK_OK;
}
extern "C" void slk_resolver_destroy(slk_resolver* r) {
  if (!r) return;
  cudaFree(r->d_r2d); cudaFree(r->d_err); cudaFree(r->d_dense);
  slk_dense_free(r->dt);
  delete r;
}

#define SLK_DISPATCH_SPANS(w, a)                         \
  switch (w) {                                           \
    case 1: slk_launch_spans_w1(a); break; case 2: slk_launch_spans_w2(a); break; \
    case 3: slk_launch_spans_w3(a); break; case 4: slk_launch_spans_w4(a); break; \
    case 5: slk_launch_spans_w5(a); break; case 6: slk_launch_spans_w6(a); break; \
    case 7: slk_launch_spans_w7(a); break; default: slk_launch_spans_w8(a); break; \
  }

extern "C" int slk_scan_spans_dev(slk_ctx* ctx, const slk_params* params, const uint8_t* bases1, const uint64_t* off1,
                                  const uint8_t* bases2, const uint64_t* off2, uint32_t n_reads, uint64_t* span_off,
                                  uint64_t* spans, uint64_t cap, uint64_t* n_spans_host) {
  if (!ctx || !params || !bases1 || !off1 || !span_off || !n_spans_host || ((bases2 == nullptr) != (off2 == nullptr)))
    return slk_fail(SLK_E_INVALID, "bad arguments");
  SLK_CU(cudaSetDevice(ctx->device));
  slk_spans_args a;
  int rc = slk_make_scan_params_checked(params, &a.sp);
  if (rc != SLK_OK) return rc;
  *n_spans_host = 0;
  SLK_CU(cudaMemsetAsync(span_off, 0, ((size_t)n_reads + 1) * 8, ctx->scan_stream));
  if (n_reads == 0) { SLK_CU(cudaStreamSynchronize(ctx->scan_stream)); return SLK_OK; }
  a.bases1 = bases1; a.off1 = off1; a.bases2 = bases2; a.off2 = off2; a.n_reads = n_reads;
  a.span_off = span_off; a.spans = nullptr; a.stream = ctx->scan_stream;
  ctx->pending_spans.valid = false;
  // A count-only call that slk_emit_spans_dev will follow: scan ONCE, into rows of the context's scratch, when the rows
  // of this batch (one slot per k-mer window of its longest reads) fit a bounded scratch; otherwise count now, scan again later.
  bool strided = false;
  if (!spans && !getenv("SLK_SPANS_TWO_PASS")) {
    if (!ctx->d_maxlen) SLK_CU(cudaMalloc(&ctx->d_maxlen, 12));
    SLK_CU(cudaMemsetAsync(ctx->d_maxlen, 0, 12, ctx->scan_stream));
    max_len_kernel<<<(n_reads + 255) / 256, 256, 0, ctx->scan_stream>>>(off1, off2, n_reads, ctx->d_maxlen);
    uint32_t ml[2] = {0, 0};
    SLK_CU(cudaMemcpyAsync(ml, ctx->d_maxlen, 8, cudaMemcpyDeviceToHost, ctx->scan_stream));
    SLK_CU(cudaStreamSynchronize(ctx->scan_stream));
    const uint64_t k = (uint64_t)a.sp.k;
    const uint64_t stride = (ml[0] >= k ? ml[0] - k + 1 : 0) + (bases2 ? 1 + (ml[1] >= k ? ml[1] - k + 1 : 0) : 0) + 1;
    const uint64_t words = stride * (uint64_t)n_reads;
    if (stride <= 4096 && words * 8 <= SLK_SPAN_SCRATCH_MAX) {
      if (ctx->span_scratch_words < words) {
        cudaFree(ctx->span_scratch); ctx->span_scratch = nullptr; ctx->span_scratch_words = 0;
        size_t free_b = 0, total_b = 0;
        SLK_CU(cudaMemGetInfo(&free_b, &total_b));
        const uint64_t want = words + words / 8;
        if (want * 8 <= free_b / 4 && cudaMalloc(&ctx->span_scratch, want * 8) == cudaSuccess) ctx->span_scratch_words = want;
        else cudaGetLastError();
      }
      if (ctx->span_scratch_words >= words) {
        a.scratch = ctx->span_scratch; a.stride = (uint32_t)stride; a.overflow = ctx->d_maxlen + 2;
        strided = true;
      }
    }
  }
  SLK_DISPATCH_SPANS(a.sp.w, a);
  SLK_CU(cudaGetLastError());
  int e = slk_exclusive_scan_u64(span_off, (uint64_t)n_reads + 1, ctx->scan_stream);
  if (e != 0) return slk_fail(SLK_E_CUDA, "prefix sum of the span counts failed (%d)", e);
  uint64_t total = 0;
  SLK_CU(cudaMemcpyAsync(&total, span_off + n_reads, 8, cudaMemcpyDeviceToHost, ctx->scan_stream));
  SLK_CU(cudaStreamSynchronize(ctx->scan_stream));
  *n_spans_host = total;
  if (strided) {
    uint32_t over = 0;
    SLK_CU(cudaMemcpyAsync(&over, ctx->d_maxlen + 2, 4, cudaMemcpyDeviceToHost, ctx->scan_stream));
    SLK_CU(cudaStreamSynchronize(ctx->scan_stream));
    if (!over) {   // (never expected: the stride is an upper bound; if it were exceeded the emit call simply scans again)
      ctx->pending_spans.bases1 = bases1; ctx->pending_spans.off1 = off1; ctx->pending_spans.bases2 = bases2; ctx->pending_spans.off2 = off2;
      ctx->pending_spans.span_off = span_off; ctx->pending_spans.n_reads = n_reads; ctx->pending_spans.stride = a.stride;
      ctx->pending_spans.valid = true;
    }
  }
  if (!spans) return SLK_OK;   // count only: the caller sizes its buffer and calls slk_emit_spans_dev
  if (total > cap) return slk_fail(SLK_E_NOSPACE, "span buffer too small: %llu spans, room for %llu", (unsigned long long)total, (unsigned long long)cap);
  a.spans = spans;
  SLK_DISPATCH_SPANS(a.sp.w, a);
  SLK_CU(cudaGetLastError());
  SLK_CU(cudaStreamSynchronize(ctx->scan_stream));
  return SLK_OK;
}
// second half of slk_scan_spans_dev for a caller that asked for the count first: span_off is what that call left
extern "C" int slk_emit_spans_dev(slk_ctx* ctx, const slk_params* params, const uint8_t* bases1, const uint64_t* off1,
                                  const uint8_t* bases2, const uint64_t* off2, uint32_t n_reads, const uint64_t* span_off,
                                  uint64_t* spans) {
  if (!ctx || !params || !bases1 || !off1 || !span_off || !spans || ((bases2 == nullptr) != (off2 == nullptr)))
    return slk_fail(SLK_E_INVALID, "bad arguments");
  SLK_CU(cudaSetDevice(ctx->device));
  if (n_reads == 0) return SLK_OK;
  auto& pd = ctx->pending_spans;
  if (pd.valid && pd.bases1 == bases1 && pd.off1 == off1 && pd.bases2 == bases2 && pd.off2 == off2 && pd.span_off == span_off &&
      pd.n_reads == n_reads) {   // the count-only call already scanned: only move the rows together
    pd.valid = false;
    compact_spans_kernel<<<(unsigned)(((uint64_t)n_reads * 32 + 255) / 256), 256, 0, ctx->scan_stream>>>(ctx->span_scratch, pd.stride, span_off,
                                                                                                    n_reads, spans);
    SLK_CU(cudaGetLastError());
    SLK_CU(cudaStreamSynchronize(ctx->scan_stream));
    return SLK_OK;
  }
  pd.valid = false;
  slk_spans_args a;
  int rc = slk_make_scan_params_checked(params, &a.sp);
  if (rc != SLK_OK) return rc;
  a.bases1 = bases1; a.off1 = off1; a.bases2 = bases2; a.off2 = off2; a.n_reads = n_reads;
  a.span_off = const_cast<uint64_t*>(span_off); a.spans = spans; a.stream = ctx->scan_stream;
  SLK_DISPATCH_SPANS(a.sp.w, a);
  SLK_CU(cudaGetLastError());
  SLK_CU(cudaStreamSynchronize(ctx->scan_stream));
  return SLK_OK;
}

extern "C" int slk_route_spans_dev(slk_ctx* ctx, const uint64_t* spans, uint64_t n_spans, uint32_t world, uint64_t* send_keys,
                                   uint32_t* send_idx, uint64_t cap, uint64_t* counts_host) {
  if (!ctx || !counts_host || world == 0 || world > 1024 || (n_spans && !spans)) return slk_fail(SLK_E_INVALID, "bad arguments");
  if (n_spans > 0xffffffffull) return slk_fail(SLK_E_UNSUPPORTED, "more than 2^32 spans in one batch");
  SLK_CU(cudaSetDevice(ctx->device));
  for (uint32_t i = 0; i < world; i++) counts_host[i] = 0;
  if (n_spans == 0) return SLK_OK;
  unsigned long long* d_cur = nullptr;
  SLK_CU(cudaMalloc(&d_cur, (size_t)world * 8));
  auto done = [&](int rc) { cudaFree(d_cur); return rc; };
  const unsigned grid = (unsigned)((n_spans + 255) / 256);
  if (cudaMemsetAsync(d_cur, 0, (size_t)world * 8, ctx->stream) != cudaSuccess) return done(slk_fail(SLK_E_CUDA, "memset failed"));
  route_kernel<false><<<grid, 256, 0, ctx->stream>>>(spans, n_spans, world, d_cur, nullptr, nullptr);
  std::vector<unsigned long long> cnt(world);
  if (cudaMemcpyAsync(cnt.data(), d_cur, (size_t)world * 8, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
      cudaStreamSynchronize(ctx->stream) != cudaSuccess)
    return done(slk_fail(SLK_E_CUDA, "routing count failed: %s", cudaGetErrorString(cudaGetLastError())));
  uint64_t total = 0;
  std::vector<unsigned long long> start(world);
  for (uint32_t i = 0; i < world; i++) { counts_host[i] = cnt[i]; start[i] = total; total += cnt[i]; }
  if (!send_keys || !send_idx) return done(SLK_OK);   // size query
  if (total > cap) return done(slk_fail(SLK_E_NOSPACE, "send buffers too small: %llu keys, room for %llu", (unsigned long long)total, (unsigned long long)cap));
  if (cudaMemcpyAsync(d_cur, start.data(), (size_t)world * 8, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess)
    return done(slk_fail(SLK_E_CUDA, "copy failed"));
  route_kernel<true><<<grid, 256, 0, ctx->stream>>>(spans, n_spans, world, d_cur, send_keys, send_idx);
  if (cudaStreamSynchronize(ctx->stream) != cudaSuccess)
    return done(slk_fail(SLK_E_CUDA, "routing failed: %s", cudaGetErrorString(cudaGetLastError())));
  return done(SLK_OK);
}

extern "C" int slk_probe_keys_dev(slk_index* idx, const uint64_t* keys, uint64_t n, int32_t* taxa) {
  if (!idx || (n && (!keys || !taxa))) return slk_fail(SLK_E_INVALID, "bad arguments");
  SLK_CU(cudaSetDevice(idx->ctx->device));
  if (n == 0) return SLK_OK;
  probe_keys_kernel<<<(unsigned)((n + 255) / 256), 256, 0, idx->ctx->stream>>>(idx->table, idx->dt.view(), keys, n, taxa);
  SLK_CU(cudaGetLastError());
  SLK_CU(cudaStreamSynchronize(idx->ctx->stream));
  return SLK_OK;
}

extern "C" int slk_resolve_spans_dev(slk_resolver* r, const slk_classify_opts* opts, const uint64_t* spans, const uint64_t* span_off,
                                     uint64_t n_spans, uint32_t n_reads, int paired, const uint32_t* send_idx, const int32_t* taxa,
                                     uint64_t n_routed, int32_t* taxon_out, uint8_t* flags_out, slk_read_detail* detail_out,
                                     slk_hit* hits_out) {
  if (!r || !opts || !span_off || !taxon_out || !flags_out || (n_spans && !spans) || (n_routed && (!send_idx || !taxa)))
    return slk_fail(SLK_E_INVALID, "bad arguments");
  if (hits_out && !detail_out) return slk_fail(SLK_E_INVALID, "hits_out needs detail_out");
  SLK_CU(cudaSetDevice(r->ctx->device));
  if (n_reads == 0) return SLK_OK;
  cudaStream_t st = r->ctx->stream;
  if (r->dense_cap < std::max<uint64_t>(n_spans, 1)) {
    cudaFree(r->d_dense); r->d_dense = nullptr; r->dense_cap = 0;
    const uint64_t cap = std::max<uint64_t>(n_spans + n_spans / 8, 1024);
    SLK_CU(cudaMalloc(&r->d_dense, cap * 2));
    r->dense_cap = cap;
  }
  uint16_t* d_dense = r->d_dense;
  auto done = [&](int rc) { return rc; };
  cudaMemsetAsync(d_dense, 0, std::max<uint64_t>(n_spans, 1) * 2, st);
  if (n_routed)
    unroute_kernel<<<(unsigned)((n_routed + 255) / 256), 256, 0, st>>>(taxa, send_idx, n_routed, r->d_r2d, (int32_t)r->tax->parents.size(),
                                                                        d_dense, r->d_err);
  resolve_spans_kernel<<<(n_reads + 127) / 128, 128, 0, st>>>(r->dt.view(), r->sp.k, spans, span_off, d_dense, n_reads, paired,
                                                            opts->confidence, opts->min_hit_groups, taxon_out, flags_out,
                                                            detail_out, hits_out, r->d_err + 0);
  uint32_t err = 0;
  if (cudaMemcpyAsync(&err, r->d_err, 4, cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess)
    return done(slk_fail(SLK_E_CUDA, "resolve failed: %s", cudaGetErrorString(cudaGetLastError())));
  if (err) {
    cudaMemset(r->d_err, 0, 4);
    return done(slk_fail(SLK_E_UNSUPPORTED, "a returned taxon is unknown to the resolver, or a fragment hit more than %d distinct taxa", SLK_KMAX));
  }
  return done(SLK_OK);
}

// ... and on the device (the distributed build keeps its records in HBM)
__global__ void __launch_bounds__(256) shard_of_records_kernel(slk_scan_params sp, const int64_t* __restrict__ id1, uint64_t n,
                                                               uint32_t world, uint8_t* __restrict__ out) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (uint8_t)slk_shard_of(slk_compress(sp, (uint64_t)id1[i]), world);
}
extern "C" int slk_shard_of_records_dev(slk_ctx* ctx, const slk_params* params, const int64_t* id1, uint64_t n, uint32_t world,
                                        uint8_t* shard_out) {
  if (!ctx || !params || world == 0 || world > 255 || (n && (!id1 || !shard_out))) return slk_fail(SLK_E_INVALID, "bad arguments");
  slk_scan_params sp;
  int rc = slk_make_scan_params_checked(params, &sp);
  if (rc != SLK_OK) return rc;
  SLK_CU(cudaSetDevice(ctx->device));
  if (n == 0) return SLK_OK;
  shard_of_records_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(sp, id1, n, world, shard_out);
  SLK_CU(cudaGetLastError());
  SLK_CU(cudaStreamSynchronize(ctx->stream));
  return SLK_OK;
}

// The export half of the distributed build: the records of a table grouped by the rank that owns their minimizer, in
// device memory. Pass 1 counts per owner, pass 2 writes (id1, raw taxon) to owner_start[d] + position; a block takes
// 1024 cells, counts in shared memory and touches the global cursors once per owner.
#define DUMP_MAX_WORLD 256
template <bool SCATTER>
__global__ void __launch_bounds__(256) dump_by_owner_kernel(slk_table_view tb, slk_scan_params sp, const int32_t* __restrict__ raw,
                                                            uint32_t world, unsigned long long* cursors, int64_t* __restrict__ id1,
                                                            int32_t* __restrict__ taxon, uint64_t cap) {
  __shared__ uint32_t s_cnt[DUMP_MAX_WORLD];
  __shared__ unsigned long long s_base[DUMP_MAX_WORLD];
  for (uint32_t d = threadIdx.x; d < world; d += 256) s_cnt[d] = 0;
  __syncthreads();
  const uint64_t ncell = tb.n_buckets * 4, i0 = (uint64_t)blockIdx.x * 1024;
  uint64_t cell[4];
  uint32_t dest[4], pos[4];
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const uint64_t i = i0 + (uint64_t)j * 256 + threadIdx.x;
    cell[j] = i < ncell ? tb.cells[i] : 0;
    if (cell[j]) { dest[j] = slk_shard_of(cell[j] >> 16, world); pos[j] = atomicAdd(&s_cnt[dest[j]], 1u); }
  }
  __syncthreads();
  for (uint32_t d = threadIdx.x; d < world; d += 256)
    if (s_cnt[d]) s_base[d] = atomicAdd(&cursors[d], (unsigned long long)s_cnt[d]);
  if (!SCATTER) return;
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 4; j++)
    if (cell[j]) {
      const unsigned long long o = s_base[dest[j]] + pos[j];
      if (o < cap) { id1[o] = (int64_t)slk_expand(sp, cell[j] >> 16); taxon[o] = raw[cell[j] & 0xffffu]; }
    }
}
extern "C" int slk_index_records_by_owner_dev(slk_index* idx, uint32_t world, int64_t* id1_out, int32_t* taxon_out, uint64_t cap,
                                              uint64_t* counts_host) {
  if (!idx || !counts_host || world == 0 || world > DUMP_MAX_WORLD) return slk_fail(SLK_E_INVALID, "bad arguments");
  slk_ctx* ctx = idx->ctx;
  SLK_CU(cudaSetDevice(ctx->device));
  for (uint32_t d = 0; d < world; d++) counts_host[d] = 0;
  if (idx->n_records == 0) return SLK_OK;
  if (!id1_out || !taxon_out || cap < idx->n_records) return slk_fail(SLK_E_NOSPACE, "records need room for %llu rows", (unsigned long long)idx->n_records);
  unsigned long long* d_cur = nullptr;
  SLK_CU(cudaMalloc(&d_cur, (size_t)world * 8));
  auto done = [&](int rc) { cudaFree(d_cur); return rc; };
  const uint64_t ncell = idx->table.n_buckets * 4;
  const unsigned grid = (unsigned)((ncell + 1023) / 1024);
  std::vector<unsigned long long> cnt(world), start(world);
  if (cudaMemsetAsync(d_cur, 0, (size_t)world * 8, ctx->stream) != cudaSuccess) return done(slk_fail(SLK_E_CUDA, "memset failed"));
  dump_by_owner_kernel<false><<<grid, 256, 0, ctx->stream>>>(idx->table, idx->sp, idx->dt.d_raw, world, d_cur, nullptr, nullptr, 0);
  if (cudaMemcpyAsync(cnt.data(), d_cur, (size_t)world * 8, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
      cudaStreamSynchronize(ctx->stream) != cudaSuccess)
    return done(slk_fail(SLK_E_CUDA, "owner count failed: %s", cudaGetErrorString(cudaGetLastError())));
  unsigned long long total = 0;
  for (uint32_t d = 0; d < world; d++) { counts_host[d] = cnt[d]; start[d] = total; total += cnt[d]; }
  if (total != idx->n_records) return done(slk_fail(SLK_E_CUDA, "table holds %llu records, expected %llu", total, (unsigned long long)idx->n_records));
  if (cudaMemcpyAsync(d_cur, start.data(), (size_t)world * 8, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess)
    return done(slk_fail(SLK_E_CUDA, "copy failed"));
  dump_by_owner_kernel<true><<<grid, 256, 0, ctx->stream>>>(idx->table, idx->sp, idx->dt.d_raw, world, d_cur, id1_out, taxon_out, cap);
  if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return done(slk_fail(SLK_E_CUDA, "record export failed: %s", cudaGetErrorString(cudaGetLastError())));
  return done(SLK_OK);
}

// ---------------------------------------------------------------------------------------------- NVLink mailbox
// The two exchanges of the split path as stores into PEER memory, fused into the kernels on either side of them
// (no NCCL, no host in the loop): the route kernel groups the keys of a block by owner in shared memory and stores every
// run straight into the owner's inbox over NVLink; the owner's lookup kernel stores every taxon straight into the
// asker's reply area. Completion travels as one 8-byte flag per (source, owner) pair, written after the data by a
// one-block kernel on the same stream (epoch << 32 | key count); the consuming kernel's blocks spin on the flag of the
// peer whose data they read. Every rank's mailbox is ONE allocation with the same layout:
//   flag1[world] | flag2[world] | keys_in[world][cap] (u64) | taxa_back[world][cap] (i32)
// keys_in[s] = keys rank s wants looked up here; taxa_back[d] = answers of owner d, in the order the keys were sent.
#define MBX_MAX_WORLD 64
#define MBX_ROUTE_ITEMS 4          // spans per thread of the route kernel (1024 per block: runs of ~1 KB per owner at 8 ranks)
#define MBX_SPIN_NS 10000000000ull // a peer that does not show up within 10 s is reported as an error instead of a hang
struct mbx_layout {
  uint32_t world; uint64_t cap;
  SLK_HD size_t flag1_off() const { return 0; }
  SLK_HD size_t flag2_off() const { return (size_t)world * 8; }
  SLK_HD size_t keys_off() const { return (((size_t)world * 16) + 255) & ~(size_t)255; }
  SLK_HD size_t taxa_off() const { return keys_off() + (size_t)world * cap * 8; }
  SLK_HD size_t bytes() const { return taxa_off() + (size_t)world * cap * 4; }
};
struct slk_mailbox {
  slk_ctx* ctx;
  uint32_t rank, world;
  uint64_t cap;
  mbx_layout lay;
  uint8_t* base = nullptr;                 // this rank's mailbox (peers store into it)
  std::vector<uint8_t*> peer;              // every rank's mailbox as seen from here (peer[rank] == base)
  std::vector<bool> opened;                // peer[i] came from cudaIpcOpenMemHandle
  uint8_t** d_peer = nullptr;
  uint32_t* d_send_idx = nullptr;          // [world][cap]: span index of every key sent, per owner
  unsigned long long* d_cursors = nullptr; // [world]: keys sent to every owner in the current batch
  uint32_t* d_err = nullptr;               // bit 0: inbox overflow, bit 1: a peer timed out, bit 2: bad taxon / overflow of taxa
  uint16_t* d_dense = nullptr; uint64_t dense_cap = 0;
  uint32_t epoch = 0;
  uint32_t blocks_per_sm = 8;              // size of the lookup / unroute grids (slk_mailbox_set_blocks_per_sm)
  bool connected = false;
  cudaStream_t copy_stream = nullptr;      // slk_mailbox_resolve_wait: the results leave here, past whatever the main stream
  cudaEvent_t ev_resolved = nullptr;       // has queued behind the resolve kernels (the next batch's route and lookups)
};

__device__ __forceinline__ unsigned long long mbx_ld_flag(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void mbx_st_flag(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long mbx_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// waits until the flag carries `epoch`; returns its low word (the count), 0 on time-out
__device__ __forceinline__ uint32_t mbx_wait(const unsigned long long* flag, uint32_t epoch, uint32_t* err) {
  const unsigned long long t0 = mbx_now();
  for (;;) {
    const unsigned long long v = mbx_ld_flag(flag);
    const uint32_t ep = (uint32_t)(v >> 32);
    if (ep == epoch) return (uint32_t)v;
    // a NEWER epoch means the peer overwrote the flag before it was read here: a protocol violation (every rank consumes
    // every owner's flag2 before it routes again, so this cannot happen), reported at once instead of after the time-out
    if ((int32_t)(ep - epoch) > 0 || mbx_now() - t0 > MBX_SPIN_NS) { atomicOr(err, 2u); return 0; }
    __nanosleep(256);
  }
}

// Route + send. A block takes 1024 consecutive span words, counts its sequence spans per owner in shared memory,
// reserves a range of every owner's inbox slice with one atomic per owner, orders its keys by owner in shared memory and
// stores the runs: consecutive threads store consecutive 8-byte words of one peer's memory.
__global__ void __launch_bounds__(256) mbx_route_kernel(const uint64_t* __restrict__ spans, uint64_t n, mbx_layout lay, uint32_t rank,
                                                        unsigned long long* cursors, uint8_t* const* __restrict__ peers,
                                                        uint32_t* __restrict__ send_idx, uint32_t* err) {
  constexpr uint32_t PER_BLOCK = 256 * MBX_ROUTE_ITEMS;
  __shared__ uint32_t s_cnt[MBX_MAX_WORLD], s_off[MBX_MAX_WORLD + 1];
  __shared__ unsigned long long s_base[MBX_MAX_WORLD];
  __shared__ uint64_t s_key[PER_BLOCK];
  __shared__ uint32_t s_idx[PER_BLOCK];
  __shared__ uint8_t s_dest[PER_BLOCK];
  const uint32_t world = lay.world;
  if (threadIdx.x < world) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  const uint64_t i0 = (uint64_t)blockIdx.x * PER_BLOCK;
  uint64_t ck[MBX_ROUTE_ITEMS];
  uint32_t dest[MBX_ROUTE_ITEMS], pos[MBX_ROUTE_ITEMS];
#pragma unroll
  for (int j = 0; j < MBX_ROUTE_ITEMS; j++) {
    const uint64_t i = i0 + (uint64_t)j * 256 + threadIdx.x;
    dest[j] = 0xffffffffu;
    if (i < n) {
      const uint64_t w = spans[i];
      if (SLK_SPAN_TYPE(w) == SLK_E_SEQ) {
        ck[j] = SLK_SPAN_KEY(w);
        dest[j] = slk_shard_of(ck[j], world);
        pos[j] = atomicAdd(&s_cnt[dest[j]], 1u);
      }
    }
  }
  __syncthreads();
  if (threadIdx.x < world && s_cnt[threadIdx.x]) s_base[threadIdx.x] = atomicAdd(&cursors[threadIdx.x], (unsigned long long)s_cnt[threadIdx.x]);
  if (threadIdx.x == 0) {
    uint32_t o = 0;
    for (uint32_t d = 0; d < world; d++) { s_off[d] = o; o += s_cnt[d]; }
    s_off[world] = o;
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < MBX_ROUTE_ITEMS; j++)
    if (dest[j] != 0xffffffffu) {
      const uint32_t q = s_off[dest[j]] + pos[j];
      s_key[q] = ck[j]; s_idx[q] = (uint32_t)(i0 + (uint64_t)j * 256 + threadIdx.x); s_dest[q] = (uint8_t)dest[j];
    }
  __syncthreads();
  const uint32_t total = s_off[world];
  for (uint32_t q = threadIdx.x; q < total; q += 256) {
    const uint32_t d = s_dest[q];
    const unsigned long long p = s_base[d] + (q - s_off[d]);
    if (p < lay.cap) {
      reinterpret_cast<uint64_t*>(peers[d] + lay.keys_off())[(uint64_t)rank * lay.cap + p] = s_key[q];
      send_idx[(uint64_t)d * lay.cap + p] = s_idx[q];
    } else {
      atomicOr(err, 1u);
    }
  }
}
// after the route kernel: tells every owner how many keys are waiting for it
__global__ void mbx_signal_keys_kernel(mbx_layout lay, uint32_t rank, uint32_t epoch, const unsigned long long* cursors,
                                       uint8_t* const* __restrict__ peers) {
  const uint32_t d = threadIdx.x;
  if (d >= lay.world) return;
  unsigned long long c = cursors[d];
  if (c > lay.cap) c = lay.cap;
  __threadfence_system();
  mbx_st_flag(reinterpret_cast<unsigned long long*>(peers[d] + lay.flag1_off()) + rank, ((unsigned long long)epoch << 32) | c);
}
// Owner side: `per` blocks per source; a block waits for the keys of its source s, then strides over them 256 at a time:
// look up, store the raw taxon into s's reply area.
__global__ void __launch_bounds__(256) mbx_probe_kernel(slk_table_view tb, slk_tax_view tx, mbx_layout lay, uint32_t rank, uint32_t epoch,
                                                        uint32_t per, const uint8_t* __restrict__ base,
                                                        uint8_t* const* __restrict__ peers, uint32_t* err) {
  __shared__ uint32_t s_n;
  const uint32_t s = blockIdx.x / per, c0 = blockIdx.x % per;
  if (threadIdx.x == 0) s_n = mbx_wait(reinterpret_cast<const unsigned long long*>(base + lay.flag1_off()) + s, epoch, err);
  __syncthreads();
  const uint64_t n = s_n;
  const uint64_t* keys = reinterpret_cast<const uint64_t*>(base + lay.keys_off()) + (uint64_t)s * lay.cap;
  int32_t* back = reinterpret_cast<int32_t*>(peers[s] + lay.taxa_off()) + (uint64_t)rank * lay.cap;
  for (uint64_t j = (uint64_t)c0 * 256 + threadIdx.x; j < n; j += (uint64_t)per * 256) {
    const uint32_t d = slk_probe(tb, __ldcg(keys + j));
    back[j] = d ? tx.raw[d] : 0;
  }
}
__global__ void mbx_signal_taxa_kernel(mbx_layout lay, uint32_t rank, uint32_t epoch, uint8_t* const* __restrict__ peers) {
  const uint32_t s = threadIdx.x;
  if (s >= lay.world) return;
  __threadfence_system();
  mbx_st_flag(reinterpret_cast<unsigned long long*>(peers[s] + lay.flag2_off()) + rank, ((unsigned long long)epoch << 32) | 1ull);
}
// Asker side: `per` blocks per owner; a block waits for the answers of its owner d and scatters them to their spans as
// dense labels.
__global__ void __launch_bounds__(256) mbx_unroute_kernel(mbx_layout lay, uint32_t epoch, uint32_t per, const uint8_t* __restrict__ base,
                                                          const unsigned long long* __restrict__ cursors,
                                                          const uint32_t* __restrict__ send_idx, const uint16_t* __restrict__ raw2dense,
                                                          int32_t n_tax, uint16_t* __restrict__ dense, uint32_t* err) {
  const uint32_t d = blockIdx.x / per, c0 = blockIdx.x % per;
  unsigned long long n = cursors[d];
  if (n > lay.cap) n = lay.cap;
  // Block 0 of every owner ALWAYS consumes the owner's flag2, even when nothing was sent to it: the owner raises flag2 only
  // after its lookup kernel has read this rank's flag1 (and count) of the batch, so once this kernel is through, the next
  // route may overwrite flag1 and the keys, and the mailbox may be freed, without racing with any peer.
  if ((uint64_t)c0 * 256 >= n && c0 != 0) return;   // nothing of this block's share was sent to d
  if (threadIdx.x == 0) mbx_wait(reinterpret_cast<const unsigned long long*>(base + lay.flag2_off()) + d, epoch, err);
  __syncthreads();
  const int32_t* back = reinterpret_cast<const int32_t*>(base + lay.taxa_off()) + (uint64_t)d * lay.cap;
  const uint32_t* idx = send_idx + (uint64_t)d * lay.cap;
  for (uint64_t j = (uint64_t)c0 * 256 + threadIdx.x; j < n; j += (uint64_t)per * 256) {
    const int32_t t = __ldcg(back + j);
    uint16_t dn = 0;
    if (t != 0) {
      if (t < 0 || t >= n_tax || (dn = raw2dense[t]) == 0) { atomicOr(err, 4u); dn = 0; }
    }
    dense[idx[j]] = dn;
  }
}

extern "C" int slk_mailbox_create(slk_ctx* ctx, uint32_t rank, uint32_t world, uint64_t cap, slk_mailbox** out, uint8_t* handle_out) {
  if (!ctx || !out || world == 0 || world > MBX_MAX_WORLD || rank >= world || cap == 0 || cap > 0xffffffffull)
    return slk_fail(SLK_E_INVALID, "bad arguments (1 <= world <= %d, rank < world, 0 < cap < 2^32)", MBX_MAX_WORLD);
  SLK_CU(cudaSetDevice(ctx->device));
  slk_mailbox* m = new (std::nothrow) slk_mailbox;
  if (!m) return slk_fail(SLK_E_NOMEM, "host allocation failed");
  m->ctx = ctx; m->rank = rank; m->world = world; m->cap = cap;
  m->lay.world = world; m->lay.cap = cap;
  m->peer.assign(world, nullptr); m->opened.assign(world, false);
  cudaError_t e = cudaMalloc(&m->base, m->lay.bytes());
  if (e == cudaSuccess) e = cudaMemset(m->base, 0, m->lay.keys_off());
  if (e == cudaSuccess) e = cudaMalloc(&m->d_peer, (size_t)world * sizeof(uint8_t*));
  if (e == cudaSuccess) e = cudaMalloc(&m->d_send_idx, (size_t)world * cap * 4);
  if (e == cudaSuccess) e = cudaMalloc(&m->d_cursors, (size_t)world * 8);
  if (e == cudaSuccess) e = cudaMemset(m->d_cursors, 0, (size_t)world * 8);
  if (e == cudaSuccess) e = cudaMalloc(&m->d_err, 4);
  if (e == cudaSuccess) e = cudaMemset(m->d_err, 0, 4);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&m->copy_stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&m->ev_resolved, cudaEventDisableTiming);
  if (e == cudaSuccess && handle_out) {
    cudaIpcMemHandle_t h;
    static_assert(sizeof(h) == SLK_IPC_HANDLE_BYTES, "IPC handle size");
    e = cudaIpcGetMemHandle(&h, m->base);
    if (e == cudaSuccess) memcpy(handle_out, &h, sizeof(h));
  }
  if (e != cudaSuccess) {
    const int rc = slk_fail(e == cudaErrorMemoryAllocation ? SLK_E_NOMEM : SLK_E_CUDA, "mailbox allocation failed: %s", cudaGetErrorString(e));
    slk_mailbox_destroy(m);
    return rc;
  }
  m->peer[rank] = m->base;
  *out = m;
  return SLK_OK;
}
// one wave of 256-thread blocks over the whole chip (8 per SM), shared out evenly among the peers
static uint32_t mbx_blocks_per_peer(const slk_mailbox* m) {
  // SLK_MBX_BLOCKS_PER_SM (1..8) overrides the mailbox's own setting (tuning runs)
  static const int env_per_sm = [] { const char* e = getenv("SLK_MBX_BLOCKS_PER_SM"); int v = e ? atoi(e) : 0; return v < 0 ? 0 : v > 8 ? 8 : v; }();
  const uint32_t per_sm = env_per_sm ? (uint32_t)env_per_sm : m->blocks_per_sm;
  const uint32_t per = (uint32_t)m->ctx->sm_count * per_sm / m->world;
  const uint32_t most = (uint32_t)((m->cap + 255) / 256);
  return std::max(1u, std::min(per, most));
}
static int mbx_upload_peers(slk_mailbox* m) {
  SLK_CU(cudaMemcpy(m->d_peer, m->peer.data(), (size_t)m->world * sizeof(uint8_t*), cudaMemcpyHostToDevice));
  m->connected = true;
  return SLK_OK;
}
// one process per GPU: handles = world x SLK_IPC_HANDLE_BYTES, gathered from slk_mailbox_create on every rank
extern "C" int slk_mailbox_connect(slk_mailbox* m, const uint8_t* handles) {
  if (!m || !handles) return slk_fail(SLK_E_INVALID, "bad arguments");
  SLK_CU(cudaSetDevice(m->ctx->device));
  for (uint32_t r = 0; r < m->world; r++) {
    if (r == m->rank || m->peer[r]) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, handles + (size_t)r * SLK_IPC_HANDLE_BYTES, sizeof(h));
    void* p = nullptr;
    SLK_CU(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    m->peer[r] = (uint8_t*)p; m->opened[r] = true;
  }
  return mbx_upload_peers(m);
}
// one process driving all ranks (tests; several ranks may share a device): peers by pointer
extern "C" int slk_mailbox_connect_local(slk_mailbox* const* boxes, uint32_t world) {
  if (!boxes || world == 0) return slk_fail(SLK_E_INVALID, "bad arguments");
  for (uint32_t r = 0; r < world; r++)
    if (!boxes[r] || boxes[r]->world != world || boxes[r]->rank != r || boxes[r]->cap != boxes[0]->cap)
      return slk_fail(SLK_E_INVALID, "mailbox %u does not belong to this group", r);
  for (uint32_t r = 0; r < world; r++) {
    SLK_CU(cudaSetDevice(boxes[r]->ctx->device));
    for (uint32_t q = 0; q < world; q++) {
      if (boxes[q]->ctx->device != boxes[r]->ctx->device) {
        int can = 0;
        SLK_CU(cudaDeviceCanAccessPeer(&can, boxes[r]->ctx->device, boxes[q]->ctx->device));
        if (!can) return slk_fail(SLK_E_UNSUPPORTED, "device %d cannot access device %d", boxes[r]->ctx->device, boxes[q]->ctx->device);
        cudaError_t e = cudaDeviceEnablePeerAccess(boxes[q]->ctx->device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) SLK_CU(e);
        cudaGetLastError();
      }
      boxes[r]->peer[q] = boxes[q]->base;
    }
    int rc = mbx_upload_peers(boxes[r]);
    if (rc != SLK_OK) return rc;
  }
  return SLK_OK;
}
extern "C" void slk_mailbox_destroy(slk_mailbox* m) {
  if (!m) return;
  cudaSetDevice(m->ctx->device);
  cudaStreamSynchronize(m->ctx->stream);
  for (uint32_t r = 0; r < m->world; r++)
    if (m->opened[r] && m->peer[r]) cudaIpcCloseMemHandle(m->peer[r]);
  cudaFree(m->base); cudaFree(m->d_peer); cudaFree(m->d_send_idx); cudaFree(m->d_cursors); cudaFree(m->d_err); cudaFree(m->d_dense);
  if (m->copy_stream) { cudaStreamSynchronize(m->copy_stream); cudaStreamDestroy(m->copy_stream); }
  if (m->ev_resolved) cudaEventDestroy(m->ev_resolved);
  delete m;
}
// 256-thread blocks per SM of the lookup and unroute kernels (1..8, default 8 = every thread slot). A caller that scans
// the next batch meanwhile (slk_scan_spans_dev runs on its own stream) leaves room for it: measured best at 4.
extern "C" int slk_mailbox_set_blocks_per_sm(slk_mailbox* m, uint32_t blocks_per_sm) {
  if (!m || blocks_per_sm < 1 || blocks_per_sm > 8) return slk_fail(SLK_E_INVALID, "blocks per SM must be 1..8");
  m->blocks_per_sm = blocks_per_sm;
  return SLK_OK;
}
// step 1 (asynchronous): this rank's sequence-span keys go to their owners' inboxes
extern "C" int slk_mailbox_route(slk_mailbox* m, const uint64_t* spans, uint64_t n_spans) {
  if (!m || !m->connected || (n_spans && !spans)) return slk_fail(SLK_E_INVALID, "bad arguments (is the mailbox connected?)");
  if (n_spans > 0xffffffffull) return slk_fail(SLK_E_UNSUPPORTED, "more than 2^32 spans in one batch");
  SLK_CU(cudaSetDevice(m->ctx->device));
  cudaStream_t st = m->ctx->stream;
  m->epoch++;
  SLK_CU(cudaMemsetAsync(m->d_cursors, 0, (size_t)m->world * 8, st));
  if (n_spans) {
    const uint64_t per = 256ull * MBX_ROUTE_ITEMS;
    mbx_route_kernel<<<(unsigned)((n_spans + per - 1) / per), 256, 0, st>>>(spans, n_spans, m->lay, m->rank, m->d_cursors, m->d_peer,
                                                                            m->d_send_idx, m->d_err);
  }
  mbx_signal_keys_kernel<<<1, MBX_MAX_WORLD, 0, st>>>(m->lay, m->rank, m->epoch, m->d_cursors, m->d_peer);
  SLK_CU(cudaGetLastError());
  return SLK_OK;
}
// step 2 (asynchronous): the owner's half of the join for the keys of every rank, answers stored into the askers
extern "C" int slk_mailbox_probe(slk_mailbox* m, slk_index* idx) {
  if (!m || !m->connected || !idx || idx->ctx != m->ctx) return slk_fail(SLK_E_INVALID, "bad arguments (index and mailbox must share a context)");
  SLK_CU(cudaSetDevice(m->ctx->device));
  cudaStream_t st = m->ctx->stream;
  const uint32_t per = mbx_blocks_per_peer(m);
  mbx_probe_kernel<<<m->world * per, 256, 0, st>>>(idx->table, idx->dt.view(), m->lay, m->rank, m->epoch, per, m->base, m->d_peer, m->d_err);
  mbx_signal_taxa_kernel<<<1, MBX_MAX_WORLD, 0, st>>>(m->lay, m->rank, m->epoch, m->d_peer);
  SLK_CU(cudaGetLastError());
  return SLK_OK;
}
// step 3a (asynchronous): slk_resolve_spans_dev on the answers in the reply area. The next batch's slk_mailbox_route /
// slk_mailbox_probe may be issued right after this call: they queue behind the resolve kernels on the library's stream,
// which is all the protocol needs (this rank's unroute kernel has consumed every owner's flag before its next route runs).
extern "C" int slk_mailbox_resolve_async(slk_mailbox* m, slk_resolver* r, const slk_classify_opts* opts, const uint64_t* spans,
                                         const uint64_t* span_off, uint64_t n_spans, uint32_t n_reads, int paired,
                                         int32_t* taxon_out, uint8_t* flags_out, slk_read_detail* detail_out, slk_hit* hits_out) {
  if (!m || !m->connected || !r || r->ctx != m->ctx || !opts || !span_off || !taxon_out || !flags_out || (n_spans && !spans))
    return slk_fail(SLK_E_INVALID, "bad arguments (resolver and mailbox must share a context)");
  if (hits_out && !detail_out) return slk_fail(SLK_E_INVALID, "hits_out needs detail_out");
  SLK_CU(cudaSetDevice(m->ctx->device));
  cudaStream_t st = m->ctx->stream;
  if (m->dense_cap < std::max<uint64_t>(n_spans, 1)) {
    SLK_CU(cudaStreamSynchronize(st));
    cudaFree(m->d_dense); m->d_dense = nullptr; m->dense_cap = 0;
    const uint64_t cap = std::max<uint64_t>(n_spans + n_spans / 8, 1024);
    SLK_CU(cudaMalloc(&m->d_dense, cap * 2));
    m->dense_cap = cap;
  }
  SLK_CU(cudaMemsetAsync(m->d_dense, 0, std::max<uint64_t>(n_spans, 1) * 2, st));
  const uint32_t per = mbx_blocks_per_peer(m);
  mbx_unroute_kernel<<<m->world * per, 256, 0, st>>>(m->lay, m->epoch, per, m->base, m->d_cursors, m->d_send_idx, r->d_r2d,
                                                        (int32_t)r->tax->parents.size(), m->d_dense, m->d_err);
  if (n_reads)
    resolve_spans_kernel<<<(n_reads + 127) / 128, 128, 0, st>>>(r->dt.view(), r->sp.k, spans, span_off, m->d_dense, n_reads, paired,
                                                                opts->confidence, opts->min_hit_groups, taxon_out, flags_out,
                                                                detail_out, hits_out, r->d_err);
  SLK_CU(cudaGetLastError());
  SLK_CU(cudaEventRecord(m->ev_resolved, st));
  return SLK_OK;
}
// step 3b: returns when the results of the last slk_mailbox_resolve_async are complete. taxon_host / flags_host (pinned
// memory, or NULL) receive n_reads results from taxon_dev / flags_dev on a stream of their own, past whatever has been
// queued behind the resolve kernels meanwhile.
extern "C" int slk_mailbox_resolve_wait(slk_mailbox* m, slk_resolver* r, uint32_t n_reads, const int32_t* taxon_dev,
                                        const uint8_t* flags_dev, int32_t* taxon_host, uint8_t* flags_host) {
  if (!m || !r || ((taxon_host || flags_host) && (!taxon_dev || !flags_dev))) return slk_fail(SLK_E_INVALID, "bad arguments");
  SLK_CU(cudaSetDevice(m->ctx->device));
  cudaStream_t st = m->copy_stream;
  uint32_t err = 0, rerr = 0;
  bool ok = cudaStreamWaitEvent(st, m->ev_resolved, 0) == cudaSuccess;
  if (ok && taxon_host && n_reads) ok = cudaMemcpyAsync(taxon_host, taxon_dev, (size_t)n_reads * 4, cudaMemcpyDeviceToHost, st) == cudaSuccess;
  if (ok && flags_host && n_reads) ok = cudaMemcpyAsync(flags_host, flags_dev, (size_t)n_reads, cudaMemcpyDeviceToHost, st) == cudaSuccess;
  ok = ok && cudaMemcpyAsync(&err, m->d_err, 4, cudaMemcpyDeviceToHost, st) == cudaSuccess &&
       cudaMemcpyAsync(&rerr, r->d_err, 4, cudaMemcpyDeviceToHost, st) == cudaSuccess && cudaStreamSynchronize(st) == cudaSuccess;
  if (!ok) return slk_fail(SLK_E_CUDA, "mailbox resolve failed: %s", cudaGetErrorString(cudaGetLastError()));
  if (err || rerr) {
    cudaStreamSynchronize(m->ctx->stream);
    cudaMemset(m->d_err, 0, 4); cudaMemset(r->d_err, 0, 4);
    if (err & 2u) return slk_fail(SLK_E_CUDA, "a peer did not deliver its part of the exchange within 10 s");
    if (err & 1u) return slk_fail(SLK_E_NOSPACE, "more than %llu keys for one owner in a batch: create the mailbox with a larger cap", (unsigned long long)m->cap);
    return slk_fail(SLK_E_UNSUPPORTED, "a returned taxon is unknown to the resolver, or a fragment hit more than %d distinct taxa", SLK_KMAX);
  }
  return SLK_OK;
}
// step 3 in one call: returns when the results are complete (device memory)
extern "C" int slk_mailbox_resolve(slk_mailbox* m, slk_resolver* r, const slk_classify_opts* opts, const uint64_t* spans,
                                   const uint64_t* span_off, uint64_t n_spans, uint32_t n_reads, int paired, int32_t* taxon_out,
                                   uint8_t* flags_out, slk_read_detail* detail_out, slk_hit* hits_out) {
  const int rc = slk_mailbox_resolve_async(m, r, opts, spans, span_off, n_spans, n_reads, paired, taxon_out, flags_out, detail_out, hits_out);
  return rc != SLK_OK ? rc : slk_mailbox_resolve_wait(m, r, n_reads, nullptr, nullptr, nullptr, nullptr);
}

// owner of every record (id1 = the uncompressed minimizer of the Parquet column), computed on the host
extern "C" int slk_shard_of_records(const slk_params* params, const int64_t* id1, uint64_t n, uint32_t world, uint8_t* shard_out) {
  if (!params || world == 0 || world > 255 || (n && (!id1 || !shard_out))) return slk_fail(SLK_E_INVALID, "bad arguments");
  slk_scan_params sp;
  int rc = slk_make_scan_params_checked(params, &sp);
  if (rc != SLK_OK) return rc;
  for (uint64_t i = 0; i < n; i++) shard_out[i] = (uint8_t)slk_shard_of(slk_compress(sp, (uint64_t)id1[i]), world);
  return SLK_OK;
}

// ---------------------------------------------------------------------------------------------- Bracken weights
// slacken/BrackenWeights.scala:312-354 on the device: scan every genome fragment into its taxon hits (two passes: count,
// emit), look the super-mers up (one thread per hit), then one thread per fragment slides the read window over the hits
// (the window is inherently sequential, and the reference's quasi-hit ordinals make it non-local) and counts the
// destination taxon of every read. Output: (destination, source, reads) triples per fragment; the caller adds them up.
__global__ void __launch_bounds__(256) bracken_probe_kernel(slk_table_view tb, slk_bhit* __restrict__ hits, uint64_t n) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (hits[i].flags & SLK_BHIT_SEQ) hits[i].taxon = slk_probe(tb, hits[i].key);
}
#define BRACKEN_DESTS 64
__global__ void __launch_bounds__(32) bracken_window_kernel(slk_tax_view tx, const slk_bhit* __restrict__ hits,
                                                            const uint64_t* __restrict__ hit_off, const uint64_t* __restrict__ frag_off,
                                                            const int32_t* __restrict__ frag_taxon, uint32_t n_frag, uint32_t read_len,
                                                            uint32_t k, slk_bracken_triple* __restrict__ out, uint64_t cap,
                                                            unsigned long long* cursor, uint32_t* error_flag) {
  const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= n_frag) return;
  uint32_t dk[BRACKEN_DESTS];
  unsigned long long dv[BRACKEN_DESTS];
  uint32_t nd = 0;
  bool lost = false;
  const bool ok = slk_bracken_window(tx, hits + hit_off[f], (uint32_t)(hit_off[f + 1] - hit_off[f]), (uint32_t)(frag_off[f + 1] - frag_off[f]),
                                     read_len, k, [&](uint32_t d) {
    for (uint32_t i = 0; i < nd; i++)
      if (dk[i] == d) { dv[i]++; return; }
    if (nd == BRACKEN_DESTS) { lost = true; return; }
    dk[nd] = d; dv[nd] = 1; nd++;
  });
  if (!ok || lost) atomicExch(error_flag, 1u);
  if (nd == 0) return;
  const unsigned long long o = atomicAdd(cursor, (unsigned long long)nd);
  for (uint32_t i = 0; i < nd; i++)
    if (o + i < cap) {
      slk_bracken_triple t;
      t.dest = tx.raw[dk[i]]; t.source = frag_taxon[f]; t.reads = dv[i];
      out[o + i] = t;
    }
}
#define SLK_DISPATCH_BRACKEN(w, a)                       \
  switch (w) {                                           \
    case 1: slk_launch_bracken_scan_w1(a); break; case 2: slk_launch_bracken_scan_w2(a); break; \
    case 3: slk_launch_bracken_scan_w3(a); break; case 4: slk_launch_bracken_scan_w4(a); break; \
    case 5: slk_launch_bracken_scan_w5(a); break; case 6: slk_launch_bracken_scan_w6(a); break; \
    case 7: slk_launch_bracken_scan_w7(a); break; default: slk_launch_bracken_scan_w8(a); break; \
  }

extern "C" int slk_bracken_weights(slk_index* idx, const uint8_t* bases, const uint64_t* frag_off, const int32_t* frag_taxon,
                                   uint32_t n_frag, uint32_t read_len, slk_bracken_triple* out, uint64_t cap, uint64_t* n_out) {
  if (!idx || !n_out || (n_frag && (!bases || !frag_off || !frag_taxon)) || (cap && !out)) return slk_fail(SLK_E_INVALID, "bad arguments");
  if (read_len < (uint32_t)idx->sp.k) return slk_fail(SLK_E_INVALID, "read length %u is shorter than k = %d", read_len, idx->sp.k);
  *n_out = 0;
  if (n_frag == 0) return SLK_OK;
  slk_ctx* ctx = idx->ctx;
  SLK_CU(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const uint64_t total = frag_off[n_frag] - frag_off[0];
  for (uint32_t f = 0; f < n_frag; f++)
    if (frag_off[f + 1] - frag_off[f] > 0xfffffff0ull) return slk_fail(SLK_E_UNSUPPORTED, "fragment %u is longer than 2^32 bases", f);
  uint8_t* d_bases = nullptr; uint64_t *d_foff = nullptr, *d_hoff = nullptr; int32_t* d_ftax = nullptr; slk_bhit* d_hits = nullptr;
  slk_bracken_triple* d_out = nullptr; unsigned long long* d_cur = nullptr; uint32_t* d_err = nullptr;
  auto done = [&](int rc) {
    cudaFree(d_bases); cudaFree(d_foff); cudaFree(d_hoff); cudaFree(d_ftax); cudaFree(d_hits); cudaFree(d_out); cudaFree(d_cur); cudaFree(d_err);
    return rc;
  };
#define BCU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return done(slk_fail(e_ == cudaErrorMemoryAllocation ? SLK_E_NOMEM : SLK_E_CUDA, \
    "%s failed: %s", #call, cudaGetErrorString(e_))); } while (0)
  std::vector<uint64_t> rel(n_frag + 1);
  for (uint32_t f = 0; f <= n_frag; f++) rel[f] = frag_off[f] - frag_off[0];
  BCU(cudaMalloc(&d_bases, total + 64)); BCU(cudaMalloc(&d_foff, ((size_t)n_frag + 1) * 8)); BCU(cudaMalloc(&d_hoff, ((size_t)n_frag + 1) * 8));
  BCU(cudaMalloc(&d_ftax, (size_t)n_frag * 4)); BCU(cudaMalloc(&d_cur, 8)); BCU(cudaMalloc(&d_err, 4));
  BCU(cudaMemcpyAsync(d_bases, bases + frag_off[0], total, cudaMemcpyHostToDevice, st));
  BCU(cudaMemcpyAsync(d_foff, rel.data(), ((size_t)n_frag + 1) * 8, cudaMemcpyHostToDevice, st));
  BCU(cudaMemcpyAsync(d_ftax, frag_taxon, (size_t)n_frag * 4, cudaMemcpyHostToDevice, st));
  BCU(cudaMemsetAsync(d_hoff, 0, ((size_t)n_frag + 1) * 8, st)); BCU(cudaMemsetAsync(d_cur, 0, 8, st)); BCU(cudaMemsetAsync(d_err, 0, 4, st));
  slk_bracken_scan_args a;
  a.sp = idx->sp; a.bases = d_bases; a.frag_off = d_foff; a.n_frag = n_frag; a.hit_off = d_hoff; a.hits = nullptr; a.stream = st;
  SLK_DISPATCH_BRACKEN(a.sp.w, a);
  BCU(cudaGetLastError());
  if (slk_exclusive_scan_u64(d_hoff, (uint64_t)n_frag + 1, st) != 0) return done(slk_fail(SLK_E_CUDA, "prefix sum of the hit counts failed"));
  uint64_t n_hits = 0;
  BCU(cudaMemcpy(&n_hits, d_hoff + n_frag, 8, cudaMemcpyDeviceToHost));
  BCU(cudaMalloc(&d_hits, std::max<uint64_t>(n_hits, 1) * sizeof(slk_bhit)));
  a.hits = d_hits;
  SLK_DISPATCH_BRACKEN(a.sp.w, a);
  BCU(cudaGetLastError());
  if (n_hits) bracken_probe_kernel<<<(unsigned)((n_hits + 255) / 256), 256, 0, st>>>(idx->table, d_hits, n_hits);
  BCU(cudaGetLastError());
  const uint64_t dcap = std::max<uint64_t>(cap, 1);
  BCU(cudaMalloc(&d_out, dcap * sizeof(slk_bracken_triple)));
  bracken_window_kernel<<<(n_frag + 31) / 32, 32, 0, st>>>(idx->dt.view(), d_hits, d_hoff, d_foff, d_ftax, n_frag, read_len, (uint32_t)idx->sp.k,
                                                         d_out, cap, d_cur, d_err);
  BCU(cudaGetLastError());
  unsigned long long used = 0;
  uint32_t err = 0;
  BCU(cudaMemcpyAsync(&used, d_cur, 8, cudaMemcpyDeviceToHost, st));
  BCU(cudaMemcpyAsync(&err, d_err, 4, cudaMemcpyDeviceToHost, st));
  BCU(cudaStreamSynchronize(st));
  *n_out = used;
  if (err) return done(slk_fail(SLK_E_UNSUPPORTED, "a read window held more than %d taxa, or a fragment more than %d destination taxa", SLK_KMAX, BRACKEN_DESTS));
  if (used > cap) return done(slk_fail(SLK_E_NOSPACE, "out needs room for %llu triples", used));
  if (used) BCU(cudaMemcpy(out, d_out, used * sizeof(slk_bracken_triple), cudaMemcpyDeviceToHost));
#undef BCU
  return done(SLK_OK);
}
