// slk_inst.cu -- one instantiation of the W-templated kernels; compiled once per window width with -DSLK_W=n.
#include "slk_kernels.cuh"

#ifndef SLK_W
#error "compile with -DSLK_W=<k-m+1>"
#endif
#define SLK_CAT_(a, b) a##b
#define SLK_CAT(a, b) SLK_CAT_(a, b)

#define SLK_CLS_ARGS(a, hb, hs, hc)                                                                                     \
  a.sp, a.tb, a.tx, a.bases1, a.off1, a.shift1, a.bases2, a.off2, a.shift2, a.mask1, a.len1, a.mask2, a.len2, a.n_reads,  \
      a.confidence, a.min_hit_groups, a.taxon_out, a.flags_out, a.detail_out, hb, hs, hc, a.hits_cursor, a.counts,         \
      a.error_flag, a.stats

// the tiles of a block (71 KB) exceed the static shared-memory limit: opt in once per instantiation and device
#define SLK_CLS_LAUNCH(HITS, PACKED, hb, hs, hc)                                                                       \
  do {                                                                                                                 \
    static bool done[64] = {};                                                                                         \
    int dev = 0;                                                                                                       \
    cudaGetDevice(&dev);                                                                                               \
    if (dev < 0 || dev >= 64 || !done[dev]) {                                                                          \
      cudaFuncSetAttribute(classify_kernel<SLK_W, HITS, PACKED>, cudaFuncAttributeMaxDynamicSharedMemorySize,          \
                           (int)SLK_SMEM_BYTES);                                                                       \
      cudaFuncSetAttribute(classify_kernel<SLK_W, HITS, PACKED>, cudaFuncAttributePreferredSharedMemoryCarveout,       \
                           cudaSharedmemCarveoutMaxShared);                                                            \
      if (dev >= 0 && dev < 64) done[dev] = true;                                                                      \
    }                                                                                                                  \
    classify_kernel<SLK_W, HITS, PACKED><<<grid, SLK_CLS_THREADS, SLK_SMEM_BYTES, a.stream>>>(SLK_CLS_ARGS(a, hb, hs, hc)); \
  } while (0)

void SLK_CAT(slk_launch_classify_w, SLK_W)(const slk_classify_args& a) {
  unsigned grid = (a.n_reads + SLK_CLS_THREADS - 1) / SLK_CLS_THREADS;
  if (a.hits) {
    if (a.packed) SLK_CLS_LAUNCH(true, true, a.hits_base, a.hits_shift_ptr, a.hits_cap);
    else SLK_CLS_LAUNCH(true, false, a.hits_base, a.hits_shift_ptr, a.hits_cap);
  } else {
    if (a.packed) SLK_CLS_LAUNCH(false, true, nullptr, nullptr, 0);
    else SLK_CLS_LAUNCH(false, false, nullptr, nullptr, 0);
  }
}
template <bool CANON>
static void launch_classify2(const slk_classify2_args& a, cudaStream_t stream) {
  static bool done[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !done[dev]) {
    cudaFuncSetAttribute(classify2_kernel<SLK_W, CANON>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SLK_G_SMEM_BYTES);
    cudaFuncSetAttribute(classify2_kernel<SLK_W, CANON>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (dev >= 0 && dev < 64) done[dev] = true;
  }
  const unsigned grid = (a.n_reads + SLK_G_THREADS - 1) / SLK_G_THREADS;
  classify2_kernel<SLK_W, CANON><<<grid, SLK_G_THREADS, SLK_G_SMEM_BYTES, stream>>>(a);
}
void SLK_CAT(slk_launch_classify2_w, SLK_W)(const slk_classify2_args& a, cudaStream_t stream) {
  if (a.sp.canonical) launch_classify2<true>(a, stream);
  else launch_classify2<false>(a, stream);
}
void SLK_CAT(slk_launch_emit_w, SLK_W)(const slk_emit_args& a) {
  emit_cells_kernel<SLK_W><<<(unsigned)((a.n_items + 127) / 128), 128, 0, a.stream>>>(
      a.sp, a.bases, a.frag_off, a.off_shift, a.frag_dense, a.item_prefix, a.n_frag, a.n_items, a.out, a.cap, a.cursor);
}
void SLK_CAT(slk_launch_spans_w, SLK_W)(const slk_spans_args& a) {
  const unsigned grid = (a.n_reads + 127) / 128;
  if (a.scratch) spans_strided_kernel<SLK_W><<<grid, 128, 0, a.stream>>>(a.sp, a.bases1, a.off1, a.bases2, a.off2, a.n_reads, a.stride,
                                                                          a.span_off, a.scratch, a.overflow);
  else if (a.spans) spans_kernel<SLK_W, true><<<grid, 128, 0, a.stream>>>(a.sp, a.bases1, a.off1, a.bases2, a.off2, a.n_reads, a.span_off, a.spans);
  else spans_kernel<SLK_W, false><<<grid, 128, 0, a.stream>>>(a.sp, a.bases1, a.off1, a.bases2, a.off2, a.n_reads, a.span_off, a.spans);
}
void SLK_CAT(slk_launch_bracken_scan_w, SLK_W)(const slk_bracken_scan_args& a) {
  const unsigned grid = (a.n_frag + 31) / 32;
  if (a.hits) bracken_scan_kernel<SLK_W, true><<<grid, 32, 0, a.stream>>>(a.sp, a.bases, a.frag_off, a.n_frag, a.hit_off, a.hits);
  else bracken_scan_kernel<SLK_W, false><<<grid, 32, 0, a.stream>>>(a.sp, a.bases, a.frag_off, a.n_frag, a.hit_off, a.hits);
}
