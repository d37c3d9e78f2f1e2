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

void SLK_CAT(slk_launch_classify_w, SLK_W)(const slk_classify_args& a) {
  unsigned grid = (a.n_reads + 127) / 128;
  if (a.hits) {
    if (a.packed)
      classify_kernel<SLK_W, true, true><<<grid, 128, 0, a.stream>>>(SLK_CLS_ARGS(a, a.hits_base, a.hits_shift_ptr, a.hits_cap));
    else
      classify_kernel<SLK_W, true, false><<<grid, 128, 0, a.stream>>>(SLK_CLS_ARGS(a, a.hits_base, a.hits_shift_ptr, a.hits_cap));
  } else {
    if (a.packed)
      classify_kernel<SLK_W, false, true><<<grid, 128, 0, a.stream>>>(SLK_CLS_ARGS(a, nullptr, nullptr, 0));
    else
      classify_kernel<SLK_W, false, false><<<grid, 128, 0, a.stream>>>(SLK_CLS_ARGS(a, nullptr, nullptr, 0));
  }
}
void SLK_CAT(slk_launch_emit_w, SLK_W)(const slk_emit_args& a) {
  emit_cells_kernel<SLK_W><<<(unsigned)((a.n_items + 127) / 128), 128, 0, a.stream>>>(
      a.sp, a.bases, a.frag_off, a.off_shift, a.frag_dense, a.item_prefix, a.n_frag, a.n_items, a.out, a.cap, a.cursor);
}
