// emu2.cpp -- TEST-ONLY host emulation of the warp-cooperative classify kernel body (slacken_b200/csrc/slk_group.h).
// simt.h runs the 32 lanes of every warp as fibres, so the very source the GPU runs is checked against the oracle on
// machines without a GPU (pytest -m "not gpu"). Not a fallback: nothing in slacken_b200/ loads it.
#include "simt.h"

#include <vector>

// the smallest legal buffer, a three-step table and a two-pair fast histogram, so that the tests close the buffer and
// take the slow paths all the time
#define SLK_G_CAP 608
#define SLK_G_STEPS 3
#define SLK_G_HIST 2
#include "../../include/slacken_gpu.h"
#include "../../slacken_b200/csrc/slk_group.h"

#define EMU_API extern "C" __attribute__((visibility("default")))

template <int W>
static void run_w(const slk_classify2_args& a, uint32_t threads) {
  const uint32_t warps = threads / 32;
  std::vector<uint8_t> smem((size_t)warps * SLK_G_WARP_BYTES + 64);
  uint8_t* base = reinterpret_cast<uint8_t*>(((uintptr_t)smem.data() + 15) & ~(uintptr_t)15);
  const uint32_t blocks = (a.n_reads + threads - 1) / threads;
  simt::bdim().x = threads;
  for (uint32_t b = 0; b < blocks; b++) {
    simt::bid().x = b;
    for (uint32_t w = 0; w < warps; w++)
      simt::runner::run_warp(w, base, [&]() {
        if (a.sp.canonical) slk_classify2_thread<W, true>(a, base);
        else slk_classify2_thread<W, false>(a, base);
      });
  }
}

EMU_API int emu2_classify(const slk_scan_params* sp, uint64_t* cells, uint64_t n_buckets, const uint16_t* parent, const uint8_t* depth,
                          const int32_t* raw, uint32_t n_dense, uint32_t root, const uint64_t* codes1, const uint32_t* mask1,
                          const uint64_t* boff1, const uint32_t* len1, const uint64_t* codes2, const uint32_t* mask2,
                          const uint64_t* boff2, const uint32_t* len2, uint32_t n, const double* confidence, uint32_t n_conf, int min_hit_groups,
                          int want_hits, int32_t* taxon_out, uint8_t* flags_out, slk_read_detail* detail_out, slk_hit* hits_out,
                          uint64_t hits_cap, uint64_t* hits_used, uint64_t* stats, uint64_t* counts) {
  slk_classify2_args a;
  a.sp = *sp;
  a.tb = slk_table_view{cells, n_buckets, 1, 0};
  a.tx = slk_tax_view{parent, depth, raw, n_dense, root};
  a.in1 = slk_group_in{codes1, mask1, boff1, len1, 0};
  a.in2 = slk_group_in{codes2, mask2, boff2, len2, 0};
  a.paired = codes2 != nullptr; a.n_reads = n; a.min_hit_groups = min_hit_groups; a.hits = want_hits != 0;
  a.mt.n = n_conf; a.mt.stride = n; a.mt.taxon_out = taxon_out; a.mt.flags_out = flags_out;   // [n_conf][n]
  for (uint32_t t = 0; t < n_conf; t++) a.mt.confidence[t] = confidence[t];
  a.taxon_out = taxon_out; a.flags_out = flags_out; a.detail_out = detail_out;
  unsigned long long cursor = 0;
  uint32_t err = 0;
  a.hits_base = hits_out; a.hits_shift_ptr = nullptr; a.hits_cap = hits_cap; a.hits_cursor = &cursor; a.hits_over = nullptr;
  a.counts = reinterpret_cast<unsigned long long*>(counts); a.error_flag = &err; a.stats = reinterpret_cast<unsigned long long*>(stats);
  switch (sp->w) {
    case 1: run_w<1>(a, 64); break; case 2: run_w<2>(a, 64); break; case 3: run_w<3>(a, 64); break; case 4: run_w<4>(a, 64); break;
    case 5: run_w<5>(a, 64); break; case 6: run_w<6>(a, 64); break; case 7: run_w<7>(a, 64); break; case 8: run_w<8>(a, 64); break;
    default: return -1;
  }
  if (hits_used) *hits_used = cursor;
  return err ? -5 : 0;
}
