// slk_host.h -- host-side handle types and helpers shared by the translation units of libslacken_gpu.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <unordered_map>
#include <vector>

#include "../../include/slacken_gpu.h"
#include "slk_core.h"

struct slk_ctx {
  int device;
  int sm_count;
  cudaStream_t stream;
  cudaStream_t scan_stream;   // slk_scan_spans_dev / slk_emit_spans_dev (synchronous calls) run here, so that a caller may
                              // scan the next batch while the mailbox kernels of the current one are still in flight
  // split path, one-pass scan: the span words of the last count-only slk_scan_spans_dev call, `stride` slots per
  // fragment, waiting for slk_emit_spans_dev to compact them (grow-only scratch owned by the context)
  uint64_t* span_scratch = nullptr;
  size_t span_scratch_words = 0;
  uint32_t* d_maxlen = nullptr;      // [3]: longest mate 1, longest mate 2, overflow flag
  struct {
    const void* bases1 = nullptr; const void* off1 = nullptr; const void* bases2 = nullptr; const void* off2 = nullptr;
    const void* span_off = nullptr; uint32_t n_reads = 0; uint32_t stride = 0; bool valid = false;
  } pending_spans;
};
struct slk_tax {
  slk_ctx* ctx;
  std::vector<int32_t> parents;  // raw-indexed
};
struct dense_tax {
  std::vector<int32_t> raw;       // dense -> raw, raw[0] = 0
  std::vector<uint16_t> parent;   // dense parent
  std::vector<uint8_t> depth;
  std::unordered_map<int32_t, uint32_t> to_dense;
  uint32_t root = 0;
  uint16_t* d_parent = nullptr;
  uint8_t* d_depth = nullptr;
  int32_t* d_raw = nullptr;
  slk_tax_view view() const {
    slk_tax_view v;
    v.parent = d_parent; v.depth = d_depth; v.raw = d_raw; v.n = (uint32_t)raw.size(); v.root = root;
    return v;
  }
};
struct slk_index {
  slk_ctx* ctx;
  slk_tax* tax;
  slk_params params;
  slk_scan_params sp;
  dense_tax dt;
  slk_table_view table{nullptr, 0, 0, 0};
  uint64_t n_records = 0;
};

// sets the thread-local error string of slk_last_error() and returns `code`
int slk_fail(int code, const char* fmt, ...);
#define SLK_CU(call)                                                                                       \
  do {                                                                                                     \
    cudaError_t e_ = (call);                                                                               \
    if (e_ != cudaSuccess)                                                                                 \
      return slk_fail(e_ == cudaErrorMemoryAllocation ? SLK_E_NOMEM : SLK_E_CUDA, "%s failed: %s (%s:%d)", #call, \
                      cudaGetErrorString(e_), __FILE__, __LINE__);                                         \
  } while (0)
int slk_make_scan_params_checked(const slk_params* p, slk_scan_params* sp);
// dense, ancestor-closed numbering of a set of taxa (dense 0 = NONE, ROOT always present)
void slk_dense_init(dense_tax& dt);
int slk_dense_add(dense_tax& dt, const slk_tax* tax, int32_t raw, uint32_t* out);
int slk_dense_upload(dense_tax& dt);
void slk_dense_free(dense_tax& dt);
