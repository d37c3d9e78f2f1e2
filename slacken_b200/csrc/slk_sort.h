// slk_sort.h -- device radix sort of 64-bit cells (K3a of the library build).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

// Sorts n keys on bits [begin_bit, end_bit). `keys` and `tmp` are both n-element device buffers and are both
// clobbered; *sorted points at whichever of them holds the result. Returns 0 or a non-zero CUDA error code.
int slk_sort_u64(uint64_t* keys, uint64_t* tmp, uint64_t n, int begin_bit, int end_bit, cudaStream_t stream,
                 uint64_t** sorted);
// The same for build cells (compressed key << 16 | taxon), ordered by slk_key_mix(key): four 8-bit passes. Cells of
// the same key end up adjacent, and the order is the order of the table's lines (slk_bucket_of).
int slk_sort_cells_by_line(uint64_t* cells, uint64_t* tmp, uint64_t n, cudaStream_t stream, uint64_t** sorted);

// In-place exclusive prefix sum of n 64-bit counters on the device (allocates its own scratch; asynchronous on `stream`
// apart from that allocation). Returns 0 or a non-zero CUDA error code.
int slk_exclusive_scan_u64(uint64_t* d, uint64_t n, cudaStream_t stream);
// ... with caller-owned scratch of slk_scan_scratch_words(n) words: no allocation, no synchronisation
uint64_t slk_scan_scratch_words(uint64_t n);
int slk_exclusive_scan_u64_async(uint64_t* d, uint64_t n, uint64_t* scratch, cudaStream_t stream);
