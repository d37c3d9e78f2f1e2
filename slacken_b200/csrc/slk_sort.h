// slk_sort.h -- device radix sort of 64-bit cells (K3a of the library build).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

// Sorts n keys on bits [begin_bit, end_bit). `keys` and `tmp` are both n-element device buffers and are both
// clobbered; *sorted points at whichever of them holds the result. Returns 0 or a non-zero CUDA error code.
int slk_sort_u64(uint64_t* keys, uint64_t* tmp, uint64_t n, int begin_bit, int end_bit, cudaStream_t stream,
                 uint64_t** sorted);
