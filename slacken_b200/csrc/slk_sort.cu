// slk_sort.cu -- K3a: radix sort of the (compressed minimizer << 16 | dense taxon) cells.
// Round-1 implementation: CUB's DeviceRadixSort (a library call, declared as such in DESIGN.md); the cells are
// sorted as plain 64-bit keys, so a (minimizer, taxon) pair costs 8 bytes per pass instead of 12.
#include "slk_sort.h"

#include <cub/device/device_radix_sort.cuh>

int slk_sort_u64(uint64_t* keys, uint64_t* tmp, uint64_t n, int begin_bit, int end_bit, cudaStream_t stream,
                 uint64_t** sorted) {
  cub::DoubleBuffer<uint64_t> db(keys, tmp);
  size_t bytes = 0;
  cudaError_t e = cub::DeviceRadixSort::SortKeys(nullptr, bytes, db, (int64_t)n, begin_bit, end_bit, stream);
  if (e != cudaSuccess) return (int)e;
  void* scratch = nullptr;
  e = cudaMalloc(&scratch, bytes ? bytes : 16);
  if (e != cudaSuccess) return (int)e;
  e = cub::DeviceRadixSort::SortKeys(scratch, bytes, db, (int64_t)n, begin_bit, end_bit, stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
  cudaFree(scratch);
  if (e != cudaSuccess) return (int)e;
  *sorted = db.Current();
  return 0;
}
