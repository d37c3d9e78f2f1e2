// How long do device allocations take on this box? (the distributed build allocates 23-46 GB buffers inside its timed
// region).  nvcc -O2 -o alloc_cost alloc_cost.cu && ./alloc_cost
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main() {
  cudaFree(0);
  const size_t sizes[] = {4096, 1ull << 20, 64ull << 20, 1ull << 30, 8ull << 30, 23ull << 30, 46ull << 30};
  for (size_t sz : sizes) {
    printf("%8.3f GB:", sz / 1073741824.0);
    for (int rep = 0; rep < 4; rep++) {
      void* p = nullptr;
      double t0 = now(); cudaError_t e = cudaMalloc(&p, sz); double t1 = now();
      if (e != cudaSuccess) { printf(" failed"); continue; }
      cudaMemset(p, 0, sz); cudaDeviceSynchronize();
      double t2 = now(); cudaFree(p); double t3 = now();
      printf("  malloc %.2f free %.2f ms |", (t1 - t0) * 1e3, (t3 - t2) * 1e3);
    }
    printf("\n");
  }
  // two live large buffers, then a third (the build's situation)
  void *a, *b, *c;
  cudaMalloc(&a, 23ull << 30); cudaMalloc(&b, 23ull << 30);
  double t0 = now(); cudaMalloc(&c, 46ull << 30); double t1 = now(); cudaFree(a); double t2 = now(); cudaFree(b); double t3 = now();
  printf("with 46 GB live: malloc 46 GB %.2f ms, free 23 GB %.2f ms, free 23 GB %.2f ms\n", (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3);
  cudaFree(c);
  return 0;
}
