// simt.h -- TEST-ONLY: runs warp-cooperative CUDA device code on the CPU, one warp at a time, so that the kernels which
// use warp collectives (shuffles, ballots, reductions) can be checked against the oracle without a GPU.
//
// Every lane of a warp is a fibre (ucontext). A collective deposits the lane's operand, yields, and computes its result
// from the 32 deposited operands once every lane has arrived; the scheduler sweeps the lanes round-robin. Operand slots
// are double-buffered by collective parity: lane 0 may already deposit for collective c+1 while lane 31 still reads c.
// Only full-mask collectives are supported (the kernels here never use partial masks), and every lane of a warp must
// run the same sequence of collectives, which is exactly the convergence rule the device code has to obey as well.
// Never part of the product: nothing under slacken_b200/ includes this file.
#pragma once
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#include <functional>

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
#define __restrict__
#define __launch_bounds__(...)
#define __grid_constant__

namespace simt {
struct dim3_ { uint32_t x, y, z; };
struct warp_state {
  uint64_t slot[2][32];
  uint32_t gen;        // collectives completed by the lane that is furthest ahead
  uint32_t lane_gen[32];
  bool done[32];
  ucontext_t sched, lane_ctx[32];
  uint32_t cur;
  uint8_t* smem;
};
inline warp_state*& W() { static warp_state* w = nullptr; return w; }
inline dim3_& tid() { static dim3_ t{0, 0, 0}; return t; }
inline dim3_& bid() { static dim3_ b{0, 0, 0}; return b; }
inline dim3_& bdim() { static dim3_ d{32, 1, 1}; return d; }

// deposit + wait for everybody; returns the buffer holding all 32 operands of this collective
inline const uint64_t* collect(uint64_t v) {
  warp_state* w = W();
  const uint32_t lane = w->cur, g = w->lane_gen[lane]++;
  w->slot[g & 1][lane] = v;
  swapcontext(&w->lane_ctx[lane], &w->sched);   // resumed after a full sweep: every lane has deposited generation g
  return w->slot[g & 1];
}

// runs body(lane) for the 32 lanes of one warp; warp_in_block sets threadIdx.x = warp_in_block * 32 + lane
struct runner {
  static void tramp(unsigned lo, unsigned hi) {
    auto* f = reinterpret_cast<std::function<void()>*>(((uintptr_t)hi << 32) | lo);
    (*f)();
    warp_state* w = W();
    w->done[w->cur] = true;
    swapcontext(&w->lane_ctx[w->cur], &w->sched);
  }
  static void run_warp(uint32_t warp_in_block, uint8_t* smem, const std::function<void()>& body) {
    static const size_t STACK = 1 << 20;
    static uint8_t* stacks = nullptr;
    if (!stacks) stacks = (uint8_t*)malloc(32 * STACK);
    warp_state* w = new warp_state;
    memset(w->slot, 0, sizeof(w->slot));
    w->gen = 0; w->smem = smem;
    W() = w;
    std::function<void()> fn = body;
    for (uint32_t l = 0; l < 32; l++) {
      w->lane_gen[l] = 0; w->done[l] = false;
      getcontext(&w->lane_ctx[l]);
      w->lane_ctx[l].uc_stack.ss_sp = stacks + l * STACK;
      w->lane_ctx[l].uc_stack.ss_size = STACK;
      w->lane_ctx[l].uc_link = nullptr;
      uintptr_t p = (uintptr_t)&fn;
      makecontext(&w->lane_ctx[l], (void (*)())tramp, 2, (unsigned)(p & 0xffffffffu), (unsigned)(p >> 32));
    }
    bool any = true;
    while (any) {
      any = false;
      for (uint32_t l = 0; l < 32; l++) {
        if (w->done[l]) continue;
        any = true;
        w->cur = l;
        tid().x = warp_in_block * 32 + l;
        swapcontext(&w->sched, &w->lane_ctx[l]);
      }
      // lockstep check: all live lanes must have reached the same collective
      uint32_t g0 = 0; bool have = false;
      for (uint32_t l = 0; l < 32; l++)
        if (!w->done[l]) { if (!have) { g0 = w->lane_gen[l]; have = true; } else if (w->lane_gen[l] != g0) abort(); }
    }
    delete w;
    W() = nullptr;
  }
};
}  // namespace simt

#define threadIdx (simt::tid())
#define blockIdx (simt::bid())
#define blockDim (simt::bdim())

inline uint32_t __shfl_sync(uint32_t, uint32_t v, int src) { return (uint32_t)simt::collect(v)[src & 31]; }
inline int __shfl_sync(uint32_t, int v, int src) { return (int)(uint32_t)simt::collect((uint32_t)v)[src & 31]; }
inline uint64_t __shfl_sync(uint32_t, uint64_t v, int src) { return simt::collect(v)[src & 31]; }
inline unsigned long long __shfl_sync(uint32_t, unsigned long long v, int src) { return simt::collect(v)[src & 31]; }
inline uint32_t __shfl_up_sync(uint32_t, uint32_t v, unsigned d) {
  const uint32_t lane = simt::W()->cur;
  const uint64_t* s = simt::collect(v);
  return lane >= d ? (uint32_t)s[lane - d] : v;
}
inline uint64_t __shfl_up_sync(uint32_t, uint64_t v, unsigned d) {
  const uint32_t lane = simt::W()->cur;
  const uint64_t* s = simt::collect(v);
  return lane >= d ? s[lane - d] : v;
}
inline uint32_t __shfl_down_sync(uint32_t, uint32_t v, unsigned d) {
  const uint32_t lane = simt::W()->cur;
  const uint64_t* s = simt::collect(v);
  return lane + d < 32 ? (uint32_t)s[lane + d] : v;
}
inline uint32_t __ballot_sync(uint32_t, int p) {
  const uint64_t* s = simt::collect(p ? 1 : 0);
  uint32_t m = 0;
  for (int i = 0; i < 32; i++) m |= (uint32_t)(s[i] & 1) << i;
  return m;
}
inline int __any_sync(uint32_t m, int p) { return __ballot_sync(m, p) != 0; }
inline int __all_sync(uint32_t m, int p) { return __ballot_sync(m, p) == 0xffffffffu; }
inline uint32_t __reduce_add_sync(uint32_t, uint32_t v) {
  const uint64_t* s = simt::collect(v);
  uint32_t r = 0;
  for (int i = 0; i < 32; i++) r += (uint32_t)s[i];
  return r;
}
inline uint32_t __reduce_max_sync(uint32_t, uint32_t v) {
  const uint64_t* s = simt::collect(v);
  uint32_t r = 0;
  for (int i = 0; i < 32; i++) r = (uint32_t)s[i] > r ? (uint32_t)s[i] : r;
  return r;
}
inline uint32_t __reduce_or_sync(uint32_t, uint32_t v) {
  const uint64_t* s = simt::collect(v);
  uint32_t r = 0;
  for (int i = 0; i < 32; i++) r |= (uint32_t)s[i];
  return r;
}
inline uint32_t __match_any_sync(uint32_t, uint32_t v) {
  const uint64_t* s = simt::collect(v);
  uint32_t m = 0;
  for (int i = 0; i < 32; i++) m |= (uint32_t)((uint32_t)s[i] == v) << i;
  return m;
}
inline void __syncwarp(uint32_t = 0xffffffffu) { simt::collect(0); }
inline int __popc(uint32_t x) { return __builtin_popcount(x); }
inline int __ffs(int x) { return __builtin_ffs(x); }
inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
inline uint32_t __brev(uint32_t x) {
  x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
  x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
  x = ((x >> 4) & 0x0f0f0f0fu) | ((x & 0x0f0f0f0fu) << 4);
  return __builtin_bswap32(x);
}
inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t sh) {
  sh &= 31;
  return sh ? (lo >> sh) | (hi << (32 - sh)) : lo;
}
inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
template <class T> inline T __ldg(const T* p) { return *p; }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { unsigned long long o = *p; *p = o + v; return o; }
inline uint32_t atomicAdd(uint32_t* p, uint32_t v) { uint32_t o = *p; *p = o + v; return o; }
inline uint32_t atomicExch(uint32_t* p, uint32_t v) { uint32_t o = *p; *p = v; return o; }
inline double __longlong_as_double(long long x) { double d; memcpy(&d, &x, 8); return d; }
