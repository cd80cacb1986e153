// tests/emu/simt/cuda_runtime.h -- TEST INFRASTRUCTURE: a minimal SIMT emulator that lets the warp-synchronous CUDA
// kernels of cvxpygen_b200/csrc (admm_kernel.cuh, matpar_kernel.cuh, grad_kernel.cuh) compile and run as plain host C++.
//
// It shadows <cuda_runtime.h> (put this directory first on the include path and define CPG_SIMT_HOST_EMU).  Every CUDA
// thread of a block is a ucontext fiber with its own stack; fibers run cooperatively and switch only at the warp / block
// synchronisation points the kernels use (__shfl*_sync, __any_sync, __syncwarp, __syncthreads, __syncthreads_or), so the
// lane-level data flow -- who reads whose register through which shuffle, which lane writes which shared-memory word
// between two barriers -- is executed exactly as written.  Arithmetic is IEEE double with std::fma, i.e. the results are
// those of the GPU up to the order of floating-point atomics.  Blocks run one after the other.
//
// Not emulated: real concurrency (races show up only as schedule dependence, see simt::Runtime::schedule), memory spaces (shared memory is one host buffer per
// block), TMA / mbarrier (admm_kernel.cuh replaces its four helpers by a memcpy under CPG_SIMT_HOST_EMU).
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>
#include <functional>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __maxnreg__(...)
#define __shared__ thread_local   /* one OS thread runs all fibers, so a thread_local object is shared by every CUDA thread:
                                     `extern __shared__ T smem[]` binds to a thread_local host array the driver defines, a
                                     block-scope `__shared__ int x;` becomes one static object per kernel */
#define __align__(n)
#define __constant__ static

struct dim3 { unsigned x = 1, y = 1, z = 1; dim3() {} dim3(unsigned a, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
struct ushort2 { unsigned short x, y; };
struct ushort4 { unsigned short x, y, z, w; };
inline ushort2 make_ushort2(unsigned short x, unsigned short y) { ushort2 r; r.x = x; r.y = y; return r; }
struct int4 { int x, y, z, w; };
struct uint4 { unsigned x, y, z, w; };
struct alignas(16) double2 { double x, y; };
inline double2 make_double2(double x, double y) { double2 r; r.x = x; r.y = y; return r; }
typedef void* cudaStream_t;

namespace simt {

constexpr int WARP = 32;
constexpr size_t STACK_BYTES = 1u << 20;

struct WarpState {
  uint64_t buf[WARP];
  int arrived = 0, gen = 0, live = 0;
};
struct BlockState {
  int arrived = 0, gen = 0, live = 0, or_acc = 0, or_result = 0;
};
struct Fiber {
  ucontext_t ctx;
  char* stack = nullptr;
  bool done = false;
  dim3 tid, bid, bdim, gdim;
  int lane = 0, warp = 0;
};
struct Runtime {
  ucontext_t main_ctx;
  std::vector<Fiber> fibers;
  std::vector<WarpState> warps;
  BlockState block;
  BlockState named[16];      // bar.sync id, count (named barriers of a subset of the block's warps)
  Fiber* cur = nullptr;
  std::function<void()> body;
  long long n_exchange = 0, n_syncwarp = 0, n_syncthreads = 0;   // per-lane counts of synchronisation points (profiling aid)
  int schedule = 0;          // 0: threads resumed in ascending order, 1: descending, >= 2: a fixed pseudo-random permutation per pass
};
inline Runtime& rt() { static Runtime r; return r; }
inline Fiber* cur() { return rt().cur; }
inline void yield() { Fiber* f = rt().cur; swapcontext(&f->ctx, &rt().main_ctx); }

inline void warp_barrier() {
  WarpState& w = rt().warps[cur()->warp];
  const int gen = w.gen;
  if (++w.arrived == w.live) { w.arrived = 0; ++w.gen; }
  else while (w.gen == gen) yield();
}
inline void block_barrier() {
  BlockState& b = rt().block;
  const int gen = b.gen;
  if (++b.arrived == b.live) { b.arrived = 0; b.or_result = b.or_acc; b.or_acc = 0; ++b.gen; }
  else while (b.gen == gen) yield();
}
// bar.sync id, count: the first `count` arrivals release each other (no participant leaves the kernel before the others arrive)
inline void named_barrier(int id, int count) {
  BlockState& b = rt().named[id & 15];
  const int gen = b.gen;
  if (++b.arrived == count) { b.arrived = 0; ++b.gen; }
  else while (b.gen == gen) yield();
}
template <class T> inline uint64_t to_bits(T v) { uint64_t u = 0; memcpy(&u, &v, sizeof(T)); return u; }
template <class T> inline T from_bits(uint64_t u) { T v; memcpy(&v, &u, sizeof(T)); return v; }
template <class T> inline T exchange(T v, int src_lane) {
  static_assert(sizeof(T) <= 8, "shuffle of at most 64 bits");
  WarpState& w = rt().warps[cur()->warp];
  ++rt().n_exchange;
  w.buf[cur()->lane] = to_bits(v);
  warp_barrier();
  const T r = from_bits<T>(w.buf[src_lane & (WARP - 1)]);
  warp_barrier();
  return r;
}
inline void fiber_entry() {
  Runtime& r = rt();
  r.body();
  Fiber* f = r.cur;
  f->done = true;
  --r.warps[f->warp].live;           // a finished thread no longer takes part in barriers
  --r.block.live;
  swapcontext(&f->ctx, &r.main_ctx);
}
// run `body` as a grid of `grid` blocks x `threads` threads; blocks sequentially, threads of a block cooperatively
inline void launch(int grid, int threads, std::function<void()> body) {
  Runtime& r = rt();
  r.body = body;
  for (int b = 0; b < grid; ++b) {
    const int nw = (threads + WARP - 1) / WARP;
    r.fibers.assign(threads, Fiber());
    r.warps.assign(nw, WarpState());
    r.block = BlockState();
    for (auto& nb : r.named) nb = BlockState();
    r.block.live = threads;
    for (int t = 0; t < threads; ++t) {
      Fiber& f = r.fibers[t];
      f.stack = static_cast<char*>(malloc(STACK_BYTES));
      f.tid = dim3(t); f.bid = dim3(b); f.bdim = dim3(threads); f.gdim = dim3(grid);
      f.lane = t % WARP; f.warp = t / WARP;
      ++r.warps[f.warp].live;
      getcontext(&f.ctx);
      f.ctx.uc_stack.ss_sp = f.stack; f.ctx.uc_stack.ss_size = STACK_BYTES; f.ctx.uc_link = &r.main_ctx;
      makecontext(&f.ctx, fiber_entry, 0);
    }
    // The order in which runnable threads are resumed is the emulator's only freedom.  A kernel whose result depends on it
    // has an unsynchronised dependency between lanes (a missing __syncwarp / __syncthreads): running the same launch under
    // several schedules and comparing bit for bit is the race check of tests/test_simt_emulation.py.
    std::vector<int> order(threads);
    uint64_t lcg = 0x9e3779b97f4a7c15ull * (uint64_t)(r.schedule + 1);
    for (int alive = threads; alive > 0;) {
      alive = 0;
      for (int t = 0; t < threads; ++t) order[t] = r.schedule == 1 ? threads - 1 - t : t;
      if (r.schedule >= 2)
        for (int t = threads - 1; t > 0; --t) {
          lcg = lcg * 6364136223846793005ull + 1442695040888963407ull;
          const int k = (int)((lcg >> 33) % (uint64_t)(t + 1));
          const int tmp = order[t]; order[t] = order[k]; order[k] = tmp;
        }
      for (int ti = 0; ti < threads; ++ti) {
        const int t = order[ti];
        Fiber& f = r.fibers[t];
        if (f.done) continue;
        r.cur = &f;
        swapcontext(&r.main_ctx, &f.ctx);
        if (!f.done) ++alive;
      }
    }
    for (auto& f : r.fibers) free(f.stack);
    r.cur = nullptr;
  }
}

}  // namespace simt

#define threadIdx (simt::cur()->tid)
#define blockIdx (simt::cur()->bid)
#define blockDim (simt::cur()->bdim)
#define gridDim (simt::cur()->gdim)

template <class T> inline T __shfl_sync(unsigned, T v, int src) { return simt::exchange(v, src); }
template <class T> inline T __shfl_xor_sync(unsigned, T v, int mask) { return simt::exchange(v, simt::cur()->lane ^ mask); }
template <class T> inline T __shfl_down_sync(unsigned, T v, int d) { const int s = simt::cur()->lane + d; return simt::exchange(v, s < simt::WARP ? s : simt::cur()->lane); }
inline int __any_sync(unsigned, int pred) {
  int acc = 0;
  for (int o = 16; o; o >>= 1) { pred |= simt::exchange(pred, simt::cur()->lane ^ o); }
  acc = pred;
  return acc != 0;
}
inline int __all_sync(unsigned m, int pred) { return !__any_sync(m, !pred); }
template <class T> inline T __reduce_max_sync(unsigned, T v) {
  for (int o = 16; o; o >>= 1) { const T w = simt::exchange(v, simt::cur()->lane ^ o); v = w > v ? w : v; }
  return v;
}
template <class T> inline T __reduce_min_sync(unsigned, T v) {
  for (int o = 16; o; o >>= 1) { const T w = simt::exchange(v, simt::cur()->lane ^ o); v = w < v ? w : v; }
  return v;
}
template <class T> inline T __reduce_add_sync(unsigned, T v) {
  for (int o = 16; o; o >>= 1) v += simt::exchange(v, simt::cur()->lane ^ o);
  return v;
}
inline unsigned __ballot_sync(unsigned, int pred) {
  unsigned v = pred ? (1u << simt::cur()->lane) : 0u;
  for (int o = 16; o; o >>= 1) v |= simt::exchange(v, simt::cur()->lane ^ o);
  return v;
}
inline void __syncwarp(unsigned = 0xffffffffu) { ++simt::rt().n_syncwarp; simt::warp_barrier(); }
inline void __syncthreads() { ++simt::rt().n_syncthreads; simt::block_barrier(); }
inline int __syncthreads_or(int pred) { simt::rt().block.or_acc |= (pred != 0); simt::block_barrier(); const int r = simt::rt().block.or_result; simt::block_barrier(); return r; }

template <class T> inline T atomicAdd(T* p, T v) { const T old = *p; *p = old + v; return old; }
template <class T> inline T atomicOr(T* p, T v) { const T old = *p; *p = old | v; return old; }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline unsigned __byte_perm(unsigned x, unsigned y, unsigned s) {      // PRMT, default mode: result byte i = byte (s >> 4i) & 7 of {x, y}
  const unsigned long long pool = ((unsigned long long)y << 32) | x;
  unsigned r = 0;
  for (int i = 0; i < 4; ++i) r |= (unsigned)((pool >> (8 * ((s >> (4 * i)) & 7))) & 0xff) << (8 * i);
  return r;
}
template <class T> inline T __ldg(const T* p) { return *p; }
inline double __longlong_as_double(long long v) { double d; memcpy(&d, &v, 8); return d; }
inline long long __double_as_longlong(double d) { long long v; memcpy(&v, &d, 8); return v; }
inline size_t __cvta_generic_to_shared(const void* p) { return reinterpret_cast<size_t>(p); }
template <class T> inline T min(T a, T b) { return a < b ? a : b; }
template <class T> inline T max(T a, T b) { return a > b ? a : b; }
