// cuda_emu.h -- TEST INFRASTRUCTURE ONLY.  A minimal CUDA execution-model emulator: every CUDA
// thread of a block is an OS thread, __syncthreads()/__syncwarp() are barriers and warp shuffles
// go through a per-warp exchange buffer.  It lets `pytest -m "not gpu"` run the UNMODIFIED kernel
// bodies (k1_symbols.cuh, k3_imdct.cuh, k3_streams.cuh) and the host engine on a machine without a GPU.  It is
// never compiled into libvpz.so (the product has no CPU path); see tests/emu/README.md.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#define VPZ_EMU 1

struct float2 {
  float x, y;
};
struct alignas(8) uint2 {
  uint32_t x, y;
};
struct alignas(16) uint4 {
  uint32_t x, y, z, w;
};
struct alignas(16) float4 {
  float x, y, z, w;
};
struct emu_dim3 {
  unsigned x = 1, y = 1, z = 1;
};

namespace emu {

class Barrier {
 public:
  explicit Barrier(int n) : n_(n) {}
  void wait() {
    std::unique_lock<std::mutex> lk(m_);
    int gen = gen_;
    if (++count_ == n_) {
      count_ = 0;
      gen_++;
      cv_.notify_all();
    } else {
      cv_.wait(lk, [&] { return gen != gen_; });
    }
  }

 private:
  std::mutex m_;
  std::condition_variable cv_;
  int n_, count_ = 0, gen_ = 0;
};

struct WarpCtx {
  Barrier bar{32};
  uint32_t slot[32];
};
struct BlockCtx {
  Barrier* bar;
  std::vector<WarpCtx*> warps;
  void* smem;
  std::mutex named_m;
  Barrier* named[16] = {nullptr};   // bar.sync id, count (created on first use)
};

extern thread_local BlockCtx* t_block;
extern thread_local int t_lane, t_warp;

// Runs `blocks` blocks of `threads` threads; body is the kernel body (reads threadIdx etc.).
void launch(unsigned blocks, unsigned threads, size_t smem_bytes, const std::function<void()>& body);

}  // namespace emu

extern thread_local emu_dim3 threadIdx, blockIdx, blockDim, gridDim;

inline void __syncthreads() { emu::t_block->bar->wait(); }
inline void emu_named_barrier(int id, int count) {
  emu::BlockCtx* b = emu::t_block;
  emu::Barrier* bar;
  {
    std::lock_guard<std::mutex> lk(b->named_m);
    if (!b->named[id]) b->named[id] = new emu::Barrier(count);
    bar = b->named[id];
  }
  bar->wait();
}
inline void __syncwarp(unsigned = 0xffffffffu) { emu::t_block->warps[emu::t_warp]->bar.wait(); }

template <typename T>
inline T emu_shfl_idx(T v, int src) {
  static_assert(sizeof(T) == 4, "32-bit shuffles only");
  emu::WarpCtx* w = emu::t_block->warps[emu::t_warp];
  memcpy(&w->slot[emu::t_lane], &v, 4);
  w->bar.wait();
  T r;
  memcpy(&r, &w->slot[src & 31], 4);
  w->bar.wait();
  return r;
}
template <typename T>
inline T __shfl_sync(unsigned, T v, int src) { return emu_shfl_idx(v, src); }
template <typename T>
inline T __shfl_xor_sync(unsigned, T v, int x) { return emu_shfl_idx(v, emu::t_lane ^ x); }
template <typename T>
inline T __shfl_up_sync(unsigned, T v, int d) { return emu_shfl_idx(v, emu::t_lane >= d ? emu::t_lane - d : emu::t_lane); }

inline int __any_sync(unsigned, int pred) {
  emu::WarpCtx* w = emu::t_block->warps[emu::t_warp];
  w->slot[emu::t_lane] = pred ? 1u : 0u;
  w->bar.wait();
  int r = 0;
  for (int i = 0; i < 32; i++) r |= (int)w->slot[i];
  w->bar.wait();
  return r;
}
inline uint32_t __ballot_sync(unsigned, int pred) {
  emu::WarpCtx* w = emu::t_block->warps[emu::t_warp];
  w->slot[emu::t_lane] = pred ? 1u : 0u;
  w->bar.wait();
  uint32_t r = 0;
  for (int i = 0; i < 32; i++) r |= w->slot[i] << i;
  w->bar.wait();
  return r;
}
inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t sh) {
  sh &= 31;
  uint64_t v = ((uint64_t)hi << 32) | lo;
  return (uint32_t)(v >> sh);
}
inline float __uint_as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
inline int __ffs(int x) { return __builtin_ffs(x); }
inline int __popc(uint32_t x) { return __builtin_popcount(x); }
inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
inline int __ffsll(long long x) { return __builtin_ffsll(x); }
inline uint32_t __brev(uint32_t n) {
  n = ((n & 0xAAAAAAAAu) >> 1) | ((n & 0x55555555u) << 1);
  n = ((n & 0xCCCCCCCCu) >> 2) | ((n & 0x33333333u) << 2);
  n = ((n & 0xF0F0F0F0u) >> 4) | ((n & 0x0F0F0F0Fu) << 4);
  n = ((n & 0xFF00FF00u) >> 8) | ((n & 0x00FF00FFu) << 8);
  return (n >> 16) | (n << 16);
}
// compile this file with -ffp-contract=off so a*b+c is never fused, like the __f*_rn intrinsics
inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }

inline uint32_t atomicAdd(uint32_t* p, uint32_t v) {
  return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST);
}
inline uint32_t atomicMin(uint32_t* p, uint32_t v) {
  uint32_t old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
  while (old > v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {
  }
  return old;
}
