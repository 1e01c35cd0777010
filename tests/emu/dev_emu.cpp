// dev_emu.cpp -- TEST INFRASTRUCTURE ONLY: devapi.h on top of cuda_emu.h (see cuda_emu.h).
#include "cuda_emu.h"

#include <stdio.h>
#include <stdlib.h>

#include <chrono>

#include "../../include/vpz.h"
#include "../../vorbispizza_b200/csrc/devapi.h"
#include "../../vorbispizza_b200/csrc/k0_pages.cuh"
#include "../../vorbispizza_b200/csrc/k4_deliver.cuh"
#include "../../vorbispizza_b200/csrc/k1_symbols.cuh"
#include "../../vorbispizza_b200/csrc/k3_streams.cuh"

thread_local emu_dim3 threadIdx, blockIdx, blockDim, gridDim;
namespace emu {
thread_local BlockCtx* t_block;
thread_local int t_lane, t_warp;

void launch(unsigned blocks, unsigned threads, size_t smem_bytes, const std::function<void()>& body) {
  for (unsigned b = 0; b < blocks; b++) {
    BlockCtx ctx;
    Barrier bar((int)threads);
    ctx.bar = &bar;
    unsigned nw = (threads + 31) / 32;
    for (unsigned w = 0; w < nw; w++) ctx.warps.push_back(new WarpCtx);
    // shared memory with a canary behind it: a kernel writing past its dynamic shared memory (an "illegal
    // memory access" on the device) fails here instead of corrupting the heap silently
    const size_t guard = 4096;
    ctx.smem = calloc(1, smem_bytes + guard);
    memset(static_cast<char*>(ctx.smem) + smem_bytes, 0xa5, guard);
    std::vector<std::thread> ts;
    for (unsigned t = 0; t < threads; t++) {
      ts.emplace_back([&, t] {
        t_block = &ctx;
        t_lane = (int)(t & 31);
        t_warp = (int)(t >> 5);
        threadIdx.x = t;
        blockIdx.x = b;
        blockDim.x = threads;
        gridDim.x = blocks;
        body();
      });
    }
    for (auto& t : ts) t.join();
    for (size_t i = 0; i < guard; i++)
      if (static_cast<unsigned char*>(ctx.smem)[smem_bytes + i] != 0xa5) {
        fprintf(stderr, "emu: block %u wrote %zu bytes past its %zu bytes of shared memory\n", b, i + 1, smem_bytes);
        abort();
      }
    for (auto* w : ctx.warps) delete w;
    for (auto* nb : ctx.named) delete nb;
    free(ctx.smem);
  }
}
}  // namespace emu

namespace vpz {
namespace dev {

struct Stream {
  int dummy;
};
struct Event {
  double t = 0;
};

int init(int device, int* resolved, std::string&) {
  if (resolved) *resolved = device < 0 ? 0 : device;
  return VPZ_OK;
}
void make_current(int) {}
int device_count() { return 1; }
int sm_count() { return 1; }
size_t max_smem_per_block() { return 227 * 1024; }
void* alloc(size_t bytes, std::string&) {
  // device memory is NOT zeroed by cudaMalloc: poison it (0xff bytes = NaN floats, huge indices) so that a
  // kernel reading something no kernel wrote shows up in the CPU suite
  void* p = malloc(bytes + 64);
  if (p) memset(p, 0xff, bytes + 64);
  return p;
}
void free(void* p) { ::free(p); }
void* host_alloc(size_t bytes) { return calloc(1, bytes + 64); }
void host_free(void* p) { ::free(p); }
Stream* stream_create() { return new Stream; }
void stream_destroy(Stream* s) { delete s; }
int stream_sync(Stream*, std::string&) { return VPZ_OK; }
Event* event_create() { return new Event; }
void event_destroy(Event* e) { delete e; }
void event_record(Event* e, Stream*) {
  e->t = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
int event_sync(Event*, std::string&) { return VPZ_OK; }
void stream_wait_event(Stream*, Event*) {}
float event_elapsed_ms(Event* a, Event* b) { return (float)(b->t - a->t); }
unsigned long long transfer_bytes(int) { return 0; }
int h2d(void* d, const void* s, size_t n, Stream*, std::string&) {
  memcpy(d, s, n);
  return VPZ_OK;
}
int d2h(void* d, const void* s, size_t n, Stream*, std::string&) {
  memcpy(d, s, n);
  return VPZ_OK;
}
int d2d(void* d, const void* s, size_t n, Stream*, std::string&) {
  memcpy(d, s, n);
  return VPZ_OK;
}
int fill(void* d, int v, size_t n, Stream*, std::string&) {
  memset(d, v, n);
  return VPZ_OK;
}

int launch_k0(const K0Params& p, Stream*, std::string&) {
  if (p.n_files == 0) return VPZ_OK;
  emu::launch(1, K0_THREADS, 0, [&] { k0_walk_cta(p); });
  emu::launch(2, K0_THREADS, K0_SMEM_WORDS * 4, [&] { k0_crc_cta(p, (uint32_t*)emu::t_block->smem); });
  K0Params q = p;
  q.only_irregular = 1;
  emu::launch(1, K0_THREADS, K0_SMEM_WORDS * 4, [&] { k0_cta(q, (uint32_t*)emu::t_block->smem); });
  return VPZ_OK;
}

int launch_k0g(const K0gParams& p, Stream*, std::string&) {
  if (p.n_files == 0) return VPZ_OK;
  emu::launch(1, 64, 0, [&] { k0g_cta(p); });
  return VPZ_OK;
}

int launch_k4(const K4Params& p, Stream*, std::string&) {
  if (p.n_segs == 0) return VPZ_OK;
  emu::launch(2, 64, 0, [&] { k4_cta(p); });
  return VPZ_OK;
}

int launch_k1a(const K1Params& p, bool debug, bool full, int blocks, Stream*, std::string&) {
  if (p.n_pkts == 0) return VPZ_OK;
  (void)blocks;  // work stealing: one emulated block drains the whole queue
  if (!debug && !full && p.k1a_smem) {
    static uint32_t s_ctl[2];  // the emulator runs one block at a time
    emu::launch(1, K1A_SM_THREADS, ((size_t)K1A_SM_WORDS + 128) * 4 + K1A_RING_BYTES(K1A_SM_THREADS), [&] { k1a_sm_loop(p, s_ctl); });
    return VPZ_OK;
  }
  emu::launch(1, 128, K1A_RING_BYTES(128) + K1A_CLS_BYTES(128), [&] {
    const int lane = threadIdx.x & 31;
    uint4* ring = reinterpret_cast<uint4*>(emu::t_block->smem) + 2 * threadIdx.x;
    uint32_t* cls = reinterpret_cast<uint32_t*>(emu::t_block->smem) + K1A_RING_BYTES(128) / 4 + threadIdx.x;
    for (;;) {
      uint32_t base = 0;
      if (lane == 0) base = atomicAdd(p.counter, 32u);
      base = __shfl_sync(0xffffffffu, base, 0);
      if (base >= p.n_pkts) break;
      const uint32_t i = base + lane;
      if (i < p.n_pkts) {
        const uint32_t k = p.order ? p.order[i] : i;
        if (debug) {
          if (full) k1a_decode_packet<true, true>(p, k, ring, cls); else k1a_decode_packet<true, false>(p, k, ring, cls);
        } else {
          if (full) k1a_decode_packet<false, true>(p, k, ring, cls); else k1a_decode_packet<false, false>(p, k, ring, cls);
        }
      }
      __syncwarp();
    }
  });
  return VPZ_OK;
}

int launch_k1b(const K1Params& p, bool debug, int blocks, int warps, Stream*, std::string&) {
  if (p.n_pkts == 0) return VPZ_OK;
  (void)blocks;
  static uint32_t s_idx;  // the emulator runs one block at a time
  emu::launch(1, K1B_THREADS, (size_t)p.smem_words_per_warp * 4 * (p.gather_ok ? warps : 1) + (p.gather_ok ? 1024 : 0), [&] {
    uint32_t* smem = (uint32_t*)emu::t_block->smem;
    if (p.gather_ok) {
      if (debug) k1b_gather_loop<true>(p, smem); else k1b_gather_loop<false>(p, smem);
    } else {
      if (debug) k1b_general_loop<true>(p, smem, &s_idx); else k1b_general_loop<false>(p, smem, &s_idx);
    }
  });
  return VPZ_OK;
}

int launch_k3(const K3Params& p, int ncb, size_t smem_bytes, Stream*, std::string&) {
  if (p.n_items == 0) return VPZ_OK;
  emu::launch(2, (unsigned)ncb * K3_THREADS_PER_CH, smem_bytes, [&] {
    float* smem = (float*)emu::t_block->smem;
    if (p.out16) k3_cta_loop<true>(p, smem, ncb); else k3_cta_loop<false>(p, smem, ncb);
  });
  return VPZ_OK;
}

int launch_k3_streams(const K3Params& p, Stream*, std::string&) {
  if (p.n_items == 0) return VPZ_OK;
  const unsigned groups = p.n_items >= 3 ? 3 : 1;   // independent 64-thread workers per emulated CTA
  emu::launch(2, groups * K3_THREADS_PER_CH, ((size_t)K3S_TAB_FLOATS + groups * K3S_GROUP_FLOATS) * 4, [&] {
    float* sm = (float*)emu::t_block->smem;
    if (p.res) {
      if (p.out16) k3s_cta<true, true>(p, sm); else k3s_cta<false, true>(p, sm);
    } else {
      if (p.out16) k3s_cta<true, false>(p, sm); else k3s_cta<false, false>(p, sm);
    }
  });
  return VPZ_OK;
}

}  // namespace dev
}  // namespace vpz
