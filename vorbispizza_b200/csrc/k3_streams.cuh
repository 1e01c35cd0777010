// k3_streams.cuh -- K3 for the common case (block sizes 256 / 2048, mono or stereo): one resident CTA
// per SM, split into independent 64-thread GROUPS; every group is a worker that takes whole work
// items (a run of <= 63 consecutive packets of one stream + its seed) from the global counter.
//
// Per packet a group transforms the channels one after the other (IMDCT through a 512-point complex
// FFT, k3_imdct.cuh) into the D slots of the stream and then writes the windowed, overlap-added,
// clipped, INTERLEAVED samples of all channels (full 32-byte sectors for stereo).  Groups never wait
// for each other: the only barriers are 64-thread named barriers, the twiddle / window tables are
// staged once per CTA, a mono stream occupies one group like a stereo one, and the spectrum of the
// next block is already on its way while the current one is written out (channel 0 into registers,
// channel 1 with cp.async into the D slots of channel 1 that the finished packet no longer needs).
//
// Replaces (reference file:line): Mdct.Reverse (Mdct.cs:15-19,77-419), StreamDecoder.OverlapBuffers
// (StreamDecoder.cs:764-791), the valid-range bookkeeping of ReadNextPacket (StreamDecoder.cs:640-694,
// geometry from the host) and StoreInterleaved<Clip> (StreamDecoder.cs:515-592, Utils.cs:44-58).
#pragma once
#include "k3_imdct.cuh"

// shared memory (floats): tables, then per group: descriptors | transpose scratch | D slots of 2 channels
#define K3S_TAB_S_TW 3072        // short block: pre/post twiddle, 64 complex
#define K3S_TAB_S_W64 3200       // short block: 64-point roots, 64 complex
#define K3S_TAB_S_SLOPE 3328     // short window slope, 128 floats
#define K3S_TAB_FLOATS 3456
#define K3S_DESC_PKTS 32
#define K3S_DESC_FLOATS 192      // 32 x (4 descriptor words + 1 exec mask) + work-stealing slot
#define K3S_CH_FLOATS 1536       // D slots of one channel: Hi[512] Lo0[512] Lo1[512]
// 1: both channels of a stereo long block are transformed in one pass (fft512_pair_to_D): a second transpose scratch
// per worker, fewer workers with more registers each
#ifndef K3S_PAIR
#define K3S_PAIR 0
#endif
#define K3S_SCRATCH_FLOATS ((K3S_PAIR ? 4 : 2) * K3_PLANE)
#define K3S_GROUP_FLOATS (K3S_DESC_FLOATS + K3S_SCRATCH_FLOATS + 2 * K3S_CH_FLOATS)
#ifndef K3S_MAX_GROUPS
#define K3S_MAX_GROUPS (K3S_PAIR ? 9 : 12)
#endif
#ifndef K3S_EMIT_UNROLL
#define K3S_EMIT_UNROLL 4
#endif
// 1: read a channel's spectrum only below VpzPktRes.end16 (K1b does not write the zero tail); 0: K1b writes
// every bin and the loads are unconditional
#ifndef VPZ_K3_END
#define VPZ_K3_END 1
#endif
// ENDS is a template parameter of the kernel: spectra that did not come from K1b (vpz_synth_create) are
// complete, and the kernel-only path must not pay for the tests
#if VPZ_K3_END
#define K3S_END16(rw, c) (ENDS ? (int)(((rw) >> (16 + 8 * (c))) & 0xffu) : 255)
#else
#define K3S_END16(rw, c) 255
#endif

// 1: the spectrum of the coming long block arrives by bulk copies (cp.async.bulk + mbarrier, one elected
// thread per group): channel 0 into the transpose scratch -- in the 72-strided layout of the first transpose,
// so every thread reads back exactly the words it will overwrite and no extra barrier is needed -- and
// channel 1 into its own free D slots.  0: channel 0 in registers (8 LDG.64 per thread, live across the output
// loop), channel 1 by per-thread 16-byte cp.async.
#ifndef K3S_BULK
#define K3S_BULK 0
#endif
// 1: the global packet index (needed only when a packet clipped) is rebuilt from shared memory instead of being
// carried -- and spilled -- through the packet loop
// 1: per-item values the packet loop needs once per packet (output base) live in shared memory, the window slopes of
// the rare general output path come from the staged tables instead of two global pointers
#ifndef K3S_LEAN
#define K3S_LEAN 1
#endif
#ifndef K3S_GP_SMEM
#define K3S_GP_SMEM 1
#endif
#define K3S_MBAR_FLOAT 176       // two mbarriers (channel 0, channel 1) in the unused tail of the descriptor area

#ifndef VPZ_EMU
VPZ_DEV unsigned k3s_saddr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
VPZ_DEV void k3s_mbar_init(uint64_t* b) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(k3s_saddr(b)) : "memory");
}
VPZ_DEV void k3s_mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// the scratch / D slots were last read through the generic proxy (ordered by the group barrier before this point)
VPZ_DEV void k3s_bulk_begin(uint64_t* b, unsigned bytes) {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(k3s_saddr(b)), "r"(bytes) : "memory");
}
VPZ_DEV void k3s_proxy_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
VPZ_DEV void k3s_bulk(float* dst_smem, const float* src, unsigned bytes, uint64_t* b) {   // 16-byte aligned, bytes % 16 == 0
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(k3s_saddr(dst_smem)),
               "l"(src), "r"(bytes), "r"(k3s_saddr(b))
               : "memory");
}
VPZ_DEV void k3s_mbar_wait(uint64_t* b, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(k3s_saddr(b)),
      "r"(parity)
      : "memory");
}
#else
// emulator: the copy happens at issue; the consumer is always behind a group barrier
VPZ_DEV void k3s_mbar_init(uint64_t*) {}
VPZ_DEV void k3s_mbar_init_fence() {}
VPZ_DEV void k3s_bulk_begin(uint64_t*, unsigned) {}
VPZ_DEV void k3s_proxy_fence() {}
VPZ_DEV void k3s_bulk(float* dst_smem, const float* src, unsigned bytes, uint64_t*) { memcpy(dst_smem, src, bytes); }
VPZ_DEV void k3s_mbar_wait(uint64_t*, unsigned) {}
#endif

#ifndef VPZ_EMU
VPZ_DEV void k3s_cp16(float* dst_smem, const float* src) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(dst_smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(a), "l"(src) : "memory");
}
VPZ_DEV void k3s_cp8(float2* dst_smem, const float2* src) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(dst_smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(a), "l"(src) : "memory");
}
VPZ_DEV void k3s_cp_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }
#else
VPZ_DEV void k3s_cp16(float* dst_smem, const float* src) { memcpy(dst_smem, src, 16); }
VPZ_DEV void k3s_cp8(float2* dst_smem, const float2* src) { memcpy(dst_smem, src, 8); }
VPZ_DEV void k3s_cp_wait() {}
#endif

// The common cases in full: a block after a block of the same size, nothing trimmed (ls = 0,
// count = L = M, previous RightStart = M; M = 1024 long after long, M = 128 short after short).
// Sample j < M/2 reads +D[j + M/2] of the current block and -Dp[M/2 - 1 - j] of the previous one; its mirror
// image M - 1 - j reads the SAME two values with the window pair swapped (time-domain aliasing
// symmetry), so a thread produces both from one set of loads:
//   out[j]         = D[M/2 + j] * w[j]            + (-Dp[M/2 - 1 - j]) * w[M - 1 - j]
//   out[M - 1 - j] = (-D[M/2 + j]) * w[M - 1 - j] + (-Dp[M/2 - 1 - j]) * w[j]
// with the rounding order of OverlapBuffers (two rounded products, one rounded sum).
template <int NC, bool CLIP, bool OUT16, int M>
VPZ_DEV bool k3s_emit_same_size(const float* hi0 /* D[M/2..] of channel 0 */, const float* plo0 /* previous D[0..M/2) */,
                                const float* ws, float* outp, int t64) {
  bool clipped = false;
  const bool pair_ok = NC == 2 && (reinterpret_cast<uintptr_t>(outp) & (OUT16 ? 3u : 7u)) == 0;
  constexpr int HALF = M / 2;
  constexpr int UNROLL = K3S_EMIT_UNROLL;
#pragma unroll UNROLL
  for (int r = 0; r < HALF / 64; r++) {
    const int j = t64 + 64 * r;
    const float w0 = ws[j], w1 = ws[M - 1 - j];
    float lo[NC], hi[NC];
#pragma unroll
    for (int c = 0; c < NC; c++) {
      const float a = hi0[c * K3S_CH_FLOATS + j];
      const float b = -plo0[c * K3S_CH_FLOATS + HALF - 1 - j];
      float x = __fadd_rn(__fmul_rn(a, w0), __fmul_rn(b, w1));
      float y = __fadd_rn(__fmul_rn(-a, w1), __fmul_rn(b, w0));
      if (CLIP) {  // Utils.ClipValue (Utils.cs:44-58): |v| > c -> +-c, anything else (NaN included) unchanged
        const bool px = fabsf(x) > 0.99999994f, py = fabsf(y) > 0.99999994f;
        x = px ? copysignf(0.99999994f, x) : x;
        y = py ? copysignf(0.99999994f, y) : y;
        clipped |= px | py;
      }
      lo[c] = x;
      hi[c] = y;
    }
    if (NC == 2) {
      if (pair_ok) {
        k3_put2<OUT16>(outp, (size_t)(2 * j), lo[0], lo[NC - 1]);
        k3_put2<OUT16>(outp, (size_t)(2 * (M - 1 - j)), hi[0], hi[NC - 1]);
      } else {
        k3_put<OUT16>(outp, (size_t)(2 * j), lo[0]);
        k3_put<OUT16>(outp, (size_t)(2 * j + 1), lo[NC - 1]);
        k3_put<OUT16>(outp, (size_t)(2 * (M - 1 - j)), hi[0]);
        k3_put<OUT16>(outp, (size_t)(2 * (M - 1 - j) + 1), hi[NC - 1]);
      }
    } else {
      k3_put<OUT16>(outp, (size_t)j, lo[0]);
      k3_put<OUT16>(outp, (size_t)(M - 1 - j), hi[0]);
    }
  }
  return clipped;
}

// one work item, C = 1 or 2 channels, by one 64-thread group.  OUT16: 16-bit PCM (k3_s16) instead of fp32
template <bool OUT16, bool ENDS>
VPZ_DEV void k3s_run_item(const K3Params& P, const VpzOlaItem& item, float* tabs, float* gbase, int grp, int t64, unsigned& mph) {
  const VpzOlaItem it = item;  // the item lives in global memory: read it once
  const uint32_t* blob = P.setups[it.setup_slot];
  const VpzSetupHdr* Hd = reinterpret_cast<const VpzSetupHdr*>(blob);
  const int C = Hd->channels;
  const int t = k3_remap64(t64);
#if !K3S_LEAN
  const float* slope0 = reinterpret_cast<const float*>(blob + Hd->slope_off[0]);
  const float* slope1 = reinterpret_cast<const float*>(blob + Hd->slope_off[1]);
#endif
  VpzPktOla* spk = reinterpret_cast<VpzPktOla*>(gbase);                       // [K3S_DESC_PKTS] descriptors
  uint32_t* smask = reinterpret_cast<uint32_t*>(gbase) + 4 * K3S_DESC_PKTS;    // [K3S_DESC_PKTS] exec masks
  volatile int* snb = reinterpret_cast<volatile int*>(gbase) + (K3S_DESC_FLOATS - 2);   // packets staged in this batch
  float* T = gbase + K3S_DESC_FLOATS;
  float* Dch = T + K3S_SCRATCH_FLOATS;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(gbase + K3S_MBAR_FLOAT);   // [0]: channel 0, [1]: channel 1; phases in mph
  (void)mbar;
  const cpx* tab = reinterpret_cast<const cpx*>(tabs);
  const cpx* tw_s = reinterpret_cast<const cpx*>(tabs + K3S_TAB_S_TW);
  const cpx* w64_s = reinterpret_cast<const cpx*>(tabs + K3S_TAB_S_W64);

  int prevM = 0, prev_rs = 0, prev_re = 0;   // previous packet: M, RightStart, RightEnd
  bool have_prev = false;
  int parity = 0;
  const int first = (int)it.first_pkt - (it.has_pre ? 1 : 0);
  const int total = (int)it.n_pkts + (it.has_pre ? 1 : 0);
  float2 xr[8];            // spectrum of channel 0 of the coming long block (prefetched)
  bool xr_valid = false;
  bool staged = false;     // spectrum of channel 1 of the coming long block sits in channel 1's free D slots
#if K3S_PAIR
  bool xb_valid = false;   // ... or, when both channels go through the paired transform, in the second scratch
#endif
  for (int pb = 0; pb < total; pb += K3S_DESC_PKTS) {
    const int nb = (total - pb) < K3S_DESC_PKTS ? (total - pb) : K3S_DESC_PKTS;
    // descriptors + exec masks of the next nb packets: one parallel fetch
    K3_GSYNC(grp);
    if (t64 == 0) {
      snb[0] = nb;
#if K3S_LEAN
      *reinterpret_cast<volatile unsigned long long*>(snb - 4) = (unsigned long long)it.out_base;   // read back per packet
#endif
#if K3S_GP_SMEM
      snb[-1] = first + pb;   // global index of the batch's first packet (only a clipped packet needs it)
#endif
    }
    if (t64 < nb) {
      spk[t64] = P.pkts[first + pb + t64];
      // exec mask | status << 8 | end16[0] << 16 | end16[1] << 24 (VpzPktRes); no K1: every channel, every bin
      smask[t64] = ENDS ? reinterpret_cast<const uint32_t*>(P.res)[first + pb + t64] : 0xffff00ffu;
    }
    K3_GSYNC(grp);
    // the bound is read back from shared memory: as a register it was spilled, and its reload at the top of every
    // packet missed L1 (7.6 % of the kernel's stall samples on one compare)
    for (int pw = 0; pw < snb[0]; pw++, parity ^= 1) {
      const int pi = pb + pw;
#if !K3S_GP_SMEM
      const uint32_t gp = (uint32_t)(first + pi);
#endif
      const VpzPktOla pk = spk[pw];
      const uint32_t rw = ENDS ? smask[pw] : 0xffff00ffu;
      const uint32_t mask = rw & 0xffu;
      const bool has_next = pw + 1 < snb[0];
      const VpzPktOla pk_next = spk[has_next ? pw + 1 : pw];
      const uint32_t rw_next = ENDS ? smask[has_next ? pw + 1 : pw] : 0xffff00ffu;
      const uint32_t mask_next = rw_next & 0xffu;
      const bool next_long = has_next && (pk_next.flags & VPZ_OLA_LONG);
      const bool is_long = pk.flags & VPZ_OLA_LONG;
      const int M = is_long ? 1024 : 128;
      const int h = M >> 1;
      const bool emit = !(pi == 0 && it.has_pre) && !(pk.flags & VPZ_OLA_NOOUT) && have_prev;

      // ---- transform the channels into their D slots -------------------------------------------------
      if (!is_long) {
        // short block: 8 threads per channel run its 64-point FFT, both channels at the same time; a
        // channel without floor energy is cleared instead (Mapping.cs:185-194)
        for (int c = 0; c < C; c++) {
          if ((mask >> c) & 1u) continue;
          float* chb = Dch + c * K3S_CH_FLOATS;
          for (int i = t64; i < 64; i += K3_THREADS_PER_CH) {
            chb[i] = 0.f;                              // D[64..128): high slot
            chb[512 + parity * 512 + i] = 0.f;         // D[0..64): low slot of this parity
          }
        }
        const int c = (t64 >> 3) < C ? (t64 >> 3) : 0;
        float* chb = Dch + c * K3S_CH_FLOATS;
        K3D D;
        D.h = h;
        D.hm = chb - h;
        D.lo = chb + 512 + parity * 512;
        const bool active = t64 < 8 * C && ((mask >> c) & 1u);
        fft64_to_D(P.spec + pk.spec_off + (size_t)c * M, T + 80 * c, D, tw_s, w64_s, t64 & 7, active, grp,
                   K3S_END16(rw, c) * 16);   // ends with a group barrier
        if (P.dbg_imdct) {
          for (int c2 = 0; c2 < C; c2++) {
            K3D D2;
            D2.h = h;
            D2.hm = Dch + c2 * K3S_CH_FLOATS - h;
            D2.lo = Dch + c2 * K3S_CH_FLOATS + 512 + parity * 512;
            float* dy = P.dbg_imdct + 2 * (size_t)pk.spec_off + (size_t)c2 * 2 * M;
            for (int i = t64; i < 2 * M; i += K3_THREADS_PER_CH) dy[i] = k3_y(D2, M, i);
          }
        }
      } else
#if K3S_PAIR
      if (C == 2 && (mask & 3u) == 3u) {
        K3D Da, Db;
        Da.h = Db.h = h;
        Da.hm = Dch - h;
        Da.lo = Dch + 512 + parity * 512;
        Db.hm = Dch + K3S_CH_FLOATS - h;
        Db.lo = Dch + K3S_CH_FLOATS + 512 + parity * 512;
        const float* X = P.spec + pk.spec_off;
        float2 xb[8];
        if (xb_valid) {
          // channel 1's pairs were copied into the second scratch by this very thread (8-byte cp.async during the
          // previous packet's output), at the places it overwrites in the first transpose: no barrier needed
          k3s_cp_wait();
          const int end2 = K3S_END16(rw, 1) * 8;
          const float2* B2 = reinterpret_cast<const float2*>(T + 2 * K3_PLANE);
#pragma unroll
          for (int q = 0; q < 8; q++) xb[q] = 64 * q < end2 ? B2[72 * q + t] : float2{0.f, 0.f};
          xb_valid = false;
        } else {
          k3_load_x(X + M, t, xb, K3S_END16(rw, 1) * 8);
        }
        if (!xr_valid) k3_load_x(X, t, xr, K3S_END16(rw, 0) * 8);
        xr_valid = false;
        fft512_pair_to_D(xr, xb, T, T + 2 * K3_PLANE, Da, Db, tab, t, grp);
        K3_GSYNC(grp);   // D complete, scratches reusable
        if (P.dbg_imdct) {
          for (int c = 0; c < 2; c++) {
            float* dy = P.dbg_imdct + 2 * (size_t)pk.spec_off + (size_t)c * 2 * M;
            for (int i = t64; i < 2 * M; i += K3_THREADS_PER_CH) dy[i] = k3_y(c ? Db : Da, M, i);
          }
        }
      } else
#endif
      for (int c = 0; c < C; c++) {
        float* chb = Dch + c * K3S_CH_FLOATS;
        K3D D;
        D.h = h;
        D.hm = chb - h;
        D.lo = chb + 512 + parity * 512;
        const bool exec = (mask >> c) & 1u;
        const float* X = P.spec + pk.spec_off + (size_t)c * M;
        if (!exec) {
          // a channel without floor energy outputs zeros (Mapping.cs:185-194) but still takes part in
          // the overlap-add: its D slots are cleared instead of transformed
          for (int i = t64; i < M; i += K3_THREADS_PER_CH) *k3_dp(D, i) = 0.f;
#if !(K3S_BULK & 1)
          k3s_cp_wait();
#endif
          K3_GSYNC(grp);
        } else {
#if K3S_BULK & 2
          if (c == 0 && xr_valid) {
            // channel 0's spectrum was bulk-copied into the transpose scratch during the previous packet's
            // output, chunk q (64 pairs) at complex offset 72 q: thread t reads T2[72 q + t], the very words it
            // writes in the first transpose
            k3s_mbar_wait(mbar, mph & 1u);
            mph ^= 1u;
            const int end2 = K3S_END16(rw, 0) * 8;
            const float2* T2 = reinterpret_cast<const float2*>(T);
#pragma unroll
            for (int q = 0; q < 8; q++) xr[q] = 64 * q < end2 ? T2[72 * q + t] : float2{0.f, 0.f};
          } else
#endif
          if (c == 1 && staged) {
            // channel 1's spectrum was copied into its own Hi / Lo[parity] slots after the previous
            // packet's output (made visible by the barrier that ended channel 0's transform)
#if K3S_BULK & 1
            k3s_mbar_wait(mbar + 1, (mph >> 1) & 1u);
            mph ^= 2u;
#endif
#pragma unroll
            for (int q = 0; q < 8; q++) {
              const int n2 = 2 * (t + 64 * q);
              xr[q] = *reinterpret_cast<const float2*>(n2 < 512 ? chb + n2 : D.lo + (n2 - 512));
            }
          } else if (!(c == 0 && xr_valid)) {
            k3_load_x(X, t, xr, K3S_END16(rw, c) * 8);
          }
          fft512_to_D(xr, T, D, tab, t, grp);
          if (c == 0) xr_valid = false;
#if !(K3S_BULK & 1)
          k3s_cp_wait();
#endif
          K3_GSYNC(grp);   // D complete, scratch reusable, staged channel-1 spectrum visible
        }
        if (P.dbg_imdct) {
          float* dy = P.dbg_imdct + 2 * (size_t)pk.spec_off + (size_t)c * 2 * M;
          for (int i = t64; i < 2 * M; i += K3_THREADS_PER_CH) dy[i] = k3_y(D, M, i);
        }
      }
      staged = false;

      // channel 0 of the next long block is requested now and lands during the output loop
      if (next_long && (mask_next & 1u)) {
#if K3S_BULK & 2
        // every transform above ended with a group barrier: the scratch is free until the next transform
        if (t64 < 8) {   // lane q issues chunk q
          int nq = (K3S_END16(rw_next, 0) * 8 + 63) >> 6;   // 64-pair chunks that were written (<= 8)
          nq = nq < 8 ? nq : 8;
          const float* Xn = P.spec + pk_next.spec_off;
          if (t64 == 0) k3s_bulk_begin(mbar, 512u * (unsigned)nq);
          if (t64 < nq) {
            if (t64 != 0) k3s_proxy_fence();
            k3s_bulk(T + 144 * t64, Xn + 128 * t64, 512u, mbar);
          }
        }
#else
        k3_load_x(P.spec + pk_next.spec_off, t, xr, K3S_END16(rw_next, 0) * 8);
#endif
        xr_valid = true;
      }
#if K3S_PAIR
      if (C == 2 && next_long && (mask_next & 3u) == 3u) {
        // channel 1 of a block that will take the paired transform: every thread copies its own 8 pairs
        const float2* Xn = reinterpret_cast<const float2*>(P.spec + pk_next.spec_off + 1024);
        float2* B2 = reinterpret_cast<float2*>(T + 2 * K3_PLANE);
        const int end2 = K3S_END16(rw_next, 1) * 8;
#pragma unroll
        for (int q = 0; q < 8; q++)
          if (64 * q < end2) k3s_cp8(B2 + 72 * q + t, Xn + t + 64 * q);
        xb_valid = true;
      }
#endif
      // ---- output: window + overlap-add + clip, all channels interleaved ----------------------------
      if (emit) {
        const int ls = pk.left_start;
        const int count = (int)pk.right_start - ls;
        const int L = prev_re - prev_rs;             // StreamDecoder.cs:654
        // element offset of the packet's first sample; an s16 element is 2 bytes, so the float-typed base
        // advances by half the element offset (out_base and out_off * C are element counts)
#if K3S_LEAN
        const size_t eoff = (size_t)*reinterpret_cast<volatile const unsigned long long*>(snb - 4) + (size_t)pk.out_off * C;
#else
        const size_t eoff = (size_t)it.out_base + (size_t)pk.out_off * C;
#endif
        float* outp = OUT16 ? reinterpret_cast<float*>(reinterpret_cast<int16_t*>(P.pcm) + eoff) : P.pcm + eoff;
        const float* Dp_lo = Dch + 512 + (parity ^ 1) * 512;
        bool clipped;
        const bool same = prevM == M && ls == 0 && count == M && L == M && prev_rs == M;
        const bool clip = P.clip != 0;   // ClipSamples applies before the 16-bit conversion, which clamps on its own
        if (same && M == 1024 && (pk.flags & VPZ_OLA_LEFT1)) {
          const float* ws = tabs + K3_TAB_SLOPE;
          if (C == 2)
            clipped = clip ? k3s_emit_same_size<2, true, OUT16, 1024>(Dch, Dp_lo, ws, outp, t64) : k3s_emit_same_size<2, false, OUT16, 1024>(Dch, Dp_lo, ws, outp, t64);
          else
            clipped = clip ? k3s_emit_same_size<1, true, OUT16, 1024>(Dch, Dp_lo, ws, outp, t64) : k3s_emit_same_size<1, false, OUT16, 1024>(Dch, Dp_lo, ws, outp, t64);
        } else if (same && M == 128) {
          const float* ws = tabs + K3S_TAB_S_SLOPE;
          if (C == 2)
            clipped = clip ? k3s_emit_same_size<2, true, OUT16, 128>(Dch, Dp_lo, ws, outp, t64) : k3s_emit_same_size<2, false, OUT16, 128>(Dch, Dp_lo, ws, outp, t64);
          else
            clipped = clip ? k3s_emit_same_size<1, true, OUT16, 128>(Dch, Dp_lo, ws, outp, t64) : k3s_emit_same_size<1, false, OUT16, 128>(Dch, Dp_lo, ws, outp, t64);
        } else {
#if K3S_LEAN
          // every setup of this launch has the block sizes 256 / 2048: both window slopes are in the staged tables
          const float* w = tabs + ((pk.flags & VPZ_OLA_LEFT1) ? K3_TAB_SLOPE : K3S_TAB_S_SLOPE);
#else
          const float* w = (pk.flags & VPZ_OLA_LEFT1) ? slope1 : slope0;
#endif
          const float* Dc_hm = Dch - h;
          const float* Dc_lo = Dch + 512 + parity * 512;
          if (C == 2)
            clipped = clip ? k3_emit<2, true, OUT16>(Dc_hm, Dc_lo, Dp_lo, K3S_CH_FLOATS, M, prevM, ls, count, prev_rs, L, w, outp, C, t64, K3_THREADS_PER_CH)
                           : k3_emit<2, false, OUT16>(Dc_hm, Dc_lo, Dp_lo, K3S_CH_FLOATS, M, prevM, ls, count, prev_rs, L, w, outp, C, t64, K3_THREADS_PER_CH);
          else
            clipped = clip ? k3_emit<1, true, OUT16>(Dc_hm, Dc_lo, Dp_lo, K3S_CH_FLOATS, M, prevM, ls, count, prev_rs, L, w, outp, C, t64, K3_THREADS_PER_CH)
                           : k3_emit<1, false, OUT16>(Dc_hm, Dc_lo, Dp_lo, K3S_CH_FLOATS, M, prevM, ls, count, prev_rs, L, w, outp, C, t64, K3_THREADS_PER_CH);
        }
        // per packet: 0 when any sample was clamped (HasClipped), else stays 0xffffffff
#if K3S_GP_SMEM
        if (P.clip_first && clipped) atomicMin(P.clip_first + (uint32_t)(snb[-1] + pw), 0u);
#else
        if (P.clip_first && clipped) atomicMin(P.clip_first + gp, 0u);
#endif
      }
      prevM = M;
      prev_rs = pk.right_start;
      prev_re = pk.right_end;
      have_prev = true;
      // the output loop read this packet's high slot and the previous packet's low slot: both are free
      // once every thread of the group is here
      K3_GSYNC(grp);
      if (C == 2 && next_long && (mask_next & 2u) && !(K3S_PAIR && (mask_next & 3u) == 3u)) {
        // channel 1 of the next long block -> its Hi slot (X[0..512)) and the Lo slot of the next parity
        // (X[512..1024)), asynchronously; channel 0's transform runs meanwhile
        const float* Xn = P.spec + pk_next.spec_off + 1024;
        float* hi = Dch + K3S_CH_FLOATS;
        float* lo = hi + 512 + (parity ^ 1) * 512;
        int end4 = K3S_END16(rw_next, 1) * 4;   // float4 groups that were written (the rest is +0)
#if K3S_BULK & 1
        end4 = end4 > 256 ? 256 : ((end4 + 31) & ~31);   // whole 128-bin units, as K1b fills them
        if (t64 == 0) {
          const unsigned nh = (unsigned)(end4 < 128 ? end4 : 128), nl = (unsigned)(end4 - (int)nh);
          k3s_bulk_begin(mbar + 1, 16u * (nh + nl));
          if (nh) k3s_bulk(hi, Xn, 16u * nh, mbar + 1);
          if (nl) k3s_bulk(lo, Xn + 512, 16u * nl, mbar + 1);
        }
#pragma unroll
        for (int r = 0; r < 4; r++) {
          const int i4 = t64 + 64 * r;
          float* dst = i4 < 128 ? hi + 4 * i4 : lo + 4 * (i4 - 128);
          if (i4 >= end4) *reinterpret_cast<float4*>(dst) = float4{0.f, 0.f, 0.f, 0.f};
        }
#else
#pragma unroll
        for (int r = 0; r < 4; r++) {
          const int i4 = t64 + 64 * r;
          float* dst = i4 < 128 ? hi + 4 * i4 : lo + 4 * (i4 - 128);
          if ((i4 & ~31) < end4)   // 32 float4 = one 128-bin unit = one warp's share: uniform
            k3s_cp16(dst, Xn + 4 * i4);
          else
            *reinterpret_cast<float4*>(dst) = float4{0.f, 0.f, 0.f, 0.f};
        }
#endif
        staged = true;
      }
    }
  }
#if !(K3S_BULK & 1)
  k3s_cp_wait();
#endif
}

// kernel body: `groups` = blockDim.x / 64 workers per CTA
template <bool OUT16, bool ENDS>
VPZ_DEV void k3s_cta(const K3Params& P, float* smem) {
  const int tid = threadIdx.x;
  const int grp = tid / K3_THREADS_PER_CH, t64 = tid % K3_THREADS_PER_CH;
  const int nthreads = blockDim.x;
  {
    // the tables depend on the two block sizes only (256 / 2048 for every setup of this launch)
    const uint32_t* blob = P.setups[P.items[0].setup_slot];
    const VpzSetupHdr* Hd = reinterpret_cast<const VpzSetupHdr*>(blob);
    cpx* tb = reinterpret_cast<cpx*>(smem);
    const cpx* tw1 = reinterpret_cast<const cpx*>(blob + Hd->tw_off[1]);
    const cpx* w512 = reinterpret_cast<const cpx*>(blob + Hd->fft_off[1]);
    const cpx* tw0 = reinterpret_cast<const cpx*>(blob + Hd->tw_off[0]);
    const cpx* w64 = reinterpret_cast<const cpx*>(blob + Hd->fft_off[0]);
    const float* slope0 = reinterpret_cast<const float*>(blob + Hd->slope_off[0]);
    const float* slope1 = reinterpret_cast<const float*>(blob + Hd->slope_off[1]);
    for (int i = tid; i < 512; i += nthreads) tb[K3_TAB_TW + i] = VPZ_LDG(tw1 + i);
    for (int i = tid; i < 448; i += nthreads) {
      int k = (i >> 6) + 1, tt = i & 63;
      tb[K3_TAB_W1 + i] = VPZ_LDG(w512 + ((tt * k) & 511));
    }
    for (int i = tid; i < 56; i += nthreads) {
      int k = (i >> 3) + 1, r = i & 7;
      tb[K3_TAB_W2 + i] = VPZ_LDG(w512 + ((r * k) << 3));
    }
    for (int i = tid; i < 1024; i += nthreads) smem[K3_TAB_SLOPE + i] = VPZ_LDG(slope1 + i);
    cpx* ts = reinterpret_cast<cpx*>(smem + K3S_TAB_S_TW);
    cpx* tr = reinterpret_cast<cpx*>(smem + K3S_TAB_S_W64);
    for (int i = tid; i < 64; i += nthreads) {
      ts[i] = VPZ_LDG(tw0 + i);
      tr[i] = VPZ_LDG(w64 + i);
    }
    for (int i = tid; i < 128; i += nthreads) smem[K3S_TAB_S_SLOPE + i] = VPZ_LDG(slope0 + i);
  }
  float* gbase = smem + K3S_TAB_FLOATS + grp * K3S_GROUP_FLOATS;
  unsigned mph = 0;   // phase bits of the group's two mbarriers
#if K3S_BULK
  if (t64 == 0) {
    k3s_mbar_init(reinterpret_cast<uint64_t*>(gbase + K3S_MBAR_FLOAT));
    k3s_mbar_init(reinterpret_cast<uint64_t*>(gbase + K3S_MBAR_FLOAT) + 1);
    k3s_mbar_init_fence();
  }
#endif
  __syncthreads();
  uint32_t* slot = reinterpret_cast<uint32_t*>(gbase) + (K3S_DESC_FLOATS - 1);
  for (;;) {
    K3_GSYNC(grp);
    if (t64 == 0) *slot = atomicAdd(P.counter, 1u);
    K3_GSYNC(grp);
    const uint32_t idx = *slot;
    if (idx >= P.n_items) break;
    k3s_run_item<OUT16, ENDS>(P, P.items[idx], smem, gbase, grp, t64, mph);
  }
}
