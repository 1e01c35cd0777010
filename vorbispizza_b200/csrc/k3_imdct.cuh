// k3_imdct.cuh -- K3: spectrum -> interleaved float PCM.  IMDCT + window + overlap-add + clip +
// interleave in one kernel; one CTA walks a run of consecutive packets of one stream.
//
// Replaces (reference file:line):
//   Mdct.Reverse / MdctImpl.CalcReverse     Mdct.cs:15-19,77-419   (different algorithm, same transform)
//   StreamDecoder.OverlapBuffers            StreamDecoder.cs:764-791
//   valid-range bookkeeping of ReadNextPacket  StreamDecoder.cs:640-694 (geometry comes from the host)
//   StoreInterleaved<Clip> / Utils.ClipValue   StreamDecoder.cs:515-592, Utils.cs:44-58
//
// Transform: y[i] = sum_k X[k] cos(pi/(2N) (2i+1+N/2)(2k+1)) is a DCT-IV of size M = N/2 in
// disguise: y[i] = D[i+M/2] (i < M/2), -D[3M/2-1-i] (M/2 <= i < 3M/2), -D[i-3M/2] (i >= 3M/2).
// D is computed with one H = N/4 point complex FFT:
//   z[n] = (X[2n] + i X[M-1-2n]) * tw[n],  tw[n] = exp(-i pi (n + 1/8) / M)
//   T = FFT_H(z);  c[p] = T[p] * tw[p];  D[2p] = Re c[p];  D[M-1-2p] = -Im c[p]
// The FFT runs as radix-8 passes in registers with conflict-free shared-memory transposes
// (8x8x8 for N = 2048, 8x8 for N = 256); other block sizes use a radix-2 Stockham loop.
// Only D (M floats per channel) is kept.  The left half of y reads D[M/2..M) only and the right half
// D[0..M/2) only, so the next packet needs just the LOW half of its predecessor's D: shared memory
// holds one high-half slot and two low-half slots (ping-pong) per channel, 1.5 M floats.
#pragma once
#include "k1_params.h"

#ifndef VPZ_EMU
#define VPZ_DEV __device__ __forceinline__
#define VPZ_LDG(p) __ldg(p)
#else
#define VPZ_DEV inline
#define VPZ_LDG(p) (*(p))
#endif

#define K3_THREADS_PER_CH 64

// D[i] of the current block lives in two places: i < h (= M/2) in the low-half slot, i >= h in the
// high-half slot.  `hm` is the high slot minus h, so both take the plain index.
struct K3D {
  float* lo;
  float* hm;
  int h;
};
VPZ_DEV float* k3_dp(const K3D& d, int i) { return (i < d.h ? d.lo : d.hm) + i; }

// barrier of one 64-thread group (2 warps) of the fast kernel: hardware barrier 1 + group index (a
// CTA of that kernel is alone on its SM, so reserving all 16 barriers costs nothing)
#ifndef VPZ_EMU
#define K3_GSYNC(g) asm volatile("bar.sync %0, 64;" ::"r"((g) + 1) : "memory")
#else
#define K3_GSYNC(g) emu_named_barrier((g) + 1, 64)
#endif

typedef float2 cpx;
// 1: complex multiplies and the W8 rotations of the 8-point DFT in packed fp32 (Blackwell FMUL2 / FFMA2).  A packed
// operand may be a scalar broadcast (`R.F32`) or a register pair with its halves swapped and one half negated
// (`-R.F32x2.LO_HI.NP`), so a * b = (a.x, a.y) * b.x + (-a.y, a.x) * b.y is TWO instructions instead of four and
// needs no moves: ptxas folds the pack / swap / negate of the inline PTX below into the operand modifiers.
// Only the transform uses it (results within the IMDCT tolerance, same contraction as the scalar code: one rounded
// product, one fused multiply-add); the overlap-add keeps its scalar, explicitly rounded form.
#ifndef K3_PACKED_MUL
#define K3_PACKED_MUL 1
#endif
#if !defined(VPZ_EMU) && K3_PACKED_MUL
VPZ_DEV unsigned long long k3_pk(float a, float b) {
  unsigned long long r;
  asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
VPZ_DEV float2 k3_up(unsigned long long v) {
  float2 r;
  asm("mov.b64 {%0,%1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
  return r;
}
VPZ_DEV float2 cmul(float2 a, float2 b) {
  unsigned long long p, r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(p) : "l"(k3_pk(a.x, a.y)), "l"(k3_pk(b.x, b.x)));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(k3_pk(-a.y, a.x)), "l"(k3_pk(b.y, b.y)), "l"(p));
  return k3_up(r);
}
VPZ_DEV float2 cscale(float2 a, float h) {   // a * h
  unsigned long long r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(k3_pk(a.x, a.y)), "l"(k3_pk(h, h)));
  return k3_up(r);
}
#else
VPZ_DEV cpx cmul(cpx a, cpx b) {
  cpx r;
  r.x = a.x * b.x - a.y * b.y;
  r.y = a.x * b.y + a.y * b.x;
  return r;
}
VPZ_DEV cpx cscale(cpx a, float h) { return cpx{a.x * h, a.y * h}; }
#endif
#ifndef VPZ_EMU
// Blackwell packed fp32: one FADD2 adds both parts of a complex number (a float2 already sits in an
// aligned register pair, so no moves are needed); same round-to-nearest result as two FADDs.
VPZ_DEV cpx cadd(cpx a, cpx b) {
  cpx r;
  unsigned long long ra, rb, rc;
  asm("mov.b64 %0, {%1,%2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1,%2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rc) : "l"(ra), "l"(rb));
  asm("mov.b64 {%0,%1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(rc));
  return r;
}
VPZ_DEV cpx csub(cpx a, cpx b) {
  cpx r;
  unsigned long long ra, rb, rc;
  asm("mov.b64 %0, {%1,%2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1,%2};" : "=l"(rb) : "f"(-b.x), "f"(-b.y));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rc) : "l"(ra), "l"(rb));
  asm("mov.b64 {%0,%1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(rc));
  return r;
}
#else
VPZ_DEV cpx cadd(cpx a, cpx b) { return cpx{a.x + b.x, a.y + b.y}; }
VPZ_DEV cpx csub(cpx a, cpx b) { return cpx{a.x - b.x, a.y - b.y}; }
#endif
VPZ_DEV cpx cmul_mi(cpx a) { return cpx{a.y, -a.x}; }  // a * (-i)

// 8-point forward DFT (e^{-2 pi i qk/8}), natural order in and out.
VPZ_DEV void dft8(cpx* v) {
  const float h = 0.70710678118654752440f;
  cpx a0 = cadd(v[0], v[4]), a1 = csub(v[0], v[4]);
  cpx a2 = cadd(v[2], v[6]), a3 = cmul_mi(csub(v[2], v[6]));
  cpx a4 = cadd(v[1], v[5]), a5 = csub(v[1], v[5]);
  cpx a6 = cadd(v[3], v[7]), a7 = cmul_mi(csub(v[3], v[7]));
  cpx b0 = cadd(a0, a2), b2 = csub(a0, a2);       // even part, 4-point DFT of (v0,v2,v4,v6)
  cpx b1 = cadd(a1, a3), b3 = csub(a1, a3);
  cpx c0 = cadd(a4, a6), c2 = cmul_mi(csub(a4, a6));  // odd part, 4-point DFT of (v1,v3,v5,v7)
  cpx c1 = cadd(a5, a7), c3 = csub(a5, a7);
  // twiddles W8^1 = (1-i)/sqrt2, W8^2 = -i (already applied to c2), W8^3 = (-1-i)/sqrt2
  // d1 = c1 * (1 - i) / sqrt2 = h * (c1 + (c1.y, -c1.x)),  d3 = c3 * (-1 - i) / sqrt2 = h * ((c3.y, -c3.x) - c3):
  // one (packed) add and one (packed) scaling each, the same rounded sums and products as the scalar form
  cpx d1 = cscale(cadd(c1, cmul_mi(c1)), h);
  cpx d3 = cscale(csub(cmul_mi(c3), c3), h);
  v[0] = cadd(b0, c0);
  v[4] = csub(b0, c0);
  v[1] = cadd(b1, d1);
  v[5] = csub(b1, d1);
  v[2] = cadd(b2, c2);
  v[6] = csub(b2, c2);
  v[3] = cadd(b3, d3);
  v[7] = csub(b3, d3);
}

// thread-in-channel remap: lanes l and 31-l of a warp hold mirrored FFT inputs (t and 63-t), so the
// X[M-1-2n] operands arrive with one shuffle instead of a second, half-used global load.
VPZ_DEV int k3_remap64(int tid64) {
  int w = tid64 >> 5, l = tid64 & 31;
  return w == 0 ? (l < 16 ? l : 32 + l) : 16 + l;
}

// transposes: plane size 576 floats; both layouts are conflict-free for the writing and the reading
// pass (idx2 found by exhaustive search over linear layouts: the reader of pass 3 is thread
// t = k1 + 8 k2, so its reads B[t + 68 r] and its outputs p = t + 64 k3 are consecutive across lanes)
VPZ_DEV int idx1(int k1, int r) { return 72 * k1 + r; }                   // r in [0,64)
VPZ_DEV int idx2(int k1, int k2, int r2) { return k1 + 8 * k2 + 68 * r2; }
#define K3_PLANE 576
// FAST path tables staged in shared memory once per work item (complex entries):
//   TW[512]  pre/post twiddle of the long block, W1[7][64] = w512^(t k), W2[7][8] = w64^(r2 k)
#define K3_TAB_TW 0
#define K3_TAB_W1 512
#define K3_TAB_W2 (512 + 448)
#define K3_TAB_CPX (512 + 448 + 56)
#define K3_TAB_SLOPE 2048    // float offset of the long window slope (1024 floats) behind the complex tables
#define K3_TAB_FLOATS 3072
// Both paths: the first K3_DESC_FLOATS words of shared memory hold the descriptors (4 words) and exec
// masks (1 word) of up to K3_DESC_PKTS packets of the current work item, fetched with one parallel
// load instead of one dependent global load per packet; the last word is the work-stealing slot.
#define K3_DESC_PKTS 64
#define K3_DESC_FLOATS 384

// N = 2048: H = 512 = 8*8*8, 64 threads.  The thread's 8 float2 of the spectrum (X[2n], X[2n+1] for
// n = t + 64 q) arrive in registers (prefetched one packet ahead).  Writes D[0..1024) (smem).
// T: transpose scratch (2 planes), used for both transposes; tab: the shared-memory tables above.
// end2: the spectrum holds end2 pairs, a multiple of 64 (K1b fills up to the next 128-bin boundary); the
// bins above are an exact +0 and were not written (VpzPktRes.end16).  Load q covers pairs [64 q, 64 q + 64):
// the test is uniform over the group, so a skipped load costs nothing.
VPZ_DEV void k3_load_x(const float* X, int t, float2* xr, int end2) {
#pragma unroll
  for (int q = 0; q < 8; q++)
    xr[q] = 64 * q < end2 ? VPZ_LDG(reinterpret_cast<const float2*>(X) + (t + 64 * q)) : float2{0.f, 0.f};
}

VPZ_DEV void fft512_to_D(const float2* xr, float* T, const K3D& D, const cpx* tab, int t, int grp) {   // M = 1024
  const cpx* tw = tab + K3_TAB_TW;
  cpx v[8];
  // X[2n+1] = X[M-1-2n'] of the mirrored element n' = 511-n, held by the mirrored lane
#pragma unroll
  for (int q = 0; q < 8; q++) {
    v[q].x = xr[q].x;
    v[q].y = __shfl_xor_sync(0xffffffffu, xr[7 - q].y, 31);
  }
#pragma unroll
  for (int q = 0; q < 8; q++) v[q] = cmul(v[q], tw[t + 64 * q]);
  dft8(v);
#pragma unroll
  for (int k = 1; k < 8; k++) v[k] = cmul(v[k], tab[K3_TAB_W1 + (k - 1) * 64 + t]);
  // transposes through shared memory with COMPLEX (8-byte) elements: layouts 72 k1 + r and
  // k1 + 8 k2 + 66 r2 are conflict-free for 64-bit accesses (every half-warp touches 16 distinct
  // 8-byte bank pairs) in the writing and in the reading pass
  cpx* T2 = reinterpret_cast<cpx*>(T);
#pragma unroll
  for (int k = 0; k < 8; k++) T2[72 * k + t] = v[k];
  K3_GSYNC(grp);
  const int k1 = t >> 3, r2 = t & 7;
#pragma unroll
  for (int q = 0; q < 8; q++) v[q] = T2[72 * k1 + r2 + 8 * q];
  K3_GSYNC(grp);  // the second transpose reuses T
  dft8(v);
#pragma unroll
  for (int k = 1; k < 8; k++) v[k] = cmul(v[k], tab[K3_TAB_W2 + (k - 1) * 8 + r2]);
#pragma unroll
  for (int k = 0; k < 8; k++) T2[k1 + 8 * k + 66 * r2] = v[k];
  K3_GSYNC(grp);
  // pass 3: this thread owns (k1, k2) = (t & 7, t >> 3): element t + 66 r
#pragma unroll
  for (int r = 0; r < 8; r++) v[r] = T2[t + 66 * r];
  dft8(v);
  // c[p] = T[p] * tw[p] gives D[2p] = Re c[p] and D[M-1-2p] = -Im c[p].  The neighbour of D[2p] in
  // memory, D[2p+1] = D[M-1-2p'] with p' = 511 - p, belongs to the mirrored thread 63 - t (the other lane of
  // k3_remap64's pairing, lane ^ 31) at index 7 - k3: one shuffle fetches it, and the pair leaves as one
  // conflict-free 64-bit store instead of two stride-2 scalar stores.
  float ny[8];
#pragma unroll
  for (int k3 = 0; k3 < 8; k3++) {
    const cpx c = cmul(v[k3], tw[t + 64 * k3]);
    v[k3].x = c.x;
    ny[k3] = -c.y;
  }
#pragma unroll
  for (int k3 = 0; k3 < 8; k3++) {
    const int p = t + 64 * k3;
    const float odd = __shfl_xor_sync(0xffffffffu, ny[7 - k3], 31);
    // 2p < 512: both values lie in the low half, else in the high half
    float* dst = (k3 < 4 ? D.lo : D.hm) + 2 * p;
    *reinterpret_cast<float2*>(dst) = float2{v[k3].x, odd};
  }
}

// Both channels of a stereo long block in ONE pass (K3S_PAIR): the same statements as fft512_to_D on two independent
// register sets and two transpose scratches.  Every twiddle is loaded once for both channels, the group barriers are
// shared (3 per block instead of 6), and twice as many independent loads / multiplies are in flight per thread between
// them -- the workers wait on exactly these chains.
VPZ_DEV void fft512_pair_to_D(const float2* xa, const float2* xb, float* Ta, float* Tb, const K3D& Da, const K3D& Db,
                              const cpx* tab, int t, int grp) {
  const cpx* tw = tab + K3_TAB_TW;
  cpx v[8], u[8];
#pragma unroll
  for (int q = 0; q < 8; q++) {
    v[q].x = xa[q].x;
    v[q].y = __shfl_xor_sync(0xffffffffu, xa[7 - q].y, 31);
    u[q].x = xb[q].x;
    u[q].y = __shfl_xor_sync(0xffffffffu, xb[7 - q].y, 31);
  }
#pragma unroll
  for (int q = 0; q < 8; q++) {
    const cpx w = tw[t + 64 * q];
    v[q] = cmul(v[q], w);
    u[q] = cmul(u[q], w);
  }
  dft8(v);
  dft8(u);
#pragma unroll
  for (int k = 1; k < 8; k++) {
    const cpx w = tab[K3_TAB_W1 + (k - 1) * 64 + t];
    v[k] = cmul(v[k], w);
    u[k] = cmul(u[k], w);
  }
  cpx* A2 = reinterpret_cast<cpx*>(Ta);
  cpx* B2 = reinterpret_cast<cpx*>(Tb);
#pragma unroll
  for (int k = 0; k < 8; k++) {
    A2[72 * k + t] = v[k];
    B2[72 * k + t] = u[k];
  }
  K3_GSYNC(grp);
  const int k1 = t >> 3, r2 = t & 7;
#pragma unroll
  for (int q = 0; q < 8; q++) {
    v[q] = A2[72 * k1 + r2 + 8 * q];
    u[q] = B2[72 * k1 + r2 + 8 * q];
  }
  K3_GSYNC(grp);  // the second transpose reuses the scratches
  dft8(v);
  dft8(u);
#pragma unroll
  for (int k = 1; k < 8; k++) {
    const cpx w = tab[K3_TAB_W2 + (k - 1) * 8 + r2];
    v[k] = cmul(v[k], w);
    u[k] = cmul(u[k], w);
  }
#pragma unroll
  for (int k = 0; k < 8; k++) {
    A2[k1 + 8 * k + 66 * r2] = v[k];
    B2[k1 + 8 * k + 66 * r2] = u[k];
  }
  K3_GSYNC(grp);
#pragma unroll
  for (int r = 0; r < 8; r++) {
    v[r] = A2[t + 66 * r];
    u[r] = B2[t + 66 * r];
  }
  dft8(v);
  dft8(u);
  float ny[8], my[8];
#pragma unroll
  for (int k3 = 0; k3 < 8; k3++) {
    const cpx w = tw[t + 64 * k3];
    const cpx c = cmul(v[k3], w), d = cmul(u[k3], w);
    v[k3].x = c.x;
    ny[k3] = -c.y;
    u[k3].x = d.x;
    my[k3] = -d.y;
  }
#pragma unroll
  for (int k3 = 0; k3 < 8; k3++) {
    const int p = t + 64 * k3;
    const float odd_a = __shfl_xor_sync(0xffffffffu, ny[7 - k3], 31);
    const float odd_b = __shfl_xor_sync(0xffffffffu, my[7 - k3], 31);
    *reinterpret_cast<float2*>((k3 < 4 ? Da.lo : Da.hm) + 2 * p) = float2{v[k3].x, odd_a};
    *reinterpret_cast<float2*>((k3 < 4 ? Db.lo : Db.hm) + 2 * p) = float2{u[k3].x, odd_b};
  }
}

// N = 256: H = 64 = 8*8, threads t < 8 of the group work; M = 128.  tw / w64: shared-memory tables.
// end: bins >= end are an exact +0 and were not written
VPZ_DEV void fft64_to_D(const float* X, float* A, const K3D& D, const cpx* tw, const cpx* w64, int t, bool active, int grp,
                        int end) {
  const int M = 128;
  cpx v[8];
  if (active) {
#pragma unroll
    for (int q = 0; q < 8; q++) {
      int n = t + 8 * q;
      v[q].x = end > 0 ? VPZ_LDG(X + 2 * n) : 0.f;   // a short block is one 128-bin unit: written or not
      v[q].y = end > 0 ? VPZ_LDG(X + M - 1 - 2 * n) : 0.f;
      v[q] = cmul(v[q], tw[n]);
    }
    dft8(v);
#pragma unroll
    for (int k = 1; k < 8; k++) v[k] = cmul(v[k], w64[(t * k) & 63]);
#pragma unroll
    for (int k = 0; k < 8; k++) {
      A[9 * k + t] = v[k].x;
      A[K3_PLANE + 9 * k + t] = v[k].y;
    }
  }
  K3_GSYNC(grp);
  if (active) {
#pragma unroll
    for (int r = 0; r < 8; r++) {
      v[r].x = A[9 * t + r];
      v[r].y = A[K3_PLANE + 9 * t + r];
    }
    dft8(v);
#pragma unroll
    for (int k2 = 0; k2 < 8; k2++) {
      int p = t + 8 * k2;
      cpx c = cmul(v[k2], tw[p]);
      *k3_dp(D, 2 * p) = c.x;
      *k3_dp(D, M - 1 - 2 * p) = -c.y;
    }
  }
  K3_GSYNC(grp);
}

// Any power-of-two N in 64..8192: radix-2 Stockham autosort between two shared buffers of 2*H
// floats (re plane, im plane), 64 threads.
VPZ_DEV void fft_generic_to_D(const float* X, float* A, float* B, const K3D& D, const cpx* tw, const cpx* roots,
                              int log2H, int t, bool active) {
  // every thread of the CTA takes every barrier in here; `active` only predicates the memory traffic
  // (a channel slot without work in this pass idles through the same barriers)
  const int H = 1 << log2H, M = 2 * H;
  if (active)
    for (int n = t; n < H; n += K3_THREADS_PER_CH) {
      cpx z = cmul(cpx{VPZ_LDG(X + 2 * n), VPZ_LDG(X + M - 1 - 2 * n)}, VPZ_LDG(tw + n));
      A[n] = z.x;
      A[H + n] = z.y;
    }
  __syncthreads();
  float* src = A;
  float* dst = B;
  // Stockham autosort, decimation in frequency: stage s works on sub-length ns = H >> s with
  // stride st = 1 << s; natural order in, natural order out after log2H stages.
  for (int s = 0; s < log2H; s++) {
    const int st = 1 << s, m = H >> (s + 1);
    if (active)
      for (int i = t; i < H / 2; i += K3_THREADS_PER_CH) {
        int p = i >> s, q = i & (st - 1);
        cpx w = VPZ_LDG(roots + (p << s));  // exp(-2 pi i p / ns)
        int ia = q + st * p, ib = q + st * (p + m);
        cpx c0 = cpx{src[ia], src[H + ia]};
        cpx c1 = cpx{src[ib], src[H + ib]};
        cpx u = cadd(c0, c1), d = cmul(csub(c0, c1), w);
        int oa = q + st * 2 * p, ob = oa + st;
        dst[oa] = u.x;
        dst[H + oa] = u.y;
        dst[ob] = d.x;
        dst[H + ob] = d.y;
      }
    __syncthreads();
    float* tmp = src;
    src = dst;
    dst = tmp;
  }
  if (active)
    for (int p = t; p < H; p += K3_THREADS_PER_CH) {
      cpx c = cmul(cpx{src[p], src[H + p]}, VPZ_LDG(tw + p));
      *k3_dp(D, 2 * p) = c.x;
      *k3_dp(D, M - 1 - 2 * p) = -c.y;
    }
  __syncthreads();
}

// y[i] of a block with M = N/2 from its D buffer
VPZ_DEV float k3_y(const K3D& D, int M, int i) {
  int h = M >> 1;
  if (i < h) return *k3_dp(D, i + h);
  if (i < M + h) return -*k3_dp(D, M + h - 1 - i);
  return -*k3_dp(D, i - M - h);
}

// ---- window + overlap-add + clip + interleaved store ---------------------------------------------
// y[] of a block is a piecewise mirrored read of its D buffer (k3_y).  The output range [0, count) of
// a packet is cut at the breakpoints of the current block, of the previous block and at the end of
// the overlap; inside one segment every sample uses the same (offset, direction, sign), so the inner
// loop is two shared loads, two window loads, two multiplies and one add per channel -- the rounding
// order of OverlapBuffers (StreamDecoder.cs:786-788): two rounded products, one rounded sum.
struct K3Piece {
  int off, dir;     // D index = off + dir * j
  float sign;
  bool low;         // the piece lies in the low half of D (i >= M), else in the high half
};
VPZ_DEV K3Piece k3_piece(int M, int start, int j) {  // piece of y[start + j]
  const int h = M >> 1, i = start + j;
  K3Piece p;
  if (i < h) { p.off = start + h; p.dir = 1; p.sign = 1.f; }
  else if (i < M + h) { p.off = M + h - 1 - start; p.dir = -1; p.sign = -1.f; }
  else { p.off = start - M - h; p.dir = 1; p.sign = -1.f; }
  p.low = i >= M;
  return p;
}
VPZ_DEV int k3_next_break(int M, int start, int a, int b) {  // first breakpoint of y[start + j] in (a, b)
  const int h = M >> 1;
  int x = h - start;
  if (x > a && x < b) b = x;
  x = M - start;
  if (x > a && x < b) b = x;
  x = M + h - start;
  if (x > a && x < b) b = x;
  return b;
}

// 16-bit PCM as the reference's own tests derive it from the float output (AssetTest.cs:131-132):
// v = (int)(x * 32768f), i.e. the rounded fp32 product truncated toward zero, clamped to [-32768, 32767]
VPZ_DEV int k3_s16(float x) {
#ifndef VPZ_EMU
  int v = __float2int_rz(__fmul_rn(x, 32768.f));
#else
  int v = (int)__fmul_rn(x, 32768.f);
#endif
  return v < -32768 ? -32768 : (v > 32767 ? 32767 : v);
}
// store sample `idx` (element index into the interleaved output) as fp32 or s16
template <bool OUT16>
VPZ_DEV void k3_put(float* outp, size_t idx, float v) {
  if (OUT16) reinterpret_cast<int16_t*>(outp)[idx] = (int16_t)k3_s16(v); else outp[idx] = v;
}
template <bool OUT16>
VPZ_DEV void k3_put2(float* outp, size_t idx, float a, float b) {   // idx even, base suitably aligned
  if (OUT16) {
    const uint32_t w = ((uint32_t)k3_s16(a) & 0xffffu) | ((uint32_t)k3_s16(b) << 16);
    *reinterpret_cast<uint32_t*>(reinterpret_cast<int16_t*>(outp) + idx) = w;
  } else {
    *reinterpret_cast<float2*>(outp + idx) = float2{a, b};
  }
}

template <int NCUR, bool CLIP, bool OUT16 = false>
VPZ_DEV bool k3_emit(const float* Dc_hm, const float* Dc_lo, const float* Dp_lo, int per_ch, int M, int prevM, int ls,
                     int count, int prev_rs, int L, const float* w, float* outp, int C, int tid, int nthreads) {
  bool clipped = false;
  // outp is the float-typed base of the packet's samples; for s16 output the same ELEMENT offsets apply
  // to an int16 buffer, so the byte address is base + 2 * element (the caller passes base accordingly)
  const bool pair_ok = NCUR == 2 && C == 2 && (reinterpret_cast<uintptr_t>(outp) & (OUT16 ? 3u : 7u)) == 0;
  int a = 0;
  while (a < count) {
    int b = k3_next_break(M, ls, a, count);
    const bool ovl = a < L;
    if (ovl) {
      if (L < b) b = L;
      b = k3_next_break(prevM, prev_rs, a, b);
    }
    const K3Piece pc = k3_piece(M, ls, a);
    const K3Piece pp = k3_piece(prevM, prev_rs, a);
    const float* Dc0 = pc.low ? Dc_lo : Dc_hm;
    const float* Dp0 = Dp_lo;  // the overlap lies in the right half of the previous block (host-checked)
    // two samples per thread and step: the loads of both are in flight together
    for (int j = a + tid; j < b; j += 2 * nthreads) {
      const int j2 = j + nthreads;
      const bool two = j2 < b;
      float v[NCUR], u[NCUR];
      float w0 = 0.f, w1 = 0.f, x0 = 0.f, x1 = 0.f;
      if (ovl) {
        // plain loads: the slope lies in global memory (generic kernel) or in the staged tables (256 / 2048 kernel)
        w0 = w[j];
        w1 = w[L - 1 - j];
        if (two) {
          x0 = w[j2];
          x1 = w[L - 1 - j2];
        }
      }
#pragma unroll
      for (int cg = 0; cg < NCUR; cg++) {
        float x = pc.sign * Dc0[cg * per_ch + pc.off + pc.dir * j];
        float y = two ? pc.sign * Dc0[cg * per_ch + pc.off + pc.dir * j2] : 0.f;
        if (ovl) {
          float pv = pp.sign * Dp0[cg * per_ch + pp.off + pp.dir * j];
          float qv = two ? pp.sign * Dp0[cg * per_ch + pp.off + pp.dir * j2] : 0.f;
          x = __fadd_rn(__fmul_rn(x, w0), __fmul_rn(pv, w1));
          y = __fadd_rn(__fmul_rn(y, x0), __fmul_rn(qv, x1));
        }
        if (CLIP) {  // Utils.ClipValue (Utils.cs:44-58): |v| > c -> +-c, anything else (NaN included) unchanged
          const bool px = fabsf(x) > 0.99999994f, py = fabsf(y) > 0.99999994f;
          x = px ? copysignf(0.99999994f, x) : x;
          y = py ? copysignf(0.99999994f, y) : y;
          clipped |= px | py;
        }
        v[cg] = x;
        u[cg] = y;
      }
      const size_t o = (size_t)j * C, o2 = (size_t)j2 * C;
      if (NCUR == 2) {
        if (pair_ok) {
          k3_put2<OUT16>(outp, o, v[0], v[NCUR - 1]);
          if (two) k3_put2<OUT16>(outp, o2, u[0], u[NCUR - 1]);
        } else {
          k3_put<OUT16>(outp, o, v[0]);
          k3_put<OUT16>(outp, o + 1, v[NCUR - 1]);
          if (two) {
            k3_put<OUT16>(outp, o2, u[0]);
            k3_put<OUT16>(outp, o2 + 1, u[NCUR - 1]);
          }
        }
      } else {
        k3_put<OUT16>(outp, o, v[0]);
        if (two) k3_put<OUT16>(outp, o2, u[0]);
      }
    }
    a = b;
  }
  return clipped;
}

// ---- generic kernel: any power-of-two block sizes, any channel count --------------------------------
// Shared memory (floats): descriptors (K3_DESC_FLOATS), then per channel slot the Stockham buffers
// A[2*Hmax] B[2*Hmax] and the D slots Hi[Mmax/2] Lo0[Mmax/2] Lo1[Mmax/2] (+16 to stagger channel bases).
// A CTA = NCB channel slots x 64 threads walks one work item; streams with more channels than slots are
// swept NCB channels at a time.
template <bool OUT16>
VPZ_DEV void k3_run_item(const K3Params& P, const VpzOlaItem& item, float* smem_raw, int NCB) {
  const VpzOlaItem it = item;  // the item lives in global memory: read it once
  const uint32_t* blob = P.setups[it.setup_slot];
  const VpzSetupHdr* Hd = reinterpret_cast<const VpzSetupHdr*>(blob);
  const int C = Hd->channels;
  const int lg0 = Hd->log2_size0, lg1 = Hd->log2_size1;
  const int Mmax = 1 << (lg1 - 1);
  const int PA = 1 << (lg1 - 2);
  const int SCR = 4 * PA;                            // Stockham scratch per channel
  const int HS = Mmax >> 1;                          // one half-slot
  const int per_ch = SCR + 3 * HS + 16;
  const int tid = threadIdx.x;
  const int cgrp = tid / K3_THREADS_PER_CH;          // channel slot inside the CTA
  const int t64 = tid % K3_THREADS_PER_CH;
  const float* slope0 = reinterpret_cast<const float*>(blob + Hd->slope_off[0]);
  const float* slope1 = reinterpret_cast<const float*>(blob + Hd->slope_off[1]);
  const int nthreads = NCB * K3_THREADS_PER_CH;
  VpzPktOla* spk = reinterpret_cast<VpzPktOla*>(smem_raw);                      // [K3_DESC_PKTS] descriptors
  uint32_t* smask = reinterpret_cast<uint32_t*>(smem_raw) + 4 * K3_DESC_PKTS;    // [K3_DESC_PKTS] exec masks
  float* smem = smem_raw + K3_DESC_FLOATS;

  for (int c0 = 0; c0 < C; c0 += NCB) {
    const int ncur = (C - c0) < NCB ? (C - c0) : NCB;  // channels handled in this sweep
    const int ch = c0 + cgrp;
    const bool ch_ok = cgrp < ncur;
    float* base = smem + cgrp * per_ch;
    float* A = base;
    float* B = base + 2 * PA;

    int prevM = 0, prev_rs = 0, prev_re = 0;   // previous packet: M, RightStart, RightEnd
    bool have_prev = false;
    int parity = 0;
    const int first = (int)it.first_pkt - (it.has_pre ? 1 : 0);
    const int total = (int)it.n_pkts + (it.has_pre ? 1 : 0);
    for (int pb = 0; pb < total; pb += K3_DESC_PKTS) {
      const int nb = (total - pb) < K3_DESC_PKTS ? (total - pb) : K3_DESC_PKTS;
      // descriptors + exec masks of the next nb packets: one parallel fetch
      __syncthreads();
      if (tid < nb) {
        spk[tid] = P.pkts[first + pb + tid];
        smask[tid] = P.res ? P.res[first + pb + tid].exec_mask : 0xffu;
      }
      __syncthreads();
      for (int pw = 0; pw < nb; pw++, parity ^= 1) {
        const int pi = pb + pw;
        const uint32_t gp = (uint32_t)(first + pi);
        const VpzPktOla pk = spk[pw];
        const uint32_t mask = smask[pw];
        const bool is_long = pk.flags & VPZ_OLA_LONG;
        const int lgN = is_long ? lg1 : lg0;
        const int M = 1 << (lgN - 1);
        const bool emit = !(pi == 0 && it.has_pre) && !(pk.flags & VPZ_OLA_NOOUT) && have_prev;

        // ---- transform every channel of this sweep into its D buffer -------------------------
        const bool exec = ch_ok && ((mask >> ch) & 1u);
        const float* X = P.spec + pk.spec_off + (size_t)ch * M;
        const cpx* tw = reinterpret_cast<const cpx*>(blob + Hd->tw_off[is_long ? 1 : 0]);
        const cpx* roots = reinterpret_cast<const cpx*>(blob + Hd->fft_off[is_long ? 1 : 0]);
        const int h = M >> 1;
        K3D D;
        D.h = h;
        D.hm = base + SCR - h;
        D.lo = base + SCR + HS + parity * HS;
        // a channel without floor energy outputs zeros (Mapping.cs:185-194) but still takes part in the
        // overlap-add: its D buffer is cleared instead of transformed
        if (ch_ok && !exec)
          for (int i = t64; i < M; i += K3_THREADS_PER_CH) *k3_dp(D, i) = 0.f;
        fft_generic_to_D(X, A, B, D, tw, roots, lgN - 2, t64, exec);
        // D buffers of the sweep are complete here (the transform ends with a barrier)

        if (P.dbg_imdct && ch_ok) {
          float* dy = P.dbg_imdct + 2 * (size_t)pk.spec_off + (size_t)ch * 2 * M;
          for (int i = t64; i < 2 * M; i += K3_THREADS_PER_CH) dy[i] = k3_y(D, M, i);
        }

        if (emit) {
          const int ls = pk.left_start;
          const int count = (int)pk.right_start - ls;
          const int L = prev_re - prev_rs;             // StreamDecoder.cs:654
          const float* w = (pk.flags & VPZ_OLA_LEFT1) ? slope1 : slope0;
          // element offset of the packet's first sample; an s16 element is 2 bytes (k3_put)
          const size_t eoff = (size_t)it.out_base + (size_t)pk.out_off * C + c0;
          float* outp = OUT16 ? reinterpret_cast<float*>(reinterpret_cast<int16_t*>(P.pcm) + eoff) : P.pcm + eoff;
          const float* Dc_hm = smem + SCR - h;                       // channel slot 0; + cg * per_ch for the others
          const float* Dc_lo = smem + SCR + HS + parity * HS;
          const float* Dp_lo = smem + SCR + HS + (parity ^ 1) * HS;
          bool clipped;
          if (ncur == 2)
            clipped = P.clip ? k3_emit<2, true, OUT16>(Dc_hm, Dc_lo, Dp_lo, per_ch, M, prevM, ls, count, prev_rs, L, w, outp, C, tid, nthreads)
                             : k3_emit<2, false, OUT16>(Dc_hm, Dc_lo, Dp_lo, per_ch, M, prevM, ls, count, prev_rs, L, w, outp, C, tid, nthreads);
          else
            clipped = P.clip ? k3_emit<1, true, OUT16>(Dc_hm, Dc_lo, Dp_lo, per_ch, M, prevM, ls, count, prev_rs, L, w, outp, C, tid, nthreads)
                             : k3_emit<1, false, OUT16>(Dc_hm, Dc_lo, Dp_lo, per_ch, M, prevM, ls, count, prev_rs, L, w, outp, C, tid, nthreads);
          // per packet: 0 when any sample was clamped (HasClipped), else stays 0xffffffff
          if (P.clip_first && clipped) atomicMin(P.clip_first + gp, 0u);
        }
        prevM = M;
        prev_rs = pk.right_start;
        prev_re = pk.right_end;
        have_prev = true;
        // the output loop reads the previous D slot, which the next packet's zero fill may overwrite
        // before its first barrier
        __syncthreads();
      }
    }
    __syncthreads();
  }
}

// CTA main loop: work items are handed out by a global counter.
template <bool OUT16>
VPZ_DEV void k3_cta_loop(const K3Params& P, float* smem_raw, int ncb) {
  uint32_t* s_next = reinterpret_cast<uint32_t*>(smem_raw) + (K3_DESC_FLOATS - 1);
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) *s_next = atomicAdd(P.counter, 1u);
    __syncthreads();
    const uint32_t idx = *s_next;
    if (idx >= P.n_items) break;
    k3_run_item<OUT16>(P, P.items[idx], smem_raw, ncb);
  }
}
