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
// Only D (M floats per channel) is kept; the right half of a block is never materialised: the
// next packet reads it straight out of the previous packet's D buffer (ping-pong).
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

typedef float2 cpx;
VPZ_DEV cpx cmul(cpx a, cpx b) {
  cpx r;
  r.x = a.x * b.x - a.y * b.y;
  r.y = a.x * b.y + a.y * b.x;
  return r;
}
VPZ_DEV cpx cadd(cpx a, cpx b) { return cpx{a.x + b.x, a.y + b.y}; }
VPZ_DEV cpx csub(cpx a, cpx b) { return cpx{a.x - b.x, a.y - b.y}; }
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
  cpx d1 = cpx{h * (c1.x + c1.y), h * (c1.y - c1.x)};
  cpx d3 = cpx{h * (c3.y - c3.x), -h * (c3.x + c3.y)};
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

// transposes: plane stride 576 floats; conflict-free for both the writing and the reading pass
VPZ_DEV int idx1(int k1, int r) { return 72 * k1 + r; }                   // r in [0,64)
VPZ_DEV int idx2(int k1, int k2, int r2) { return 72 * k1 + 9 * r2 + k2; }
#define K3_PLANE 576

// N = 2048: H = 512 = 8*8*8, 64 threads.  X: M = 1024 floats in global.  Writes D[0..1024) (smem).
VPZ_DEV void fft512_to_D(const float* X, float* A, float* B, float* D, const cpx* tw, const cpx* w512,
                         int t, int lane) {
  const int M = 1024;
  cpx v[8];
  float other[8];
#pragma unroll
  for (int q = 0; q < 8; q++) {
    const float2 f = VPZ_LDG(reinterpret_cast<const float2*>(X) + (t + 64 * q));
    v[q].x = f.x;       // X[2n]
    other[q] = f.y;     // X[2n+1] = X[M-1-2n'] of the mirrored element n' = 511-n
  }
#pragma unroll
  for (int q = 0; q < 8; q++) v[q].y = __shfl_xor_sync(0xffffffffu, other[7 - q], 31);
#pragma unroll
  for (int q = 0; q < 8; q++) v[q] = cmul(v[q], VPZ_LDG(tw + t + 64 * q));
  dft8(v);
#pragma unroll
  for (int k = 1; k < 8; k++) v[k] = cmul(v[k], VPZ_LDG(w512 + ((t * k) & 511)));
#pragma unroll
  for (int k = 0; k < 8; k++) {
    A[idx1(k, t)] = v[k].x;
    A[K3_PLANE + idx1(k, t)] = v[k].y;
  }
  __syncthreads();
  const int k1 = t >> 3, r2 = t & 7;
#pragma unroll
  for (int q = 0; q < 8; q++) {
    v[q].x = A[idx1(k1, r2 + 8 * q)];
    v[q].y = A[K3_PLANE + idx1(k1, r2 + 8 * q)];
  }
  dft8(v);
#pragma unroll
  for (int k = 1; k < 8; k++) v[k] = cmul(v[k], VPZ_LDG(w512 + ((r2 * k) << 3)));  // W_64^{r2 k} = W_512^{8 r2 k}
#pragma unroll
  for (int k = 0; k < 8; k++) {
    B[idx2(k1, k, r2)] = v[k].x;
    B[K3_PLANE + idx2(k1, k, r2)] = v[k].y;
  }
  __syncthreads();
  const int kk1 = t >> 3, kk2 = t & 7;  // this thread now owns (k1, k2)
#pragma unroll
  for (int r = 0; r < 8; r++) {
    v[r].x = B[idx2(kk1, kk2, r)];
    v[r].y = B[K3_PLANE + idx2(kk1, kk2, r)];
  }
  dft8(v);
#pragma unroll
  for (int k3 = 0; k3 < 8; k3++) {
    int p = kk1 + 8 * kk2 + 64 * k3;
    cpx c = cmul(v[k3], VPZ_LDG(tw + p));
    D[2 * p] = c.x;
    D[M - 1 - 2 * p] = -c.y;
  }
  (void)lane;
}

// N = 256: H = 64 = 8*8, threads t < 8 of the channel group work; M = 128.
VPZ_DEV void fft64_to_D(const float* X, float* A, float* D, const cpx* tw, const cpx* w64, int t, bool active) {
  const int M = 128;
  cpx v[8];
  if (active) {
#pragma unroll
    for (int q = 0; q < 8; q++) {
      int n = t + 8 * q;
      v[q].x = VPZ_LDG(X + 2 * n);
      v[q].y = VPZ_LDG(X + M - 1 - 2 * n);
      v[q] = cmul(v[q], VPZ_LDG(tw + n));
    }
    dft8(v);
#pragma unroll
    for (int k = 1; k < 8; k++) v[k] = cmul(v[k], VPZ_LDG(w64 + ((t * k) & 63)));
#pragma unroll
    for (int k = 0; k < 8; k++) {
      A[9 * k + t] = v[k].x;
      A[K3_PLANE + 9 * k + t] = v[k].y;
    }
  }
  __syncthreads();
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
      cpx c = cmul(v[k2], VPZ_LDG(tw + p));
      D[2 * p] = c.x;
      D[M - 1 - 2 * p] = -c.y;
    }
  }
  __syncthreads();
}

// Any power-of-two N in 64..8192: radix-2 Stockham autosort between two shared buffers of 2*H
// floats (re plane, im plane), 64 threads.
VPZ_DEV void fft_generic_to_D(const float* X, float* A, float* B, float* D, const cpx* tw, const cpx* roots,
                              int log2H, int t) {
  const int H = 1 << log2H, M = 2 * H;
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
  for (int p = t; p < H; p += K3_THREADS_PER_CH) {
    cpx c = cmul(cpx{src[p], src[H + p]}, VPZ_LDG(tw + p));
    D[2 * p] = c.x;
    D[M - 1 - 2 * p] = -c.y;
  }
  __syncthreads();
}

// y[i] of a block with M = N/2 from its D buffer
VPZ_DEV float k3_y(const float* D, int M, int i) {
  int h = M >> 1;
  if (i < h) return D[i + h];
  if (i < M + h) return -D[M + h - 1 - i];
  return -D[i - M - h];
}

// Shared memory per channel (floats): A[2*PA] B[2*PA] D0[Mmax] D1[Mmax] where PA = plane size.
// FAST: block sizes 256/2048.  Otherwise generic (PA = Hmax).
template <bool FAST>
VPZ_DEV void k3_run_item(const K3Params& P, const VpzOlaItem& it, float* smem, int NCB) {
  const uint32_t* blob = P.setups[it.setup_slot];
  const VpzSetupHdr* Hd = reinterpret_cast<const VpzSetupHdr*>(blob);
  const int C = Hd->channels;
  const int lg0 = Hd->log2_size0, lg1 = Hd->log2_size1;
  const int Mmax = 1 << (lg1 - 1);
  const int PA = FAST ? K3_PLANE : (1 << (lg1 - 2));
  const int per_ch = 4 * PA + 2 * Mmax + 16;  // +16: stagger channel bases across banks
  const int tid = threadIdx.x;
  const int cgrp = tid / K3_THREADS_PER_CH;          // channel slot inside the CTA
  const int t64 = tid % K3_THREADS_PER_CH;
  const int t = FAST ? k3_remap64(t64) : t64;
  const float* slope0 = reinterpret_cast<const float*>(blob + Hd->slope_off[0]);
  const float* slope1 = reinterpret_cast<const float*>(blob + Hd->slope_off[1]);
  const int nthreads = NCB * K3_THREADS_PER_CH;

  for (int c0 = 0; c0 < C; c0 += NCB) {
    const int ncur = (C - c0) < NCB ? (C - c0) : NCB;  // channels handled in this sweep
    const int ch = c0 + cgrp;
    const bool ch_ok = cgrp < ncur;
    float* base = smem + cgrp * per_ch;
    float* A = base;
    float* B = base + 2 * PA;
    float* Dbuf[2] = {base + 4 * PA, base + 4 * PA + Mmax};

    int prevM = 0, prev_rs = 0, prev_re = 0;   // previous packet: M, RightStart, RightEnd
    bool have_prev = false;
    uint32_t zero_bits = 0;                    // bit (cg*2 + parity): that D buffer is all zero
    int parity = 0;
    const int first = (int)it.first_pkt - (it.has_pre ? 1 : 0);
    const int total = (int)it.n_pkts + (it.has_pre ? 1 : 0);
    for (int pi = 0; pi < total; pi++, parity ^= 1) {
      const uint32_t gp = (uint32_t)(first + pi);
      const VpzPktOla pk = P.pkts[gp];
      const bool is_long = pk.flags & VPZ_OLA_LONG;
      const int lgN = is_long ? lg1 : lg0;
      const int M = 1 << (lgN - 1);
      const uint32_t mask = P.res ? P.res[gp].exec_mask : 0xffu;
      const bool emit = !(pi == 0 && it.has_pre) && !(pk.flags & VPZ_OLA_NOOUT) && have_prev;

      // ---- transform every channel of this sweep into its D buffer -------------------------
      const bool exec = ch_ok && ((mask >> ch) & 1u);
      const float* X = P.spec + pk.spec_off + (size_t)ch * M;
      const cpx* tw = reinterpret_cast<const cpx*>(blob + Hd->tw_off[is_long ? 1 : 0]);
      const cpx* roots = reinterpret_cast<const cpx*>(blob + Hd->fft_off[is_long ? 1 : 0]);
      float* D = Dbuf[parity];
      if (FAST) {
        if (is_long) {
          if (exec) {
            fft512_to_D(X, A, B, D, tw, roots, t, tid & 31);
          } else {
            __syncthreads();
            __syncthreads();
          }
          __syncthreads();
        } else {
          fft64_to_D(exec ? X : P.spec, A, D, tw, roots, t64, exec && t64 < 8);
        }
      } else {
        // every thread must take the barriers inside; silent channels transform a dummy but skip stores
        if (exec) {
          fft_generic_to_D(X, A, B, D, tw, roots, lgN - 2, t64);
        } else {
          for (int s = 0; s < lgN - 2 + 2; s++) __syncthreads();
        }
      }
      // D buffers of the sweep are complete here (each path ends with a barrier)

      if (P.dbg_imdct && ch_ok) {
        float* dy = P.dbg_imdct + 2 * (size_t)pk.spec_off + (size_t)ch * 2 * M;
        for (int i = t64; i < 2 * M; i += K3_THREADS_PER_CH) dy[i] = exec ? k3_y(D, M, i) : 0.f;
      }

      // ---- window + overlap-add + clip + interleaved store ---------------------------------------
      // Every warp takes 32 consecutive samples per step.  y[] is a piecewise mirrored read of D
      // (k3_y) whose breakpoints, like LeftStart / RightStart of untrimmed packets, are multiples
      // of 64, so one warp step never straddles a breakpoint: the piece (base, direction, sign) is
      // picked once per step with warp-uniform branches and each lane does two shared loads, two
      // window loads, two multiplies and one add per channel -- same rounding order as
      // OverlapBuffers (StreamDecoder.cs:786-788): two rounded products, one rounded sum.
      if (emit) {
        const int ls = pk.left_start, rs = pk.right_start;
        const int count = rs - ls;
        const int L = prev_re - prev_rs;             // StreamDecoder.cs:654
        const float* w = (pk.flags & VPZ_OLA_LEFT1) ? slope1 : slope0;
        float* outp = P.pcm + it.out_base + (size_t)pk.out_off * C + c0;
        const int lane = tid & 31, wid = tid >> 5, nw = nthreads >> 5;
        const int h = M >> 1, ph = prevM >> 1;
        const bool pair_ok = (reinterpret_cast<uintptr_t>(outp) & 7u) == 0;  // runs are packed back to back
        const float* Dbase = smem + 4 * PA;
        uint32_t clipped_at = 0xffffffffu;
        uint32_t cur_on = 0, prev_on = 0;
        for (int cg = 0; cg < ncur; cg++) {
          if ((mask >> (c0 + cg)) & 1u) cur_on |= 1u << cg;
          if (!((zero_bits >> (cg * 2 + (parity ^ 1))) & 1u)) prev_on |= 1u << cg;
        }
        for (int j0 = wid * 32; j0 < count; j0 += nw * 32) {
          const int j = j0 + lane;
          // current block, index i = ls + j
          const int i0 = ls + j0;
          int cidx;
          float csign;
          if (i0 < h) { cidx = i0 + h + lane; csign = 1.f; }
          else if (i0 < M + h) { cidx = M + h - 1 - i0 - lane; csign = -1.f; }
          else { cidx = i0 - M - h + lane; csign = -1.f; }
          // previous block, index prev_rs + j (only inside the overlap)
          const int p0 = prev_rs + j0;
          int pidx;
          float psign;
          if (p0 < ph) { pidx = p0 + ph + lane; psign = 1.f; }
          else if (p0 < prevM + ph) { pidx = prevM + ph - 1 - p0 - lane; psign = -1.f; }
          else { pidx = p0 - prevM - ph + lane; psign = -1.f; }
          const bool in_ovl = j < L;
          float w0 = 1.f, w1 = 0.f;
          if (in_ovl) {
            w0 = VPZ_LDG(w + j);
            w1 = VPZ_LDG(w + (L - 1 - j));
          }
          if (j < count) {
            // breakpoints are multiples of 64 except after an end-of-stream trim, where the tail
            // of the last step may run past a piece; clamp keeps the (discarded) reads in bounds
            cidx = cidx < 0 ? 0 : (cidx >= M ? M - 1 : cidx);
            pidx = pidx < 0 ? 0 : (pidx >= prevM ? prevM - 1 : pidx);
            float v[2];
#pragma unroll
            for (int cg = 0; cg < 2; cg++) {
              if (cg < ncur) {
                const float* Dc = Dbase + cg * per_ch + parity * Mmax;
                const float* Dp = Dbase + cg * per_ch + (parity ^ 1) * Mmax;
                float x = ((cur_on >> cg) & 1u) ? csign * Dc[cidx] : 0.f;
                if (in_ovl) {
                  float pv = ((prev_on >> cg) & 1u) ? psign * Dp[pidx] : 0.f;
                  x = __fadd_rn(__fmul_rn(x, w0), __fmul_rn(pv, w1));
                }
                if (P.clip) {  // Utils.ClipValue (Utils.cs:44-58)
                  if (x > 0.99999994f) {
                    x = 0.99999994f;
                    clipped_at = clipped_at < (uint32_t)j ? clipped_at : (uint32_t)j;
                  } else if (x < -0.99999994f) {
                    x = -0.99999994f;
                    clipped_at = clipped_at < (uint32_t)j ? clipped_at : (uint32_t)j;
                  }
                }
                v[cg] = x;
              }
            }
            float* o = outp + (size_t)j * C;
            if (C == 2 && pair_ok) {
              *reinterpret_cast<float2*>(o) = float2{v[0], v[1]};
            } else {
              o[0] = v[0];
              if (ncur > 1) o[1] = v[1];
            }
          }
        }
        if (P.clip_first && clipped_at != 0xffffffffu) atomicMin(P.clip_first + gp, clipped_at);
      }
      // remember this packet as "previous"; its D buffer stays untouched until the packet after next
      for (int cg = 0; cg < ncur; cg++) {
        uint32_t bit = 1u << (cg * 2 + parity);
        if ((mask >> (c0 + cg)) & 1u) zero_bits &= ~bit; else zero_bits |= bit;
      }
      prevM = M;
      prev_rs = pk.right_start;
      prev_re = pk.right_end;
      have_prev = true;
      // The output loop above reads Dbuf[parity^1]; the next packet writes Dbuf[parity^1] only after
      // its own transform barriers, and A/B are rewritten only after a barrier too, except the very
      // first stores of the next transform (A) which race with nothing read here.  One barrier keeps
      // the D ping-pong safe when the next transform is barrier-free up to its D stores (N=256 path
      // writes A before its first barrier, D after it).
      __syncthreads();
    }
    __syncthreads();
  }
}
