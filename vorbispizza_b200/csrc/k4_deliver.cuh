// k4_deliver.cuh -- K4: hands the samples of a batch of random-access excerpts out on the device.
//
// vpz_decode_excerpts decodes the windows of many excerpts in one batch; what the caller gets of a window is
// what StreamDecoder.SeekTo + Read deliver: the target packet from the roll-forward point on, then packet after
// packet until `count` samples (StreamDecoder.cs:851-879, 418-498).  Which float ranges those are depends only
// on the packets' sample counts, so the host replays the reader's bookkeeping WITHOUT touching a sample and
// records one copy segment per (excerpt, packet); this kernel moves the segments from the batch PCM into the
// caller's layout (and clips like Utils.ClipValue, Utils.cs:44-58, when ClipSamples is set), and one
// device->host copy of the dense result follows -- instead of copying every window to the host and gathering there.
#pragma once
#include "k1_params.h"

#ifndef VPZ_EMU
#define K4_DEV __device__ __forceinline__
#else
#define K4_DEV inline
#endif

#define K4_THREADS 256

// one warp per segment, grid-stride over the segments
K4_DEV void k4_cta(const K4Params& P) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t warps_per_cta = blockDim.x >> 5;
  const uint32_t nwarps = gridDim.x * warps_per_cta;
  for (uint32_t s = blockIdx.x * warps_per_cta + (threadIdx.x >> 5); s < P.n_segs; s += nwarps) {
    const VpzCopySeg g = P.segs[s];
    const float* src = P.pcm + g.src;
    float* dst = P.out + g.dst;
    for (uint32_t i = lane; i < g.n; i += 32u) {
      float v = src[i];
      if (P.clip) {   // ClipValue: NaN passes, like the reference's two comparisons
        if (v > 0.99999994f) v = 0.99999994f;
        if (v < -0.99999994f) v = -0.99999994f;
      }
      dst[i] = v;
    }
  }
}
