/*
 * vo_bench.c -- TEST / BENCH INFRASTRUCTURE ONLY (see vorbis_oracle.h).
 *
 * Times the CPU restatement of the reference decode path the way the reference's own TestApp
 * drives it (TestApp/Program.cs:42,155: ReadSamples into a 48,000-float buffer until it returns
 * 0), one stream per thread, so bench.py can report a host-core baseline beside the GPU number.
 */
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "vorbis_oracle.h"

typedef struct {
  const uint8_t* const* datas;
  const size_t* lens;
  int nfiles;
  int njobs;
  int* next;                 /* shared job cursor: threads take the next undecoded stream (the files differ
                                100x in cost, a static split would time the unluckiest thread) */
  int64_t samples;           /* channel-samples produced */
  int errors;
  /* excerpt jobs (vo_bench_excerpts) */
  const uint32_t* file_of;
  const int64_t* start;
  const int32_t* count;
} bench_arg;

static void* bench_thread(void* p) {
  bench_arg* a = (bench_arg*)p;
  float* buf = (float*)malloc(48000 * sizeof(float));
  for (;;) {
    int j = __atomic_fetch_add(a->next, 1, __ATOMIC_RELAXED);
    if (j >= a->njobs) break;
    int f = j % a->nfiles;
    int err = 0;
    vo_stream* s = vo_open(a->datas[f], a->lens[f], &err);
    if (!s) {
      a->errors++;
      continue;
    }
    int ch = vo_channels(s);
    int n = 48000 - 48000 % ch;
    for (;;) {
      int got = vo_read(s, buf, n);
      if (got <= 0) break; /* 0 = end of stream; <0 = the reference faults here */
      a->samples += (int64_t)got * ch;
    }
    vo_close(s);
  }
  free(buf);
  return NULL;
}

/* Decodes `njobs` whole streams (job j = file j % nfiles) on `nthreads` threads.  Returns the
 * channel-samples decoded; *seconds receives the wall-clock time. */
int64_t vo_bench_decode(const uint8_t* const* datas, const size_t* lens, int nfiles, int njobs, int nthreads,
                        double* seconds) {
  if (nthreads < 1) nthreads = 1;
  pthread_t* th = (pthread_t*)calloc((size_t)nthreads, sizeof(pthread_t));
  bench_arg* args = (bench_arg*)calloc((size_t)nthreads, sizeof(bench_arg));
  struct timespec t0, t1;
  int next = 0;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (int t = 0; t < nthreads; t++) {
    args[t].datas = datas;
    args[t].lens = lens;
    args[t].nfiles = nfiles;
    args[t].njobs = njobs;
    args[t].next = &next;
    pthread_create(&th[t], NULL, bench_thread, &args[t]);
  }
  int64_t total = 0;
  for (int t = 0; t < nthreads; t++) {
    pthread_join(th[t], NULL);
    total += args[t].samples;
  }
  clock_gettime(CLOCK_MONOTONIC, &t1);
  if (seconds) *seconds = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
  free(th);
  free(args);
  return total;
}

/* ---- random access (BASELINE config 5): SeekTo(start) + read `count` samples per channel --------------
 * Every thread keeps ONE open reader per file and seeks it from excerpt to excerpt (what a host program
 * doing random access with the reference would do: VorbisReader.SeekTo, StreamDecoder.cs:817-880, then
 * ReadSamples); excerpts are handed out by a shared cursor. */
static void* excerpt_thread(void* p) {
  bench_arg* a = (bench_arg*)p;
  vo_stream** rd = (vo_stream**)calloc((size_t)a->nfiles, sizeof(vo_stream*));
  float* buf = (float*)malloc(8192 * 8 * sizeof(float));
  for (;;) {
    int j = __atomic_fetch_add(a->next, 1, __ATOMIC_RELAXED);
    if (j >= a->njobs) break;
    int f = (int)a->file_of[j];
    if (!rd[f]) {
      int err = 0;
      rd[f] = vo_open(a->datas[f], a->lens[f], &err);
      if (!rd[f]) {
        a->errors++;
        continue;
      }
    }
    vo_stream* s = rd[f];
    if (vo_seek(s, a->start[j]) != VO_OK) {
      a->errors++;
      continue;
    }
    int ch = vo_channels(s);
    int left = a->count[j];
    while (left > 0) {
      int want = left < 8192 ? left : 8192;
      int got = vo_read(s, buf, want * ch);
      if (got <= 0) break;
      a->samples += (int64_t)got * ch;
      left -= got;
    }
  }
  for (int f = 0; f < a->nfiles; f++)
    if (rd[f]) vo_close(rd[f]);
  free(rd);
  free(buf);
  return NULL;
}

int64_t vo_bench_excerpts(const uint8_t* const* datas, const size_t* lens, int nfiles, int n, const uint32_t* file_of,
                          const int64_t* start, const int32_t* count, int nthreads, double* seconds) {
  if (nthreads < 1) nthreads = 1;
  pthread_t* th = (pthread_t*)calloc((size_t)nthreads, sizeof(pthread_t));
  bench_arg* args = (bench_arg*)calloc((size_t)nthreads, sizeof(bench_arg));
  struct timespec t0, t1;
  int next = 0;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (int t = 0; t < nthreads; t++) {
    args[t].datas = datas;
    args[t].lens = lens;
    args[t].nfiles = nfiles;
    args[t].njobs = n;
    args[t].next = &next;
    args[t].file_of = file_of;
    args[t].start = start;
    args[t].count = count;
    pthread_create(&th[t], NULL, excerpt_thread, &args[t]);
  }
  int64_t total = 0;
  for (int t = 0; t < nthreads; t++) {
    pthread_join(th[t], NULL);
    total += args[t].samples;
  }
  clock_gettime(CLOCK_MONOTONIC, &t1);
  if (seconds) *seconds = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
  free(th);
  free(args);
  return total;
}


/* ---- kernel-only baseline (BASELINE config 3): Mdct.Reverse + OverlapBuffers + StoreInterleaved ---------------- */
typedef struct {
  const float* spectra;
  const uint8_t* flags;
  const int64_t* stream_off;   /* float offset of every stream's first spectrum */
  int n_streams, n_blocks, channels, size0, size1;
  int* next;
  int64_t samples;
  double sum;
} ola_arg;

static void* ola_thread(void* p) {
  ola_arg* a = (ola_arg*)p;
  const int C = a->channels, size0 = a->size0, size1 = a->size1;
  float* bufs[2];
  bufs[0] = (float*)malloc(sizeof(float) * (size_t)size1 * (size_t)C);
  bufs[1] = (float*)malloc(sizeof(float) * (size_t)size1 * (size_t)C);
  float* scratch = (float*)malloc(sizeof(float) * (size_t)size1);
  float* out = (float*)malloc(sizeof(float) * (size_t)size1 * (size_t)C);
  float* slope[2];
  slope[0] = (float*)malloc(sizeof(float) * (size_t)(size0 / 2));
  slope[1] = (float*)malloc(sizeof(float) * (size_t)(size1 / 2));
  vo_window_slope(slope[0], size0 / 2);
  vo_window_slope(slope[1], size1 / 2);
  for (;;) {
    const int s = __sync_fetch_and_add(a->next, 1);
    if (s >= a->n_streams) break;
    const uint8_t* fl = a->flags + (size_t)s * (size_t)a->n_blocks;
    const float* X = a->spectra + a->stream_off[s];
    float *prev = NULL, *cur = bufs[0];
    int prev_end = 0, prev_stop = 0;
    for (int i = 0; i < a->n_blocks; i++) {
      const int lb = fl[i] & 1;
      const int pf = i == 0 ? 1 : (fl[i - 1] & 1), nf = i + 1 == a->n_blocks ? 1 : (fl[i + 1] & 1);
      const int n = lb ? size1 : size0;
      int32_t info[6];
      vo_packet_info(size0, size1, lb, pf, nf, info);
      const int left_use1 = info[1], left_start = info[2], right_start = info[4], right_end = info[5];
      for (int ch = 0; ch < C; ch++) {   /* Mdct.Reverse works in place on a buffer of n floats whose first half is the spectrum */
        float* b = cur + (size_t)size1 * ch;
        memcpy(b, X, sizeof(float) * (size_t)(n / 2));
        X += n / 2;
        vo_imdct2(b, n, scratch);
      }
      int prev_start;
      if (prev) {   /* StreamDecoder.OverlapBuffers (StreamDecoder.cs:764-791) */
        const int L = prev_stop - prev_end;
        const float* w = slope[left_use1 ? 1 : 0];
        for (int ch = 0; ch < C; ch++) {
          const float* pv = prev + (size_t)size1 * ch + prev_end;
          float* chan = cur + (size_t)size1 * ch + left_start;
          for (int j = 0; j < L; j++) chan[j] = chan[j] * w[j] + pv[j] * w[L - 1 - j];
        }
        prev_start = left_start;
      } else {
        prev_start = right_start;
      }
      prev_end = right_start;
      prev_stop = right_end;
      float* t = prev ? prev : bufs[1];
      prev = cur;
      cur = t;
      /* Read: [prev_start, prev_end) of every channel, clipped, interleaved (StoreInterleaved) */
      const int count = prev_end - prev_start;
      for (int ch = 0; ch < C; ch++) {
        const float* src = prev + (size_t)size1 * ch + prev_start;
        for (int j = 0; j < count; j++) {
          float v = src[j];
          if (v > 0.99999994f) v = 0.99999994f;
          if (v < -0.99999994f) v = -0.99999994f;
          out[(size_t)j * C + ch] = v;
        }
      }
      {
        float acc = 0.f;   /* consumed, so the store loop cannot be elided; also what the test compares */
        for (int j = 0; j < count * C; j++) acc += out[j];
        a->sum += acc;
      }
      a->samples += (int64_t)count * C;
    }
  }
  free(bufs[0]);
  free(bufs[1]);
  free(scratch);
  free(out);
  free(slope[0]);
  free(slope[1]);
  return NULL;
}

int64_t vo_bench_imdct_ola(const float* spectra, const uint8_t* flags, int n_streams, int n_blocks, int channels,
                           int size0, int size1, int nthreads, double* seconds, double* checksum) {
  if (nthreads < 1) nthreads = 1;
  int64_t* off = (int64_t*)calloc((size_t)n_streams + 1, sizeof(int64_t));
  for (int s = 0; s < n_streams; s++) {
    int64_t fl = 0;
    for (int i = 0; i < n_blocks; i++) fl += ((flags[(size_t)s * n_blocks + i] & 1) ? size1 : size0) / 2 * channels;
    off[s + 1] = off[s] + fl;
  }
  pthread_t* th = (pthread_t*)calloc((size_t)nthreads, sizeof(pthread_t));
  ola_arg* args = (ola_arg*)calloc((size_t)nthreads, sizeof(ola_arg));
  struct timespec t0, t1;
  int next = 0;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (int t = 0; t < nthreads; t++) {
    args[t].spectra = spectra;
    args[t].flags = flags;
    args[t].stream_off = off;
    args[t].n_streams = n_streams;
    args[t].n_blocks = n_blocks;
    args[t].channels = channels;
    args[t].size0 = size0;
    args[t].size1 = size1;
    args[t].next = &next;
    pthread_create(&th[t], NULL, ola_thread, &args[t]);
  }
  int64_t total = 0;
  double sum = 0;
  for (int t = 0; t < nthreads; t++) {
    pthread_join(th[t], NULL);
    total += args[t].samples;
    sum += args[t].sum;
  }
  clock_gettime(CLOCK_MONOTONIC, &t1);
  if (seconds) *seconds = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
  if (checksum) *checksum = sum;
  free(th);
  free(args);
  free(off);
  return total;
}
