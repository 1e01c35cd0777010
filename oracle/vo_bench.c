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
