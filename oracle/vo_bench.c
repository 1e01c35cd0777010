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
  int first, stride, njobs;  /* this thread decodes jobs first, first+stride, ... < njobs */
  int64_t samples;           /* channel-samples produced */
  int errors;
} bench_arg;

static void* bench_thread(void* p) {
  bench_arg* a = (bench_arg*)p;
  float* buf = (float*)malloc(48000 * sizeof(float));
  for (int j = a->first; j < a->njobs; j += a->stride) {
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
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (int t = 0; t < nthreads; t++) {
    args[t].datas = datas;
    args[t].lens = lens;
    args[t].nfiles = nfiles;
    args[t].first = t;
    args[t].stride = nthreads;
    args[t].njobs = njobs;
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
