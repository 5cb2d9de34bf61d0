/* oracle/orc_mt.c - multi-thread harness around the CPU oracle (orc_scan).  TEST / BASELINE INFRASTRUCTURE
 * ONLY (see orc.h): this is what `bench.py --impl reference` and the `cpu_baseline` leg time on the host
 * cores of the GPU box.  Files are independent units, so the harness is a static
 * partition of the files by BYTES over a pool of persistent POSIX threads:
 *
 *   - thread count = number of CPUs in the calling thread's affinity mask (sched_getaffinity: honours
 *     cgroup cpusets and taskset), or the count the caller asks for;
 *   - the pool lives across calls (no thread spawn and no allocation inside the timed region);
 *   - every thread runs the plain single-thread orc_scan() over its slice with its own count tables,
 *     the caller's thread sums the tables afterwards.
 *
 * Nothing here restates reference code (the package ships none: SURVEY.md section 0). */
#define _GNU_SOURCE
#include "orc.h"
#include <pthread.h>
#include <sched.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  const uint8_t* arena; const int32_t* off; const int32_t* len; const uint8_t* ext; const uint16_t* grp;
  int32_t n_files, n_groups;
  orc_file_stat* stats;
} job_t;

typedef struct {
  pthread_t tid;
  int index;
  int32_t f0, f1;            /* slice of this call */
  int64_t* counts;           /* [(n_groups_cap + 1) * ORC_K] */
  int rc;
} worker_t;

static struct {
  int n;                     /* workers (the calling thread is not one of them) */
  worker_t* w;
  int32_t groups_cap;
  pthread_mutex_t mu;
  pthread_cond_t go, done;
  unsigned long generation;  /* bumped per call */
  int pending;
  int quit;
  job_t job;
} P;

static void* worker_main(void* arg) {
  worker_t* me = (worker_t*)arg;
  unsigned long seen = 0;
  for (;;) {
    pthread_mutex_lock(&P.mu);
    while (!P.quit && P.generation == seen) pthread_cond_wait(&P.go, &P.mu);
    if (P.quit) { pthread_mutex_unlock(&P.mu); return NULL; }
    seen = P.generation;
    const job_t j = P.job;
    pthread_mutex_unlock(&P.mu);
    const int32_t n = me->f1 - me->f0;
    me->rc = 0;
    if (n > 0)
      me->rc = orc_scan(j.arena, j.off + me->f0, j.len + me->f0, j.ext ? j.ext + me->f0 : NULL, j.grp ? j.grp + me->f0 : NULL,
                        n, j.n_groups, j.stats ? j.stats + me->f0 : NULL, me->counts, me->counts + (size_t)j.n_groups * ORC_K,
                        NULL, 0, NULL, NULL, 0, NULL, NULL, NULL);
    else
      memset(me->counts, 0, sizeof(int64_t) * (size_t)(j.n_groups + 1) * ORC_K);
    pthread_mutex_lock(&P.mu);
    if (--P.pending == 0) pthread_cond_signal(&P.done);
    pthread_mutex_unlock(&P.mu);
  }
}

int orc_mt_affinity_cpus(void) {
  cpu_set_t set;
  if (sched_getaffinity(0, sizeof set, &set) == 0) {
    const int c = CPU_COUNT(&set);
    if (c > 0) return c;
  }
  return 1;
}

void orc_mt_destroy(void) {
  if (!P.w) return;
  pthread_mutex_lock(&P.mu);
  P.quit = 1;
  pthread_cond_broadcast(&P.go);
  pthread_mutex_unlock(&P.mu);
  for (int i = 0; i < P.n; ++i) { pthread_join(P.w[i].tid, NULL); free(P.w[i].counts); }
  free(P.w);
  pthread_mutex_destroy(&P.mu); pthread_cond_destroy(&P.go); pthread_cond_destroy(&P.done);
  memset(&P, 0, sizeof P);
}

/* Start (or restart) the pool: n_threads <= 0 means one per CPU of the affinity mask.  Returns the thread
 * count, or -1. */
int orc_mt_create(int n_threads, int32_t max_groups) {
  orc_mt_destroy();
  if (n_threads <= 0) n_threads = orc_mt_affinity_cpus();
  if (max_groups < 1) max_groups = 1;
  memset(&P, 0, sizeof P);
  P.w = (worker_t*)calloc((size_t)n_threads, sizeof(worker_t));
  if (!P.w) return -1;
  P.groups_cap = max_groups;
  pthread_mutex_init(&P.mu, NULL); pthread_cond_init(&P.go, NULL); pthread_cond_init(&P.done, NULL);
  for (int i = 0; i < n_threads; ++i) {
    P.w[i].index = i;
    P.w[i].counts = (int64_t*)calloc((size_t)(max_groups + 1) * ORC_K, sizeof(int64_t));
    if (!P.w[i].counts || pthread_create(&P.w[i].tid, NULL, worker_main, &P.w[i]) != 0) {
      P.n = i; orc_mt_destroy(); return -1;
    }
    P.n = i + 1;
  }
  return P.n;
}

/* One scan over the whole corpus with the pool: static partition by bytes.  Same outputs as orc_scan
 * (per-file records and the two count tables; no events).  Returns 0, -1 (malformed corpus) or -2 (no pool /
 * n_groups above what the pool was created for). */
int orc_mt_scan(const uint8_t* arena, const int32_t* off, const int32_t* len, const uint8_t* ext, const uint16_t* grp,
                int32_t n_files, int32_t n_groups, orc_file_stat* stats, int64_t* group_counts, int64_t* global_counts) {
  if (!P.w || n_groups > P.groups_cap || n_groups < 1 || n_files < 0) return -2;
  int64_t total = 0;
  for (int32_t f = 0; f < n_files; ++f) total += len[f] + 64;   /* a per-file constant keeps many empty files spread out */
  int32_t f = 0;
  int64_t acc = 0;
  for (int t = 0; t < P.n; ++t) {
    const int64_t want = total * (t + 1) / P.n;
    P.w[t].f0 = f;
    while (f < n_files && acc + len[f] + 64 <= want) { acc += len[f] + 64; ++f; }
    if (t == P.n - 1) f = n_files;
    P.w[t].f1 = f;
  }
  pthread_mutex_lock(&P.mu);
  P.job.arena = arena; P.job.off = off; P.job.len = len; P.job.ext = ext; P.job.grp = grp;
  P.job.n_files = n_files; P.job.n_groups = n_groups; P.job.stats = stats;
  P.pending = P.n;
  P.generation++;
  pthread_cond_broadcast(&P.go);
  while (P.pending) pthread_cond_wait(&P.done, &P.mu);
  pthread_mutex_unlock(&P.mu);
  int rc = 0;
  if (group_counts) memset(group_counts, 0, sizeof(int64_t) * (size_t)n_groups * ORC_K);
  if (global_counts) memset(global_counts, 0, sizeof(int64_t) * ORC_K);
  for (int t = 0; t < P.n; ++t) {
    if (P.w[t].rc) rc = -1;
    if (group_counts) for (size_t i = 0; i < (size_t)n_groups * ORC_K; ++i) group_counts[i] += P.w[t].counts[i];
    if (global_counts) for (size_t i = 0; i < ORC_K; ++i) global_counts[i] += P.w[t].counts[(size_t)n_groups * ORC_K + i];
  }
  return rc;
}
