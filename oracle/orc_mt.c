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
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  int kind;                  /* 0 = scan, 1 = revision-pair diff with hunks (orc_diff_pairs_detail) */
  const uint8_t* arena; const int32_t* off; const int32_t* len; const uint8_t* ext; const uint16_t* grp;
  int32_t n_files, n_groups;
  orc_file_stat* stats;
  const uint8_t* arena2; const int32_t* off2; const int32_t* len2; const uint8_t* ext2;   /* the new side of the pairs */
  int64_t* added; int64_t* removed; orc_diff_detail* detail;
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
    if (j.kind == 1) {
      if (n > 0)
        me->rc = orc_diff_pairs_detail(j.arena, j.off + me->f0, j.len + me->f0, j.ext ? j.ext + me->f0 : NULL,
                                       j.arena2, j.off2 + me->f0, j.len2 + me->f0, j.ext2 ? j.ext2 + me->f0 : NULL, n,
                                       j.added + me->f0, j.removed + me->f0, j.detail ? j.detail + me->f0 : NULL);
    } else if (n > 0)
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

/* CPUs this process may use: the affinity mask, capped by the cgroup CPU quota when there is one (a container with
 * `cpu.max` = "800000 100000" gets 8 threads even if it sees 128 CPUs). */
int orc_mt_affinity_cpus(void) {
  cpu_set_t set;
  int c = 1;
  if (sched_getaffinity(0, sizeof set, &set) == 0 && CPU_COUNT(&set) > 0) c = CPU_COUNT(&set);
  long long quota = -1, period = 0;
  FILE* f = fopen("/sys/fs/cgroup/cpu.max", "r");                       /* cgroup v2 */
  if (f) {
    char q[32];
    if (fscanf(f, "%31s %lld", q, &period) == 2 && strcmp(q, "max") != 0) quota = atoll(q);
    fclose(f);
  } else {
    f = fopen("/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "r");              /* cgroup v1 */
    if (f) { if (fscanf(f, "%lld", &quota) != 1) quota = -1; fclose(f); }
    f = fopen("/sys/fs/cgroup/cpu/cpu.cfs_period_us", "r");
    if (f) { if (fscanf(f, "%lld", &period) != 1) period = 0; fclose(f); }
  }
  if (quota > 0 && period > 0) {
    const int q = (int)((quota + period - 1) / period);
    if (q >= 1 && q < c) c = q;
  }
  return c;
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
  memset(&P.job, 0, sizeof P.job);
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

/* Revision pairs (docs/SPEC.md section 8) with the pool: pairs partitioned statically by bytes (old + new), every thread
 * runs the plain orc_diff_pairs_detail over its slice.  Returns 0, -1 or -2 (no pool). */
int orc_mt_diff(const uint8_t* arena_old, const int32_t* off_old, const int32_t* len_old, const uint8_t* ext_old,
                const uint8_t* arena_new, const int32_t* off_new, const int32_t* len_new, const uint8_t* ext_new,
                int32_t n_pairs, int64_t* added, int64_t* removed, orc_diff_detail* detail) {
  if (!P.w || n_pairs < 0) return -2;
  int64_t total = 0;
  for (int32_t f = 0; f < n_pairs; ++f) total += (int64_t)len_old[f] + len_new[f] + 64;
  int32_t f = 0;
  int64_t acc = 0;
  for (int t = 0; t < P.n; ++t) {
    const int64_t want = total * (t + 1) / P.n;
    P.w[t].f0 = f;
    while (f < n_pairs && acc + len_old[f] + len_new[f] + 64 <= want) { acc += (int64_t)len_old[f] + len_new[f] + 64; ++f; }
    if (t == P.n - 1) f = n_pairs;
    P.w[t].f1 = f;
  }
  pthread_mutex_lock(&P.mu);
  memset(&P.job, 0, sizeof P.job);
  P.job.kind = 1;
  P.job.arena = arena_old; P.job.off = off_old; P.job.len = len_old; P.job.ext = ext_old;
  P.job.arena2 = arena_new; P.job.off2 = off_new; P.job.len2 = len_new; P.job.ext2 = ext_new;
  P.job.added = added; P.job.removed = removed; P.job.detail = detail; P.job.n_groups = 1;
  P.pending = P.n;
  P.generation++;
  pthread_cond_broadcast(&P.go);
  while (P.pending) pthread_cond_wait(&P.done, &P.mu);
  pthread_mutex_unlock(&P.mu);
  for (int t = 0; t < P.n; ++t) if (P.w[t].rc) return -1;
  return 0;
}
