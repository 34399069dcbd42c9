/* Synthetic (read, target) pairs at memory speed -- TEST / BENCH INFRASTRUCTURE (used by workloads.py for BASELINE configs[4], where the
 * numpy generator would take minutes per pass: 10 M pairs of 100 bp .. 20 kb are ~80 GB of bases).
 * Every pair is a pure function of (seed, global pair index): any chunking of the workload, on any rank, yields the same bases.
 *   target: uniform A,C,G,T;  read: the target from a random offset on, with `err` errors split 1/3 substitution, 1/3 insertion,
 *   1/3 deletion (the distribution of SURVEY.md section 8d).
 * gcc -O2 -shared -fPIC -pthread -o libseqgen.so seqgen.c */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>

static inline uint64_t splitmix(uint64_t* s) { uint64_t z = (*s += 0x9e3779b97f4a7c15ull); z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull; z = (z ^ (z >> 27)) * 0x94d049bb133111ebull; return z ^ (z >> 31); }
static inline uint64_t xs(uint64_t* s) { uint64_t x = *s; x ^= x << 13; x ^= x >> 7; x ^= x << 17; return *s = x; }

typedef struct {
    uint64_t seed; int64_t first_index, lo, hi; const int64_t *read_off, *ref_off; double err; int8_t *reads, *refs;
} job_t;

static void gen_pair(uint64_t seed, int64_t gidx, int64_t rl, int64_t fl, double err, int8_t* read, int8_t* ref)
{
    uint64_t sm = seed ^ ((uint64_t)gidx * 0xd6e8feb86659fd93ull);
    uint64_t st = splitmix(&sm) | 1ull;
    int64_t i = 0;
    while (i < fl) {                                   /* 32 bases per random word */
        uint64_t r = xs(&st);
        for (int k = 0; k < 32 && i < fl; ++k, ++i, r >>= 2) ref[i] = (int8_t)(r & 3);
    }
    if (fl <= 0) { for (i = 0; i < rl; ++i) read[i] = (int8_t)(xs(&st) & 3); return; }
    int64_t slack = fl - rl - (int64_t)(rl * err) - 4;
    if (slack < 0) slack = 0;
    int64_t pos = (int64_t)(xs(&st) % (uint64_t)(slack + 1));
    const uint32_t t1 = (uint32_t)(err / 3.0 * 4294967296.0), t2 = 2 * t1, t3 = 3 * t1;
    for (i = 0; i < rl; ++i) {
        const uint64_t r = xs(&st);
        const uint32_t u = (uint32_t)r;
        if (pos > fl - 1) pos = fl - 1;
        int8_t b = ref[pos];
        if (u >= t3) { ++pos; }
        else if (u < t1) { b = (int8_t)((b + 1 + ((r >> 32) % 3)) & 3); ++pos; }            /* substitution */
        else if (u < t2) { b = (int8_t)((r >> 32) & 3); }                                     /* insertion: target position stays */
        else { ++pos; if (pos > fl - 1) pos = fl - 1; b = ref[pos]; ++pos; }                  /* deletion: skip one target base */
        read[i] = b;
    }
}

static void* worker(void* arg)
{
    const job_t* j = (const job_t*)arg;
    for (int64_t i = j->lo; i < j->hi; ++i)
        gen_pair(j->seed, j->first_index + i, j->read_off[i + 1] - j->read_off[i], j->ref_off[i + 1] - j->ref_off[i], j->err,
                 j->reads + (j->read_off[i] - j->read_off[0]), j->refs + (j->ref_off[i] - j->ref_off[0]));
    return 0;
}

/* pairs [0, n) of a chunk whose first pair has global index first_index; offsets are the chunk's CSR arrays (n + 1 entries) */
void seqgen_pairs(uint64_t seed, int64_t first_index, int64_t n, const int64_t* read_off, const int64_t* ref_off, double err,
                  int8_t* reads, int8_t* refs, int threads)
{
    if (n <= 0) return;
    if (threads < 1) threads = 1;
    if (threads > 64) threads = 64;
    if (threads > n) threads = (int)n;
    pthread_t th[64]; job_t jobs[64];
    const int64_t total = (read_off[n] - read_off[0]) + (ref_off[n] - ref_off[0]);
    int64_t at = 0;
    for (int t = 0; t < threads; ++t) {                /* split by bases */
        const int64_t want = total * (t + 1) / threads;
        int64_t hi = at;
        while (hi < n && ((read_off[hi] - read_off[0]) + (ref_off[hi] - ref_off[0])) < want) ++hi;
        if (t == threads - 1) hi = n;
        jobs[t] = (job_t){seed, first_index, at, hi, read_off, ref_off, err, reads, refs};
        at = hi;
    }
    for (int t = 1; t < threads; ++t) pthread_create(&th[t], 0, worker, &jobs[t]);
    worker(&jobs[0]);
    for (int t = 1; t < threads; ++t) pthread_join(th[t], 0);
}
