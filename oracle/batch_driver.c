/*
 * batch_driver.c -- TEST INFRASTRUCTURE ONLY.  Runs a batch of (read, target) pairs through an ssw.h-ABI library
 * (the compiled reference oracle/_ref/libssw_ref.so, or any drop-in) with one pair per call and `threads` POSIX
 * threads pulling pairs from a shared counter -- "ssw.c one pair per thread across all host cores" as BASELINE.md
 * section 3 prescribes for the CPU baseline -- and records the 7 scalar result fields plus the CIGAR words.
 * The function pointers are passed in (dlsym'd by the caller), so this file links against nothing.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <time.h>

typedef struct { uint16_t score1, score2; int32_t ref_begin1, ref_end1, read_begin1, read_end1, ref_end2; uint32_t* cigar; int32_t cigarLen; } ref_align_t;
typedef void* (*init_fn)(const int8_t*, int32_t, const int8_t*, int32_t, int8_t);
typedef ref_align_t* (*align_fn)(const void*, const int8_t*, int32_t, uint8_t, uint8_t, uint8_t, uint16_t, int32_t, int32_t);
typedef void (*init_destroy_fn)(void*);
typedef void (*align_destroy_fn)(ref_align_t*);

typedef struct {
    init_fn init; align_fn align; init_destroy_fn idestroy; align_destroy_fn adestroy;
    const int8_t* reads; const int64_t* read_off; const int8_t* refs; const int64_t* ref_off;   /* CSR style, n+1 offsets */
    const int32_t* masklen;          /* per pair */
    const int8_t* mat; int32_t n; int32_t gapO, gapE, flag, filters, filterd, score_size;
    int64_t npairs;
    int32_t* out;                    /* npairs x 8: score1, score2, ref_begin1, ref_end1, read_begin1, read_end1, ref_end2, cigarLen (-1 if NULL) */
    uint32_t* cigar; int64_t* cigar_off; int32_t cigar_cap; /* per pair fixed-capacity slot of cigar_cap words (may be NULL) */
    volatile int64_t next;
} job_t;

static void* worker(void* arg)
{
    job_t* J = (job_t*)arg;
    for (;;) {
        int64_t i0 = __sync_fetch_and_add(&J->next, 16), i;
        if (i0 >= J->npairs) break;
        for (i = i0; i < i0 + 16 && i < J->npairs; ++i) {
            const int8_t* rd = J->reads + J->read_off[i]; int32_t rl = (int32_t)(J->read_off[i + 1] - J->read_off[i]);
            const int8_t* rf = J->refs + J->ref_off[i];   int32_t fl = (int32_t)(J->ref_off[i + 1] - J->ref_off[i]);
            void* p = J->init(rd, rl, J->mat, J->n, (int8_t)J->score_size);
            ref_align_t* a = J->align(p, rf, fl, (uint8_t)J->gapO, (uint8_t)J->gapE, (uint8_t)J->flag, (uint16_t)J->filters, J->filterd, J->masklen[i]);
            int32_t* o = J->out + i * 8;
            if (a) {
                o[0] = a->score1; o[1] = a->score2; o[2] = a->ref_begin1; o[3] = a->ref_end1; o[4] = a->read_begin1; o[5] = a->read_end1; o[6] = a->ref_end2; o[7] = a->cigarLen;
                if (J->cigar && a->cigarLen > 0) {
                    int32_t k, m = a->cigarLen < J->cigar_cap ? a->cigarLen : J->cigar_cap;
                    for (k = 0; k < m; ++k) J->cigar[i * (int64_t)J->cigar_cap + k] = a->cigar[k];
                }
                J->adestroy(a);
            } else { memset(o, 0, 32); o[7] = -1; }
            J->idestroy(p);
        }
    }
    return NULL;
}

/* returns elapsed wall seconds */
double oracle_run_batch(void* init, void* align, void* idestroy, void* adestroy,
                        const int8_t* reads, const int64_t* read_off, const int8_t* refs, const int64_t* ref_off, const int32_t* masklen,
                        const int8_t* mat, int32_t n, int32_t gapO, int32_t gapE, int32_t flag, int32_t filters, int32_t filterd, int32_t score_size,
                        int64_t npairs, int32_t threads, int32_t* out, uint32_t* cigar, int32_t cigar_cap)
{
    job_t J; pthread_t* th; int t; struct timespec a, b;
    memset(&J, 0, sizeof J);
    J.init = (init_fn)init; J.align = (align_fn)align; J.idestroy = (init_destroy_fn)idestroy; J.adestroy = (align_destroy_fn)adestroy;
    J.reads = reads; J.read_off = read_off; J.refs = refs; J.ref_off = ref_off; J.masklen = masklen; J.mat = mat; J.n = n;
    J.gapO = gapO; J.gapE = gapE; J.flag = flag; J.filters = filters; J.filterd = filterd; J.score_size = score_size;
    J.npairs = npairs; J.out = out; J.cigar = cigar; J.cigar_cap = cigar_cap; J.next = 0;
    if (threads < 1) threads = 1;
    th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)threads);
    clock_gettime(CLOCK_MONOTONIC, &a);
    for (t = 0; t < threads; ++t) pthread_create(&th[t], NULL, worker, &J);
    for (t = 0; t < threads; ++t) pthread_join(th[t], NULL);
    clock_gettime(CLOCK_MONOTONIC, &b);
    free(th);
    return (double)(b.tv_sec - a.tv_sec) + 1e-9 * (double)(b.tv_nsec - a.tv_nsec);
}

/* Same batch through the scalar restatement (ssw_oracle.c), single entry so tests can diff oracle vs reference quickly. */
typedef struct { uint16_t score1, score2; int32_t ref_begin1, ref_end1, read_begin1, read_end1, ref_end2, cigarLen, word_mode, status; } oracle_align_t;
void oracle_ssw_align(const int8_t*, int32_t, const int8_t*, int32_t, int32_t, const int8_t*, int32_t, int32_t, int32_t, int32_t, int32_t, int32_t, int32_t,
                      oracle_align_t*, uint32_t*, int32_t);

typedef struct { job_t J; } ojob_t;
static void* oworker(void* arg)
{
    job_t* J = (job_t*)arg;
    uint32_t* tmp = (uint32_t*)malloc(sizeof(uint32_t) * 1 << 20);
    for (;;) {
        int64_t i0 = __sync_fetch_and_add(&J->next, 16), i;
        if (i0 >= J->npairs) break;
        for (i = i0; i < i0 + 16 && i < J->npairs; ++i) {
            const int8_t* rd = J->reads + J->read_off[i]; int32_t rl = (int32_t)(J->read_off[i + 1] - J->read_off[i]);
            const int8_t* rf = J->refs + J->ref_off[i];   int32_t fl = (int32_t)(J->ref_off[i + 1] - J->ref_off[i]);
            oracle_align_t r; int32_t* o = J->out + i * 8;
            oracle_ssw_align(rd, rl, J->mat, J->n, J->score_size, rf, fl, J->gapO, J->gapE, J->flag, J->filters, J->filterd, J->masklen[i], &r, tmp, 1 << 20);
            if (r.status == 0) {
                o[0] = r.score1; o[1] = r.score2; o[2] = r.ref_begin1; o[3] = r.ref_end1; o[4] = r.read_begin1; o[5] = r.read_end1; o[6] = r.ref_end2; o[7] = r.cigarLen;
                if (J->cigar && r.cigarLen > 0) {
                    int32_t k, m = r.cigarLen < J->cigar_cap ? r.cigarLen : J->cigar_cap;
                    for (k = 0; k < m; ++k) J->cigar[i * (int64_t)J->cigar_cap + k] = tmp[k];
                }
            } else { memset(o, 0, 32); o[7] = -1; }
        }
    }
    free(tmp);
    return NULL;
}

double oracle_run_batch_port(const int8_t* reads, const int64_t* read_off, const int8_t* refs, const int64_t* ref_off, const int32_t* masklen,
                             const int8_t* mat, int32_t n, int32_t gapO, int32_t gapE, int32_t flag, int32_t filters, int32_t filterd, int32_t score_size,
                             int64_t npairs, int32_t threads, int32_t* out, uint32_t* cigar, int32_t cigar_cap)
{
    job_t J; pthread_t* th; int t; struct timespec a, b;
    memset(&J, 0, sizeof J);
    J.reads = reads; J.read_off = read_off; J.refs = refs; J.ref_off = ref_off; J.masklen = masklen; J.mat = mat; J.n = n;
    J.gapO = gapO; J.gapE = gapE; J.flag = flag; J.filters = filters; J.filterd = filterd; J.score_size = score_size;
    J.npairs = npairs; J.out = out; J.cigar = cigar; J.cigar_cap = cigar_cap; J.next = 0;
    if (threads < 1) threads = 1;
    th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)threads);
    clock_gettime(CLOCK_MONOTONIC, &a);
    for (t = 0; t < threads; ++t) pthread_create(&th[t], NULL, oworker, &J);
    for (t = 0; t < threads; ++t) pthread_join(th[t], NULL);
    clock_gettime(CLOCK_MONOTONIC, &b);
    free(th);
    return (double)(b.tv_sec - a.tv_sec) + 1e-9 * (double)(b.tv_nsec - a.tv_nsec);
}
