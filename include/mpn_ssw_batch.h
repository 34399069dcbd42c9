/*
 * mpn_ssw_batch.h -- batched submit/collect C ABI of the B200 Smith-Waterman engine.
 *
 * This is the new entry point the reference does not have: where MegaPath-Nano calls
 *     ssw_init -> ssw_align -> align_destroy -> init_destroy            (ssw.h:77-130)
 * once per (read, target) pair -- pyssw.py:137-147 for the ONT path, ssw_cpp.cpp:326-359 called from
 * realigner.cpp:338,368 for the Illumina path -- a caller hands over many pairs at once and gets back one record per
 * pair with exactly the fields of s_align (ssw.h:47-57) plus the CIGAR words in one arena.
 * The legacy per-pair ABI in include/ssw.h is implemented on top of this one (batch of one).
 *
 * Plain C types only.  Sequences are int8 codes in [0, n) exactly as ssw_align takes them (ssw.h:87-89); all pairs of
 * a batch share one scoring (mat, n, gapO, gapE), one score_size and one (flag, filters, filterd) triple; maskLen is
 * per pair.  Results are bit-identical to ssw.c for gapO > gapE (every reference caller uses 8/2).
 *
 * Errors: functions return 0 on success, a negative MPN_E_* otherwise; CUDA failures abort() with a message on stderr
 * (the reference's callers never check for NULL, SURVEY.md section 8b, so failing loudly is the only safe behaviour).
 * There is no CPU fallback.
 *
 * Threading: an mpn_engine is NOT re-entrant -- one thread at a time per engine (it owns one device, its streams and its
 * staging buffers).  Use one engine per thread, or an mpn_pool (below), which owns one engine + one host thread per device.
 */
#ifndef MPN_SSW_BATCH_H
#define MPN_SSW_BATCH_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct mpn_engine mpn_engine;
typedef struct mpn_batch mpn_batch;

/* scoring + reporting options of one batch: the arguments of ssw_init (ssw.h:77) and ssw_align (ssw.h:117-125) */
typedef struct {
    const int8_t* mat;      /* n*n substitution matrix, row = target code, column = read code (ssw.c:108,191) */
    int32_t n;
    int32_t gapO, gapE;     /* absolute values, 0..255 (uint8_t in ssw_align) */
    int32_t score_size;     /* 0: 8-bit only, 1: 16-bit only, 2: 8-bit with 16-bit re-run (ssw.h:64-66) */
    int32_t flag;           /* ssw_align flag byte */
    int32_t filters;        /* uint16_t score filter */
    int32_t filterd;        /* distance filter */
} mpn_params;

/* one record per pair: the scalar fields of s_align (ssw.h:47-57); the CIGAR is cigar[cigar_off .. cigar_off+cigar_len) */
typedef struct {
    uint16_t score1;
    uint16_t score2;
    int32_t ref_begin1;
    int32_t ref_end1;
    int32_t read_begin1;
    int32_t read_end1;
    int32_t ref_end2;
    int32_t cigar_len;
    int32_t status;         /* 0 ok; MPN_ST_NULL / MPN_ST_NULL_TRACE: ssw_align would have returned NULL for this pair */
    int64_t cigar_off;
} mpn_result;

/* MPN_ST_NULL: 8-bit scores saturated and no 16-bit profile was asked for (ssw.c:793-796); MPN_ST_NULL_TRACE: banded traceback failed (ssw.c:840-843) */
enum { MPN_ST_OK = 0, MPN_ST_NULL = 1, MPN_ST_NULL_TRACE = 2 };
enum { MPN_E_ARG = -1, MPN_E_NOGPU = -2, MPN_E_CIGAR_SPACE = -3, MPN_E_UNSUPPORTED = -4 };

/* engine = one CUDA device + one stream + reusable device/pinned buffers.  device < 0: current device. */
mpn_engine* mpn_engine_create(int device);
void mpn_engine_destroy(mpn_engine* e);
/* run on a caller-owned cudaStream_t (e.g. PyTorch's current stream) instead of the engine's own; 0 restores the own stream */
int mpn_engine_set_stream(mpn_engine* e, void* cuda_stream);
/* optional phase timing (CUDA events on the engine's stream): enable, run a batch, then read the milliseconds of
 * {forward score kernels, second-best/mode epilogue, reverse score kernels, traceback+CIGAR} of the last mpn_batch_run */
int mpn_engine_set_profile(mpn_engine* e, int on);
int mpn_engine_phase_ms(mpn_engine* e, float* ms4);
/* the same averaged over the profiled runs since profiling was switched on (at most the last 16 runs); *nruns receives the count */
int mpn_engine_phase_ms_mean(mpn_engine* e, float* ms4, int* nruns);
/* counters since creation: kernel launches, pairs, forward cells, pairs re-run in the 32-bit kernel */
int mpn_engine_stats(const mpn_engine* e, int64_t* launches, int64_t* pairs, int64_t* cells, int64_t* wide_pairs);

/*
 * One call, host buffers in, host records out (the call a user makes; copies are inside):
 *   reads/refs     concatenated int8 codes; read i = reads[read_off[i] .. read_off[i+1]), same for refs (npairs+1 offsets)
 *   masklen        per pair (ssw.h:103-110)
 *   out            npairs records
 *   cigar, cigar_cap   caller's CIGAR arena in uint32 words; returns MPN_E_CIGAR_SPACE if too small
 *                      (npairs * 2 * max read length is always enough)
 */
int mpn_align_batch(mpn_engine* e, const mpn_params* p, const int8_t* reads, const int64_t* read_off, const int8_t* refs,
                    const int64_t* ref_off, const int32_t* masklen, int64_t npairs, mpn_result* out, uint32_t* cigar, int64_t cigar_cap);

/*
 * The same call with NIBBLE-PACKED sequences: base i of the concatenated reads (targets) is the low nibble (i even) or the high nibble
 * (i odd) of reads4[i / 2] (refs4[i / 2]); read_off / ref_off stay in BASES.  Codes must be below 16 (n <= 16).  Half the host->device
 * bytes of mpn_align_batch -- on a box whose GPUs share PCIe uplinks that copy is what limits short-read batches;
 * a small kernel expands the nibbles on the device, everything after that is identical.  A caller that holds ASCII (every reference
 * caller does: ssw_cpp.cpp:311-323, pyssw.py:86-98) can translate straight to nibbles; mpn_pack4 converts int8 codes on host threads.
 */
int mpn_align_batch_packed4(mpn_engine* e, const mpn_params* p, const uint8_t* reads4, const int64_t* read_off, const uint8_t* refs4,
                            const int64_t* ref_off, const int32_t* masklen, int64_t npairs, mpn_result* out, uint32_t* cigar, int64_t cigar_cap);
void mpn_pack4(const int8_t* codes, int64_t n, uint8_t* out /* (n + 1) / 2 bytes */);

/*
 * The same with FOUR bases per byte: base i of a stream sits in bits 2 (i & 3) .. 2 (i & 3) + 1 of reads2[i / 4] (refs2[i / 4]); read_off / ref_off
 * stay in BASES.  Codes above 3 (N) are stored as 0 and listed per stream in read_exc / ref_exc as (position << 4 | code), sorted by position
 * (position counted over the whole stream, like the offsets).  A quarter of the host->device bytes of mpn_align_batch: on an 8-GPU box every
 * rank's batch has to arrive through the same host before its kernels can start.  mpn_pack2 converts int8 codes and returns
 * the number of exceptions found (call again with a larger array if it exceeds exc_cap).
 */
int mpn_align_batch_packed2(mpn_engine* e, const mpn_params* p, const uint8_t* reads2, const int64_t* read_off, const int64_t* read_exc, int64_t n_read_exc,
                            const uint8_t* refs2, const int64_t* ref_off, const int64_t* ref_exc, int64_t n_ref_exc,
                            const int32_t* masklen, int64_t npairs, mpn_result* out, uint32_t* cigar, int64_t cigar_cap);
int64_t mpn_pack2(const int8_t* codes, int64_t n, uint8_t* out /* (n + 3) / 4 bytes */, int64_t* exc, int64_t exc_cap);

/*
 * Same call for pairs that SHARE sequences (one haplotype against many reads, realigner.cpp:351-384): one arena of int8
 * codes plus, per pair, the start and length of its read and of its target inside the arena.  Spans may overlap or repeat;
 * every distinct sequence is uploaded once.
 */
int mpn_align_batch_spans(mpn_engine* e, const mpn_params* p, const int8_t* seq, int64_t seq_bytes, const int64_t* rd_start, const int32_t* rd_len,
                          const int64_t* rf_start, const int32_t* rf_len, const int32_t* masklen, int64_t npairs,
                          mpn_result* out, uint32_t* cigar, int64_t cigar_cap);

/*
 * The same work split into phases, so that inputs can stay resident in HBM across repeated runs:
 *   upload : host -> device copies + scheduling (length binning, task lists)
 *   run    : enqueue every kernel of the batch on the engine's stream (no host synchronisation)
 *   fetch  : device -> host copies of the records + CIGAR arena, synchronises
 */
mpn_batch* mpn_batch_upload(mpn_engine* e, const mpn_params* p, const int8_t* reads, const int64_t* read_off, const int8_t* refs,
                            const int64_t* ref_off, const int32_t* masklen, int64_t npairs);
mpn_batch* mpn_batch_upload_spans(mpn_engine* e, const mpn_params* p, const int8_t* seq, int64_t seq_bytes, const int64_t* rd_start, const int32_t* rd_len,
                                  const int64_t* rf_start, const int32_t* rf_len, const int32_t* masklen, int64_t npairs);
int mpn_batch_run(mpn_batch* b);
int mpn_batch_fetch(mpn_batch* b, mpn_result* out, uint32_t* cigar, int64_t cigar_cap);
/* bytes copied host->device by upload and device->host by fetch for this batch */
int mpn_batch_io_bytes(const mpn_batch* b, int64_t* h2d, int64_t* d2h);
void mpn_batch_free(mpn_batch* b);

/*
 * ---- one batch over several GPUs of a box ---------------------------------------------------------------
 * Replaces the reference's process fan-out (bin/realignment/realignment.sh:34-39, 50-60: GNU parallel, one process per chromosome /
 * candidate position) inside one process: a pool owns one engine and one host thread per device.  A batch is cut, in the caller's
 * order, into ranges of about equal cost (forward cells x relative cost of the score kernel a pair takes), heaviest first, that the
 * device threads pull from one queue; inside a range the engine bins by read length and sorts by target length as for any batch.
 * No collective: every device copies its ranges from the caller's buffers and writes its records into out[] at the range's position;
 * CIGAR words go to per-range regions of the caller's arena (not compact; mpn_result.cigar_off is the absolute index as everywhere).
 * Results are identical to mpn_align_batch on one device, pair by pair.
 *
 *   devices / ndev   device ordinals; devices == NULL: 0 .. ndev-1; ndev <= 0: every visible device.  NULL without a CUDA device.
 * A pool runs one batch at a time (calls from several threads serialise).
 */
typedef struct mpn_pool mpn_pool;
mpn_pool* mpn_pool_create(const int* devices, int ndev);
void mpn_pool_destroy(mpn_pool* pl);
int mpn_pool_ndev(const mpn_pool* pl);
mpn_engine* mpn_pool_engine(mpn_pool* pl, int k);      /* the engine of device k (statistics); owned by the pool */
int mpn_pool_align_batch(mpn_pool* pl, const mpn_params* p, const int8_t* reads, const int64_t* read_off, const int8_t* refs,
                         const int64_t* ref_off, const int32_t* masklen, int64_t npairs, mpn_result* out, uint32_t* cigar, int64_t cigar_cap);
/* spans form (pairs sharing sequences): one range per device, every device uploads the arena once */
int mpn_pool_align_batch_spans(mpn_pool* pl, const mpn_params* p, const int8_t* seq, int64_t seq_bytes, const int64_t* rd_start, const int32_t* rd_len,
                               const int64_t* rf_start, const int32_t* rf_len, const int32_t* masklen, int64_t npairs,
                               mpn_result* out, uint32_t* cigar, int64_t cigar_cap);
/* per device, for the last batch: host wall milliseconds of its share, pairs and forward cells it processed (arrays of mpn_pool_ndev entries, any may be NULL) */
int mpn_pool_last_shares(const mpn_pool* pl, double* ms, int64_t* pairs, int64_t* cells);

/*
 * ---- k-mer fast pass of the region realigner on the GPU ----------------------------------------
 * Replaces, for many regions at once, the reference's BuildIndex + FastAlignReadsToHaplotype + FastAlignStrings
 * (realigner.cpp:429-451, :170-230, :232-253): every read of a region against every haplotype of that region, ungapped, at most
 * 2 mismatches (N on either side matches), only at placements that share an exact 32-mer with the haplotype; best placement per
 * (haplotype, read) with the reference's evaluation order as tie-break; haplotype score = sum of its reads' best scores, 0 when a
 * non-reference haplotype has a base inside the window that no read covered by the time the scan reached it (:248-252).
 *
 *   text                     ASCII bases of all haplotypes and reads (any layout; spans below index it)
 *   hap_start/hap_len        nhaps spans; haplotypes of one region are consecutive
 *   hap_is_ref               1 if the haplotype equals the region's reference string (never dropped, :172)
 *   read_start/read_len      nreads spans; reads of one region are consecutive
 *   regions                  per region: its haplotypes, its reads, ref_prefix / ref_suffix, and where its placements start
 *   places                   out: for region g, haplotype h (local), read r (local): places[g.place_first + h * g.nread + r]
 *                            = {score, pos}; score 0 / pos -1 when the read has no placement on that haplotype
 *   hap_score                out, nhaps entries
 *   region_flag              out, nregions entries: 1 = the region holds a base other than A,C,G,T,N; its outputs are undefined and
 *                            the caller must use its own string-exact path for that region
 * Limits (MPN_E_UNSUPPORTED otherwise, nothing is computed): reads of at most 256 bases, haplotypes of at most 2816 bases.
 */
typedef struct { int32_t score; int32_t pos; } mpn_placement;
typedef struct {
    int64_t place_first;
    int32_t hap_first, nhap;
    int32_t read_first, nread;
    int32_t prefix, suffix;
} mpn_fp_region;
enum { MPN_FP_MAX_READ = 256, MPN_FP_MAX_HAP = 2816 };
int mpn_fastpass(mpn_engine* e, const char* text, int64_t text_bytes,
                 const int64_t* hap_start, const int32_t* hap_len, const uint8_t* hap_is_ref, int32_t nhaps,
                 const int64_t* read_start, const int32_t* read_len, int32_t nreads,
                 const mpn_fp_region* regions, int32_t nregions,
                 mpn_placement* places, int32_t* hap_score, uint8_t* region_flag);
/* milliseconds the kernel of the last mpn_fastpass call took on the device (CUDA events) */
float mpn_fastpass_last_kernel_ms(const mpn_engine* e);

#ifdef __cplusplus
}
#endif
#endif
