/*
 * ssw.h -- per-pair C ABI of the B200 Smith-Waterman engine.
 *
 * Drop-in for MegaPath-Nano's bin/realignment/realign/ssw.h (reference ssw.h:29-57 types, :77 ssw_init, :82 init_destroy,
 * :117-125 ssw_align, :130 align_destroy, :137-183 CIGAR helpers): same symbol names, same argument lists, same
 * s_align layout (40 bytes on x86-64, as mirrored by pyssw.py:7-16 `CAlignRes`), same ownership rules
 *   - ssw_init BORROWS `read` and `mat`: keep them alive until the last ssw_align on that profile (reference ssw.c:749-750)
 *   - ssw_align returns a malloc'ed record (+ malloc'ed cigar) that the caller releases with align_destroy (ssw.c:854-857)
 * so `ctypes.cdll.LoadLibrary("libssw.so")` in pyssw.py:33 / fast_align_reads2ref.py:35 and the C++ wrapper ssw_cpp.cpp
 * keep working unchanged.  Every call is executed on the GPU as a batch of one through include/mpn_ssw_batch.h; there is
 * no CPU implementation behind it.  Use the batched ABI when more than a handful of pairs are available.
 */
#ifndef SSW_H
#define SSW_H

#include <stdint.h>
#include <stdio.h>
#include <string.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MAPSTR "MIDNSHP=X"
#ifndef BAM_CIGAR_SHIFT
#define BAM_CIGAR_SHIFT 4
#endif

struct _profile;
typedef struct _profile s_profile;

/* result record; field order and widths are part of the ABI (reference ssw.h:47-57) */
typedef struct {
    uint16_t score1;       /* best score */
    uint16_t score2;       /* second best score outside the mask window */
    int32_t ref_begin1;    /* 0-based, -1 if not computed */
    int32_t ref_end1;
    int32_t read_begin1;   /* 0-based, -1 if not computed */
    int32_t read_end1;
    int32_t ref_end2;
    uint32_t* cigar;       /* BAM encoding: length << 4 | op, op 0/1/2 = M/I/D; NULL if not computed */
    int32_t cigarLen;
} s_align;

s_profile* ssw_init(const int8_t* read, const int32_t readLen, const int8_t* mat, const int32_t n, const int8_t score_size);
void init_destroy(s_profile* p);
s_align* ssw_align(const s_profile* prof, const int8_t* ref, int32_t refLen, const uint8_t weight_gapO, const uint8_t weight_gapE,
                   const uint8_t flag, const uint16_t filters, const int32_t filterd, const int32_t maskLen);
void align_destroy(s_align* a);

/* CIGAR word helpers (reference ssw.h:137-183) */
static inline uint32_t to_cigar_int(uint32_t length, char op_letter)
{
    const char* p = strchr(MAPSTR, op_letter);
    const uint32_t code = (p && *p) ? (uint32_t)(p - MAPSTR) : 0u;      /* unknown letters encode as M, like the reference's default case */
    return (length << BAM_CIGAR_SHIFT) | code;
}
static inline char cigar_int_to_op(uint32_t cigar_int)
{
    const uint32_t code = cigar_int & 0xfU;
    return code > 8 ? 'M' : MAPSTR[code];
}
static inline uint32_t cigar_int_to_len(uint32_t cigar_int) { return cigar_int >> BAM_CIGAR_SHIFT; }

#ifdef __cplusplus
}
#endif
#endif /* SSW_H */
