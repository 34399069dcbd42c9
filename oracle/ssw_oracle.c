/*
 * ssw_oracle.c -- TEST INFRASTRUCTURE ONLY.  Scalar CPU restatement of the Striped Smith-Waterman path
 * of MegaPath-Nano (reference: bin/realignment/realign/ssw.c).  Nothing in the product path
 * (megapath-nano_b200/) may include, link or call this file; only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs use it, and only as the checker.
 *
 * Parity status: PINNED.  The reference has no tests or golden vectors for this path,
 * so this restatement is pinned against the reference itself: oracle/Makefile compiles the reference
 * ssw.c where it lies into oracle/_ref/libssw_ref.so and tests/test_oracle_vs_ref.py fuzzes the two
 * against each other (all 7 scalar fields + CIGAR words); tests/golden/ holds vectors generated from that
 * compiled reference (tests/golden/make_golden.py).
 *
 * The restatement is deliberately NOT striped: it states what the SSE2 kernels compute, cell by cell.
 *   - plain Gotoh affine local alignment, gap of length k costs gapO + (k-1)*gapE
 *   - the query is padded to a multiple of W rows (W = 16 in the 8-bit kernel, 8 in the 16-bit kernel) with
 *     rows that score 0 against everything; pad rows take part in the per-column maximum
 *     (ssw.c:95,108 byte profile; ssw.c:335,346 word profile)
 *   - 16-bit kernel saturates the diagonal add at 32767 (ssw.c:425, _mm_adds_epi16)
 *   - 8-bit kernel result is only used when score + bias < 255 (ssw.c:271,302; ssw.c:789)
 * The "lazy-F does not update E" detail of ssw.c:226,450 does not change any H value when gapO > gapE
 * (SURVEY.md section 8a); the fuzz against the compiled reference confirms it.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

typedef struct {
    uint16_t score1;
    uint16_t score2;
    int32_t ref_begin1;
    int32_t ref_end1;
    int32_t read_begin1;
    int32_t read_end1;
    int32_t ref_end2;
    int32_t cigarLen;   /* number of words written to the caller's cigar buffer (0 if none) */
    int32_t word_mode;  /* 1 if the 16-bit kernel produced the result */
    int32_t status;     /* 0 ok, 1 = reference would return NULL (8-bit overflow without word profile), 2 = no profile,
                           3 = traceback error, 4 = cigar buffer too small */
} oracle_align_t;

typedef struct {
    int32_t score;   /* best score (already mapped to 255 on 8-bit overflow) */
    int32_t ref;     /* end_ref  (ssw.c:145 / :371 initial values kept) */
    int32_t read;    /* end_read */
    int32_t score2;
    int32_t ref2;
} oracle_ends_t;

static inline int32_t imax(int32_t a, int32_t b) { return a > b ? a : b; }
static inline int32_t imin(int32_t a, int32_t b) { return a < b ? a : b; }

/* One score pass.  Restates sw_sse2_byte (ssw.c:123-328, byte_mode=1) and sw_sse2_word (ssw.c:354-530, byte_mode=0).
 *   reverse   : 0 = columns 0..refLen-1, 1 = columns refLen-1..0 (ssw.c:178-182 / :401-405)
 *   terminate : stop after the first column whose maximum equals it (ssw.c:281 / :483); pass -1 for "never"
 *   colmax_out: optional, refLen entries, the maxColumn[] array of the reference (unvisited columns stay 0)
 */
void oracle_score_pass(const int8_t* read, int32_t readLen, const int8_t* ref, int32_t refLen, const int8_t* mat, int32_t n,
                       int32_t gapO, int32_t gapE, int32_t byte_mode, int32_t bias, int32_t reverse, int32_t terminate, int32_t maskLen,
                       oracle_ends_t* out, int32_t* colmax_out)
{
    const int32_t W = byte_mode ? 16 : 8;
    const int32_t segLen = (readLen + W - 1) / W;
    const int32_t Lp = segLen * W;                      /* padded row count */
    const int32_t cap = byte_mode ? 255 : 32767;
    int32_t* H = (int32_t*)calloc((size_t)Lp + 1, sizeof(int32_t));      /* H of the previous column */
    int32_t* E = (int32_t*)calloc((size_t)Lp + 1, sizeof(int32_t));      /* E entering the current column */
    int32_t* Hn = (int32_t*)calloc((size_t)Lp + 1, sizeof(int32_t));
    int32_t* saved = (int32_t*)calloc((size_t)Lp + 1, sizeof(int32_t)); /* pvHmax: calloc'ed, so all 0 until a column is saved */
    int32_t* colmax = (int32_t*)calloc((size_t)(refLen > 0 ? refLen : 1), sizeof(int32_t));
    int32_t best = 0, overflow = 0;
    int32_t end_ref = byte_mode ? -1 : 0;               /* ssw.c:145 vs ssw.c:371 */
    int32_t end_read = readLen - 1;                     /* ssw.c:144 / :370 */
    int32_t i, r, k;

    for (k = 0; k < refLen; ++k) {
        int32_t cm = 0, F = 0;
        i = reverse ? refLen - 1 - k : k;
        for (r = 0; r < Lp; ++r) {
            int32_t s = r < readLen ? (int32_t)mat[(int32_t)ref[i] * n + (int32_t)read[r]] : 0;   /* pad rows score 0 */
            int32_t h = (r > 0 ? H[r - 1] : 0) + s;
            int32_t v;
            if (h > cap) h = cap;                       /* _mm_adds_epi16 / _mm_adds_epu8 saturation */
            v = imax(imax(0, h), imax(E[r], F));
            Hn[r] = v;
            E[r] = imax(0, imax(E[r] - gapE, v - gapO));
            F = imax(0, imax(F - gapE, v - gapO));
            if (v > cm) cm = v;
        }
        memcpy(H, Hn, (size_t)Lp * sizeof(int32_t));
        if (cm > best) {                                /* strict: the first column in processing order wins (ssw.c:269 / :474) */
            best = cm;
            if (byte_mode && best + bias >= 255) { overflow = 1; break; }   /* ssw.c:271 */
            end_ref = i;
            memcpy(saved, Hn, (size_t)Lp * sizeof(int32_t));
        }
        colmax[i] = cm;
        if (cm == terminate) break;
    }
    /* smallest row of the saved column holding the best score (ssw.c:284-293 / :487-495) */
    for (r = 0; r < Lp; ++r)
        if (saved[r] == best && r < end_read) end_read = r;

    out->score = byte_mode ? ((overflow || best + bias >= 255) ? 255 : best) : best;
    out->ref = end_ref;
    out->read = end_read;
    /* second best: first strictly greater column maximum outside the mask (ssw.c:306-323 / :508-525) */
    out->score2 = 0;
    out->ref2 = 0;
    {
        int32_t edge = imax(end_ref - maskLen, 0);
        for (i = 0; i < edge; ++i)
            if (colmax[i] > out->score2) { out->score2 = colmax[i]; out->ref2 = i; }
        edge = (end_ref + maskLen) > refLen ? refLen : (end_ref + maskLen);
        for (i = edge + (byte_mode ? 1 : 0); i < refLen; ++i)        /* byte: edge+1 (ssw.c:318); word: edge (ssw.c:520) */
            if (colmax[i] > out->score2) { out->score2 = colmax[i]; out->ref2 = i; }
    }
    if (colmax_out) memcpy(colmax_out, colmax, (size_t)refLen * sizeof(int32_t));
    free(H); free(E); free(Hn); free(saved); free(colmax);
}

/* banded traceback, restating banded_sw (ssw.c:532-718).  Direction codes per cell: three slots (E, F, H) as in ssw.c:587-589.
 * Returns number of cigar words (BAM encoded len<<4|op, M=0 I=1 D=2), or -3 on a traceback error (reference returns 0/NULL),
 * or -4 if cap is too small. */
static int32_t band_x(int32_t i, int32_t w) { int32_t x = i - w; return x > 0 ? x : 0; }

int32_t oracle_banded_cigar(const int8_t* ref, const int8_t* read, int32_t refLen, int32_t readLen, int32_t score,
                            int32_t gapO, int32_t gapE, int32_t band_width, const int8_t* mat, int32_t n,
                            uint32_t* cigar_out, int32_t cigar_cap, int32_t* final_band)
{
    int32_t bw = band_width, width = 0, width_d = 0, max = 0;
    int32_t *hb = NULL, *eb = NULL, *hc = NULL;
    int8_t* dir = NULL;
    int32_t i, j;
    do {
        size_t cells;
        width = bw * 2 + 3; width_d = bw * 2 + 1;
        free(hb); free(eb); free(hc); free(dir);
        hb = (int32_t*)calloc((size_t)width + 1, sizeof(int32_t));   /* previous row H, band coordinates of that row */
        eb = (int32_t*)calloc((size_t)width + 1, sizeof(int32_t));   /* previous row E */
        hc = (int32_t*)calloc((size_t)width + 1, sizeof(int32_t));   /* current row H */
        cells = (size_t)width_d * (size_t)readLen * 3;
        dir = (int8_t*)calloc(cells + 3, 1);
        /* NOTE: the reference keeps e_b across attempts (realloc); every e_b entry it reads was written in the same attempt
           or zeroed through the `edge` slot, so a fresh zeroed array is equivalent. */
        for (i = 0; i < readLen; ++i) {
            int32_t beg = imax(0, i - bw), end = imin(refLen - 1, i + bw);
            int32_t edge = imin(end + 1, width - 1);
            int32_t f = 0, u = 0;
            int8_t* line = dir + (size_t)width_d * (size_t)i * 3;
            hb[0] = 0; eb[0] = 0; hb[edge] = 0; eb[edge] = 0; hc[0] = 0;     /* ssw.c:580, including the clipped-band quirk */
            for (j = beg; j <= end; ++j) {
                int32_t xi = band_x(i, bw), xp = band_x(i - 1, bw);
                int32_t e_idx = j - xp + 1;          /* (i-1, j)   in previous-row coordinates */
                int32_t d_idx = j - 1 - xp + 1;      /* (i-1, j-1) */
                int32_t b_idx = j - 1 - xi + 1;      /* (i,   j-1) */
                int32_t cell = (j - xi) * 3;
                int32_t open, ext, E, e1, f1, t1, t2, hval;
                int8_t de, df, dh;
                u = j - xi + 1;
                open = i == 0 ? -gapO : hb[e_idx] - gapO;
                ext = i == 0 ? -gapE : eb[e_idx] - gapE;
                E = open > ext ? open : ext;  de = open > ext ? 3 : 2;          /* ties extend (ssw.c:593-594) */
                eb[u] = E;
                open = hc[b_idx] - gapO;
                ext = f - gapE;
                f = open > ext ? open : ext;  df = open > ext ? 5 : 4;          /* ssw.c:596-599 */
                e1 = E > 0 ? E : 0;  f1 = f > 0 ? f : 0;
                t1 = e1 > f1 ? e1 : f1;
                t2 = hb[d_idx] + mat[(int32_t)ref[j] * n + (int32_t)read[i]];
                hval = t1 > t2 ? t1 : t2;
                hc[u] = hval;
                if (hval > max) max = hval;
                if (t1 <= t2) dh = 1; else dh = e1 > f1 ? de : df;              /* ssw.c:609-610 */
                line[cell + 0] = de; line[cell + 1] = df; line[cell + 2] = dh;
            }
            for (j = 1; j <= u; ++j) hb[j] = hc[j];
        }
        bw *= 2;
    } while (max < score);
    bw /= 2;
    if (final_band) *final_band = bw;

    /* trace back from the bottom-right corner until row 0 (ssw.c:618-679) */
    {
        int32_t l = 0, run = 0, state = 2, cap = 64, k, a, b;
        uint32_t* c = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)cap);
        char op = 'M', prev = 'M';
        size_t total = (size_t)width_d * (size_t)readLen * 3;
        i = readLen - 1; j = refLen - 1;
        while (i > 0) {
            long idx = (long)width_d * i * 3 + (long)(j - band_x(i, bw)) * 3 + state;
            int8_t d = (idx >= 0 && (size_t)idx < total) ? dir[idx] : 0;
            switch (d) {
                case 1: --i; --j; state = 2; op = 'M'; break;
                case 2: --i; state = 0; op = 'I'; break;
                case 3: --i; state = 2; op = 'I'; break;
                case 4: --j; state = 1; op = 'D'; break;
                case 5: --j; state = 2; op = 'D'; break;
                default:
                    free(c); free(hb); free(eb); free(hc); free(dir);
                    return -3;
            }
            if (op == prev) ++run;
            else {
                if (l + 3 >= cap) { cap *= 2; c = (uint32_t*)realloc(c, sizeof(uint32_t) * (size_t)cap); }
                c[l++] = ((uint32_t)run << 4) | (prev == 'M' ? 0u : prev == 'I' ? 1u : 2u);
                prev = op; run = 1;
            }
        }
        if (l + 3 >= cap) { cap *= 2; c = (uint32_t*)realloc(c, sizeof(uint32_t) * (size_t)cap); }
        if (op == 'M') c[l++] = ((uint32_t)(run + 1) << 4);                       /* ssw.c:680-687 */
        else { c[l++] = ((uint32_t)run << 4) | (op == 'I' ? 1u : 2u); c[l++] = (1u << 4); }   /* ssw.c:688-697 */
        if (l > cigar_cap) { free(c); free(hb); free(eb); free(hc); free(dir); return -4; }
        for (a = 0, b = l - 1, k = 0; k < l; ++k, ++a, --b) cigar_out[a] = c[b];   /* reversed (ssw.c:699-708) */
        free(c); free(hb); free(eb); free(hc); free(dir);
        return l;
    }
}

/* ssw_init + ssw_align in one call (ssw.c:733-754, 762-852). */
void oracle_ssw_align(const int8_t* read, int32_t readLen, const int8_t* mat, int32_t n, int32_t score_size,
                      const int8_t* ref, int32_t refLen, int32_t gapO, int32_t gapE, int32_t flag, int32_t filters, int32_t filterd,
                      int32_t maskLen, oracle_align_t* r, uint32_t* cigar_out, int32_t cigar_cap)
{
    oracle_ends_t fw, rv;
    int32_t bias = 0, i, word = 0;
    int32_t have_byte = (score_size == 0 || score_size == 2), have_word = (score_size == 1 || score_size == 2);
    memset(r, 0, sizeof(*r));
    r->ref_begin1 = -1; r->read_begin1 = -1;
    if (have_byte) {
        for (i = 0; i < n * n; ++i) if (mat[i] < bias) bias = mat[i];
        bias = abs(bias) & 0xff;                                                   /* stored in a uint8_t (ssw.c:85,745) */
        oracle_score_pass(read, readLen, ref, refLen, mat, n, gapO, gapE, 1, bias, 0, -1, maskLen, &fw, NULL);
        if (fw.score == 255) {
            if (!have_word) { r->status = 1; return; }
            oracle_score_pass(read, readLen, ref, refLen, mat, n, gapO, gapE, 0, 0, 0, -1, maskLen, &fw, NULL);
            word = 1;
        }
    } else if (have_word) {
        oracle_score_pass(read, readLen, ref, refLen, mat, n, gapO, gapE, 0, 0, 0, -1, maskLen, &fw, NULL);
        word = 1;
    } else { r->status = 2; return; }
    r->word_mode = word;
    r->score1 = (uint16_t)fw.score; r->ref_end1 = fw.ref; r->read_end1 = fw.read;
    if (maskLen >= 15) { r->score2 = (uint16_t)fw.score2; r->ref_end2 = fw.ref2; }
    else { r->score2 = 0; r->ref_end2 = -1; }
    if (flag == 0 || (flag == 2 && r->score1 < filters)) return;                   /* ssw.c:817 */

    /* reverse pass over read[0..read_end1] reversed and ref[0..ref_end1] walked backwards (ssw.c:820-832) */
    {
        int32_t rl = r->read_end1 + 1, k;
        int8_t* rr = (int8_t*)malloc((size_t)(rl > 0 ? rl : 1));
        for (k = 0; k < rl; ++k) rr[k] = read[rl - 1 - k];
        oracle_score_pass(rr, rl, ref, r->ref_end1 + 1, mat, n, gapO, gapE, word ? 0 : 1, bias, 1, r->score1, maskLen, &rv, NULL);
        free(rr);
        r->ref_begin1 = rv.ref;
        r->read_begin1 = r->read_end1 - rv.read;
    }
    if ((7 & flag) == 0 || ((2 & flag) != 0 && r->score1 < filters) ||
        ((4 & flag) != 0 && (r->ref_end1 - r->ref_begin1 > filterd || r->read_end1 - r->read_begin1 > filterd))) return;   /* ssw.c:833 */
    {
        int32_t sub_ref = r->ref_end1 - r->ref_begin1 + 1, sub_read = r->read_end1 - r->read_begin1 + 1;
        int32_t band = abs(sub_ref - sub_read) + 1;
        int32_t l = oracle_banded_cigar(ref + r->ref_begin1, read + r->read_begin1, sub_ref, sub_read, r->score1, gapO, gapE, band, mat, n,
                                        cigar_out, cigar_cap, NULL);
        if (l == -3) { r->status = 3; return; }
        if (l == -4) { r->status = 4; return; }
        r->cigarLen = l;
    }
}
