// Generic banded traceback + CIGAR kernel (sm_100a): banded_sw (ssw.c:532-718) and the begin-position / CIGAR part of
// ssw_align (ssw.c:830-848) for ANY band width, one thread per pair.  It is the last stage of the traceback chain -- the
// register-row kernels (sw_trace_rows.cuh) take bands of up to 15 cells, the warp kernel (sw_trace_warp.cuh) up to 544 -- and
// only sees what those hand on (status 8).  This file also defines the types the whole chain shares.  Direction bytes live in a
// global scratch arena carved up with an atomic bump pointer (the band is doubled until the banded maximum reaches the score, so
// the size is only known on the device); CIGAR words go to a second arena the same way.
//
// banded_sw semantics kept:
//   band attempt bw: row i covers columns max(0,i-bw) .. min(refLen-1,i+bw); E/F are NOT floored; ties prefer the
//   diagonal, then F; E and F open only when strictly better than extending (ssw.c:593-599, 609-610); the previous-row
//   arrays are addressed in band coordinates and slot `edge` is zeroed before every row (ssw.c:580) -- including the case
//   where that zeroes the valid cell above the last target column; bw doubles until max >= score (ssw.c:555,614-615);
//   traceback runs from the bottom-right corner while i > 0 (ssw.c:625), run-length encodes, appends the row-0 match
//   (ssw.c:680-697) and reverses.
#pragma once
#include "sw_finish.cuh"

namespace mpn {

struct TraceParams {
    int32_t flag, filters, filterd;
    int32_t gapO, gapE, n;
    const int8_t* mat;            // n*n, device
    int32_t lane_max_rows;        // narrow bands: reads with more rows go to the warp-per-pair kernel instead of the lane-per-pair one
};

struct FinalResult {              // second half of s_align (ssw.h:47-57)
    int32_t ref_begin1, read_begin1;
    int32_t cigar_len;
    int32_t status;               // 0 ok; 3 traceback error (reference returns NULL, ssw.c:840-843); 5 scratch arena exhausted; 6 cigar arena exhausted
    int64_t cigar_off;            // word offset into the CIGAR arena
};

struct Arena {
    uint8_t* base;
    unsigned long long bytes;
    unsigned long long* used;
};

__device__ __forceinline__ int band_x(int i, int w) { int x = i - w; return x > 0 ? x : 0; }

constexpr int TRACE_BLOCK = 64;
constexpr int TRACE_SMEM_BW = 24;                         // bands up to this half-width keep their row buffers in shared memory
constexpr int TRACE_SMEM_W = 2 * TRACE_SMEM_BW + 4;       // entries per row buffer
constexpr size_t TRACE_SMEM_BYTES = 3ull * TRACE_SMEM_W * TRACE_BLOCK * sizeof(int);
inline size_t trace_smem_bytes(int n) { return TRACE_SMEM_BYTES + (((size_t)n * n + 15) & ~(size_t)15); }

// The banded DP of every lane is driven by ONE warp-wide loop that advances each lane by one cell per iteration (row changes and
// band restarts are predicated inside it), so lanes with different read lengths / band widths stay converged instead of
// serialising their nested row x column loops.
__global__ void __launch_bounds__(TRACE_BLOCK)
sw_trace_wide_kernel(const SwTask* __restrict__ order, int ntasks, const int8_t* __restrict__ seq, const FwdResult* __restrict__ fr,
                const SwEnds* __restrict__ rev, TraceParams tp, Arena scratch, uint32_t* __restrict__ cig, unsigned long long cig_cap,
                unsigned long long* __restrict__ cig_used, FinalResult* __restrict__ out, int only_flagged)
{
    extern __shared__ int tsm[];                          // [3][TRACE_SMEM_W][TRACE_BLOCK] row buffers, then the n*n matrix
    int8_t* smat = reinterpret_cast<int8_t*>(tsm + 3 * TRACE_SMEM_W * TRACE_BLOCK);
    const int n = tp.n, gapO = tp.gapO, gapE = tp.gapE;
    for (int q = threadIdx.x; q < n * n; q += TRACE_BLOCK) smat[q] = tp.mat[q];
    __syncthreads();

    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = k < ntasks;
    SwTask tk; tk.out = 0; tk.rd_base = 0; tk.rf_base = 0;
    if (valid) tk = order[k];
    const int i = tk.out;
    FwdResult f; f.want_rev = 0; f.score1 = 0; f.ref_end1 = 0; f.read_end1 = 0; f.word_mode = 0;
    if (valid) f = fr[i];
    if (valid && only_flagged && out[i].status != 8) f.want_rev = -1;      // already done by the narrow-band / warp kernels
    FinalResult r;
    r.ref_begin1 = -1; r.read_begin1 = -1; r.cigar_len = 0; r.status = 0; r.cigar_off = 0;
    if (f.want_rev < 0) { f.want_rev = 0; r = out[i]; }

    bool running = false;                                 // lane has a banded DP to do
    int sub_ref = 1, sub_read = 1;
    if (valid && f.want_rev) {
        if (f.score1 > 0) {
            const SwEnds e = rev[i];
            r.ref_begin1 = f.ref_end1 - e.col;
            r.read_begin1 = f.read_end1 - e.row;
        } else {
            // nothing aligned: the reverse pass of the reference sees an empty (byte mode) or 1x1 (word mode) matrix (ssw.c:820-831)
            r.ref_begin1 = f.word_mode ? 0 : -1;
            r.read_begin1 = 0;
        }
        const bool no_cigar = (7 & tp.flag) == 0 || ((2 & tp.flag) != 0 && f.score1 < tp.filters) ||
            ((4 & tp.flag) != 0 && (f.ref_end1 - r.ref_begin1 > tp.filterd || f.read_end1 - r.read_begin1 > tp.filterd));   // ssw.c:833
        if (!no_cigar) {
            sub_ref = f.ref_end1 - r.ref_begin1 + 1;
            sub_read = f.read_end1 - r.read_begin1 + 1;
            if (f.score1 <= 0) {
                // 1x1 problem, traceback loop never runs: "1M" (ssw.c:625,680-687)
                unsigned long long o = atomicAdd(cig_used, 1ull);
                if (o + 1 > cig_cap) r.status = 6;
                else { cig[o] = 1u << 4; r.cigar_off = (int64_t)o; r.cigar_len = 1; }
            } else running = true;
        }
    }
    const int8_t* ref = seq + tk.rf_base + r.ref_begin1;
    const int8_t* read = seq + tk.rd_base + r.read_begin1;
    const int score = f.score1;

    // ---- banded DP, one cell per lane per iteration
    int bw = abs(sub_ref - sub_read) + 1;
    int width = 0, width_d = 0, maxv = 0;
    uint8_t* dir = nullptr;
    int* hb = tsm; int* hc = tsm; int* eb = tsm; int st = 1;
    int ii = 0, j = 0, end = -1, xi = 0, xp = 0, fv = 0, hleft = 0, hdiag = 0;
    const int8_t* mrow = smat;
    uint8_t* line = nullptr;
    bool new_attempt = running, new_row = false;
    while (__any_sync(0xffffffffu, running)) {
        if (running) {
            if (new_attempt) {
                width = bw * 2 + 3; width_d = bw * 2 + 1;
                const bool in_smem = bw <= TRACE_SMEM_BW;
                const unsigned long long need_dir = ((unsigned long long)width_d * (unsigned long long)sub_read + 15ull) & ~15ull;
                const unsigned long long need = need_dir + (in_smem ? 0ull : 3ull * (unsigned long long)(width + 1) * 4ull);
                const unsigned long long o = atomicAdd(scratch.used, (need + 15ull) & ~15ull);
                if (o + need > scratch.bytes) { r.status = 5; running = false; }
                else {
                    dir = scratch.base + o;
                    // previous-row H, previous-row E, current-row H in band coordinates; shared memory (stride = block) for narrow bands
                    st = in_smem ? TRACE_BLOCK : 1;
                    hb = in_smem ? tsm + threadIdx.x : reinterpret_cast<int*>(dir + need_dir);
                    eb = hb + (in_smem ? TRACE_SMEM_W * TRACE_BLOCK : (width + 1));
                    hc = eb + (in_smem ? TRACE_SMEM_W * TRACE_BLOCK : (width + 1));
                    for (int q = 0; q <= width; ++q) { hb[q * st] = 0; eb[q * st] = 0; hc[q * st] = 0; }
                    ii = 0; new_row = true; new_attempt = false;
                }
            }
            if (running && new_row) {
                const int beg = max(0, ii - bw);
                end = min(sub_ref - 1, ii + bw);
                const int edge = min(end + 1, width - 1);
                xi = band_x(ii, bw); xp = band_x(ii - 1, bw);
                hb[0] = 0; eb[0] = 0; hb[edge * st] = 0; eb[edge * st] = 0; hc[0] = 0;      // ssw.c:580
                fv = 0; hleft = 0; j = beg;
                hdiag = hb[(beg - xp) * st];                                                 // H(ii-1, beg-1), 0 on the matrix edge
                line = dir + (size_t)width_d * (size_t)ii;
                mrow = smat + (int)read[ii];
                new_row = false;
            }
            if (running) {
                const int e_idx = j - xp + 1, u = j - xi + 1;
                const int hup = hb[e_idx * st], eup = eb[e_idx * st];
                int open = ii == 0 ? -gapO : hup - gapO;
                int ext = ii == 0 ? -gapE : eup - gapE;
                const int ev = open > ext ? open : ext;
                const int de3 = open > ext ? 1 : 0;
                open = hleft - gapO; ext = fv - gapE;
                fv = open > ext ? open : ext;
                const int df5 = open > ext ? 1 : 0;
                const int e1 = ev > 0 ? ev : 0, f1 = fv > 0 ? fv : 0;
                const int t1 = e1 > f1 ? e1 : f1;
                const int t2 = hdiag + (int)mrow[(int)ref[j] * n];
                const int hv = t1 > t2 ? t1 : t2;
                int dh;
                if (t1 <= t2) dh = 1; else dh = e1 > f1 ? (de3 ? 3 : 2) : (df5 ? 5 : 4);
                eb[u * st] = ev;                      // e_idx >= u (xp <= xi): the old E of this column was read above
                hc[u * st] = hv;
                if (hv > maxv) maxv = hv;
                line[j - xi] = (uint8_t)(de3 | (df5 << 1) | (dh << 2));
                hleft = hv; hdiag = hup;
                if (++j > end) {
                    // row done: the current row becomes the previous one (the reference copies h_c[1..u] into h_b, ssw.c:612; entries
                    // beyond u are never read before being rewritten or zeroed through `edge`, so swapping the buffers is equivalent)
                    int* tmp = hb; hb = hc; hc = tmp;
                    new_row = true;
                    if (++ii >= sub_read) {
                        bw *= 2;
                        if (maxv < score) new_attempt = true;      // ssw.c:614-615
                        else running = false;
                    }
                }
            }
        }
    }
    bw /= 2;
    if (!(valid && f.want_rev)) { if (valid && !only_flagged) out[i] = r; return; }
    if (r.status != 0 || r.cigar_len == 1 || dir == nullptr) { out[i] = r; return; }

    // ---- traceback, pass 1 counts the CIGAR words, pass 2 writes them back to front
    const long long total = (long long)width_d * sub_read;
    int l = 0;
    unsigned long long coff = 0;
    for (int pass = 0; pass < 2; ++pass) {
        int ii = sub_read - 1, j = sub_ref - 1, state = 2, run = 0, cnt = 0;
        int op = 0, prev = 0;                        // 0 = M, 1 = I, 2 = D (BAM codes)
        while (ii > 0) {
            const long long idx = (long long)width_d * ii + (j - band_x(ii, bw));
            const int cell = (idx >= 0 && idx < total) ? (int)dir[idx] : 0;
            int d;
            if (state == 2) d = cell >> 2; else if (state == 0) d = (cell & 1) ? 3 : 2; else d = (cell & 2) ? 5 : 4;
            if ((cell >> 2) == 0) d = 0;              // unwritten cell
            switch (d) {
                case 1: --ii; --j; state = 2; op = 0; break;
                case 2: --ii; state = 0; op = 1; break;
                case 3: --ii; state = 2; op = 1; break;
                case 4: --j; state = 1; op = 2; break;
                case 5: --j; state = 2; op = 2; break;
                default: r.status = 3; r.cigar_len = 0; out[i] = r; return;
            }
            if (op == prev) ++run;
            else {
                if (pass) cig[coff + (unsigned)(l - 1 - cnt)] = ((uint32_t)run << 4) | (uint32_t)prev;
                ++cnt; prev = op; run = 1;
            }
        }
        if (op == 0) {
            if (pass) cig[coff + (unsigned)(l - 1 - cnt)] = ((uint32_t)(run + 1) << 4);
            ++cnt;
        } else {
            if (pass) { cig[coff + (unsigned)(l - 1 - cnt)] = ((uint32_t)run << 4) | (uint32_t)op; cig[coff + (unsigned)(l - 2 - cnt)] = 1u << 4; }
            cnt += 2;
        }
        if (!pass) {
            l = cnt;
            coff = atomicAdd(cig_used, (unsigned long long)l);
            if (coff + (unsigned long long)l > cig_cap) { r.status = 6; out[i] = r; return; }
        }
    }
    r.cigar_len = l;
    r.cigar_off = (int64_t)coff;
    out[i] = r;
}

}  // namespace mpn
