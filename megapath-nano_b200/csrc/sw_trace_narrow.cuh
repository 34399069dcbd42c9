// Narrow-band traceback + CIGAR kernel (sm_100a): the fast path of banded_sw (ssw.c:532-718) for bands up to
// NARROW_BW (2*bw+1 <= 15 cells per read row), which covers short-read realignment.
// One thread per pair, all 32 lanes advance one DP cell per iteration of a single warp-wide loop.  Compared with the
// generic kernel in sw_trace.cuh nothing in the cell loop touches global memory:
//   * previous-row H / E and current-row H live in shared memory, indexed in band coordinates exactly as the reference
//     indexes h_b / e_b / h_c (ssw.c:585-612), so the `edge` zeroing quirk (ssw.c:580) is reproduced literally;
//   * the target bases under the band sit in a 16-byte register window that slides by one base per row (one global byte
//     load per ROW, issued ~a dozen rows before it is needed);
//   * the substitution scores of the current read base against every target code are packed in one 64-bit register
//     (n <= 8), refreshed once per row from the shared-memory copy of the matrix;
//   * direction flags are 4 bits per cell (E-open, F-open, H source in 2 bits), one aligned 64-bit store per row.
// Pairs whose band has to grow beyond NARROW_BW, or with n > 8, are flagged (status 7) and redone by sw_trace_wide_kernel.
#pragma once
#include "sw_trace.cuh"

namespace mpn {

constexpr int NARROW_BW = 7;
constexpr int NARROW_W = 2 * NARROW_BW + 4;            // row-buffer entries (band coordinates 0 .. 2*bw+2)
constexpr int NARROW_BLOCK = 64;
inline size_t narrow_smem_bytes() { return 3ull * NARROW_W * NARROW_BLOCK * sizeof(int) + 8 * sizeof(unsigned long long); }

// per-pair hand-over from the DP kernel to the traceback kernel
constexpr int NARROW_MAX_ROWS = 1536;       // longer reads go to the warp-parallel kernel whatever their band

struct BandRec {
    unsigned long long dir_off;   // byte offset of the pair's direction words in the scratch arena
    int32_t bw;                   // band half-width of the successful attempt
    int32_t kind;                 // 0: nothing to trace, 2: trace from dir_off
};

// ---- DP: every LANE is a persistent worker that pulls the next pair from a global counter when its current one is done, so the 32
//      lanes of a warp stay busy although band widths (3..15 cells per row) and read lengths differ from pair to pair.
__global__ void __launch_bounds__(NARROW_BLOCK)
sw_band_dp_kernel(const SwTask* __restrict__ order, int ntasks, int* __restrict__ counter, const int8_t* __restrict__ seq,
                  const FwdResult* __restrict__ fr, const SwEnds* __restrict__ rev, TraceParams tp, Arena scratch, uint32_t* __restrict__ cig,
                  unsigned long long cig_cap, unsigned long long* __restrict__ cig_used, FinalResult* __restrict__ out, BandRec* __restrict__ recs, int* __restrict__ flag_list, int* __restrict__ nflag)
{
    extern __shared__ int nsm[];                       // [3][NARROW_W][NARROW_BLOCK] row buffers, then 8 packed score words
    unsigned long long* srow = reinterpret_cast<unsigned long long*>(nsm + 3 * NARROW_W * NARROW_BLOCK);   // srow[q] = bytes mat[t*n+q], t = 0..7
    const int n = tp.n, gapO = tp.gapO, gapE = tp.gapE;
    if (threadIdx.x < 8) {
        unsigned long long v = 0;
        if ((int)threadIdx.x < n && n <= 8)
            for (int t = 0; t < n; ++t) v |= (unsigned long long)(uint8_t)tp.mat[t * n + threadIdx.x] << (8 * t);
        srow[threadIdx.x] = v;
    }
    __syncthreads();
    int* const bufA = nsm + threadIdx.x;
    int* const ebuf = bufA + NARROW_W * NARROW_BLOCK;
    int* const bufB = ebuf + NARROW_W * NARROW_BLOCK;
    constexpr int ST = NARROW_BLOCK;

    int state = 0;                                     // 0: needs a pair, 1: DP in progress, 2: no pairs left
    int kcur = 0;
    int i = 0, sub_ref = 1, sub_read = 1, score = 0, bw = 1, width = 0, maxv = 0;
    const uint8_t* useq = reinterpret_cast<const uint8_t*>(seq);     // unsigned loads: no sign-extension op hanging on the load
    const uint8_t* ref = useq; const uint8_t* read = useq;
    unsigned long long* dirrow = nullptr;
    unsigned long long dir_off = 0;
    int* hb = bufA; int* hc = bufB;
    int ii = 0, j = 0, end = -1, xi = 0, xp = 0, wbase = 0, fv = 0, hleft = 0, hdiag = 0;
    unsigned rnext = 0, pending = 0;
    unsigned long long win_lo = 0, win_hi = 0, rscore = 0, dirword = 0;
    bool new_attempt = false, new_row = false;

    for (;;) {
        if (state == 0) {
            const int k = atomicAdd(counter, 1);
            if (k >= ntasks) state = 2;
            else {
                const SwTask tk = order[k];
                kcur = k;
                i = tk.out;
                const FwdResult f = fr[i];
                FinalResult r;
                r.ref_begin1 = -1; r.read_begin1 = -1; r.cigar_len = 0; r.status = 0; r.cigar_off = 0;
                BandRec br; br.dir_off = 0; br.bw = 0; br.kind = 0;
                if (f.want_rev) {
                    if (f.score1 > 0) {
                        const SwEnds e = rev[i];
                        r.ref_begin1 = f.ref_end1 - e.col;
                        r.read_begin1 = f.read_end1 - e.row;
                    } else {
                        r.ref_begin1 = f.word_mode ? 0 : -1;       // empty / 1x1 reverse matrix (ssw.c:820-831)
                        r.read_begin1 = 0;
                    }
                    const bool no_cigar = (7 & tp.flag) == 0 || ((2 & tp.flag) != 0 && f.score1 < tp.filters) ||
                        ((4 & tp.flag) != 0 && (f.ref_end1 - r.ref_begin1 > tp.filterd || f.read_end1 - r.read_begin1 > tp.filterd));   // ssw.c:833
                    if (!no_cigar) {
                        sub_ref = f.ref_end1 - r.ref_begin1 + 1;
                        sub_read = f.read_end1 - r.read_begin1 + 1;
                        bw = abs(sub_ref - sub_read) + 1;
                        if (f.score1 <= 0) {
                            unsigned long long o = atomicAdd(cig_used, 1ull);          // "1M" (ssw.c:625,680-687)
                            if (o + 1 > cig_cap) r.status = 6;
                            else { cig[o] = 1u << 4; r.cigar_off = (int64_t)o; r.cigar_len = 1; }
                        } else if (n > 8 || bw > NARROW_BW || sub_read > NARROW_MAX_ROWS) {
                            // wide bands, and reads so long that one lane would serialise millions of cells: one warp per pair instead
                            r.status = 7; br.bw = bw; flag_list[atomicAdd(nflag, 1)] = k;
                        }
                        else {
                            ref = useq + tk.rf_base + r.ref_begin1;
                            read = useq + tk.rd_base + r.read_begin1;
                            score = f.score1; maxv = 0;
                            state = 1; new_attempt = true;
                        }
                    }
                }
                out[i] = r;
                if (state != 1) recs[i] = br;
            }
        }
        if (!__any_sync(0xffffffffu, state != 2)) break;
        if (state != 1) continue;

        if (new_attempt) {
            width = bw * 2 + 3;
            const unsigned long long need = (unsigned long long)sub_read * 8ull;
            const unsigned long long o = atomicAdd(scratch.used, (need + 15ull) & ~15ull);
            if (o + need > scratch.bytes) { out[i].status = 5; BandRec br; br.dir_off = 0; br.bw = 0; br.kind = 0; recs[i] = br; state = 0; continue; }
            dir_off = o;
            dirrow = reinterpret_cast<unsigned long long*>(scratch.base + o);
            hb = bufA; hc = bufB;
            for (int q = 0; q <= width; ++q) { bufA[q * ST] = 0; ebuf[q * ST] = 0; bufB[q * ST] = 0; }
            win_lo = 0; win_hi = 0;
            for (int q = 0; q < 16; ++q) {
                const unsigned long long c = q < sub_ref ? (unsigned long long)ref[q] : 0ull;
                if (q < 8) win_lo |= c << (8 * q); else win_hi |= c << (8 * (q - 8));
            }
            pending = 16 < sub_ref ? ref[16] : 0u;                                       // base that enters the window at the next slide
            wbase = 0; ii = 0; rnext = read[0];
            new_row = true; new_attempt = false;
        }
        if (new_row) {
            const int beg = max(0, ii - bw);
            end = min(sub_ref - 1, ii + bw);
            const int edge = min(end + 1, width - 1);
            xi = band_x(ii, bw); xp = band_x(ii - 1, bw);
            hb[0] = 0; ebuf[0] = 0; hb[edge * ST] = 0; ebuf[edge * ST] = 0; hc[0] = 0;      // ssw.c:580
            fv = 0; hleft = 0; j = beg;
            hdiag = hb[(beg - xp) * ST];                                                     // H(ii-1, beg-1); 0 on the matrix edge
            if (beg > wbase) {                                                               // slide the target window by one base
                win_lo = (win_lo >> 8) | (win_hi << 56);
                win_hi = (win_hi >> 8) | ((unsigned long long)pending << 56);
                wbase = beg;
                const int nb = wbase + 16;                                                   // fetched now, needed one row from now
                pending = nb < sub_ref ? ref[nb] : 0u;
            }
            rscore = srow[rnext & 7];                                                        // scores of this read base vs target codes 0..7
            rnext = ii + 1 < sub_read ? read[ii + 1] : 0u;
            dirword = 0;
            new_row = false;
        }
        {
            const int e_idx = j - xp + 1, u = j - xi + 1;
            const int hup = hb[e_idx * ST], eup = ebuf[e_idx * ST];
            const int wpos = j - wbase;
            const unsigned code = (unsigned)((wpos < 8 ? win_lo >> (8 * wpos) : win_hi >> (8 * (wpos - 8))) & 7ull);
            const int sc = (int)(int8_t)(rscore >> (8 * code));
            int open = ii == 0 ? -gapO : hup - gapO;
            int ext = ii == 0 ? -gapE : eup - gapE;
            const int ev = open > ext ? open : ext;
            const unsigned de3 = open > ext ? 1u : 0u;                                        // ties extend (ssw.c:593-594)
            open = hleft - gapO; ext = fv - gapE;
            fv = open > ext ? open : ext;
            const unsigned df5 = open > ext ? 1u : 0u;                                        // ssw.c:596-599
            const int e1 = ev > 0 ? ev : 0, f1 = fv > 0 ? fv : 0;
            const int t1 = e1 > f1 ? e1 : f1;
            const int t2 = hdiag + sc;
            const int hv = t1 > t2 ? t1 : t2;
            const unsigned src = t1 <= t2 ? 1u : (e1 > f1 ? 2u : 3u);                         // 1 diagonal, 2 from E, 3 from F (ssw.c:609-610)
            ebuf[u * ST] = ev;                                                                // e_idx >= u: the old E of this column was read above
            hc[u * ST] = hv;
            if (hv > maxv) maxv = hv;
            dirword |= (unsigned long long)(de3 | (df5 << 1) | (src << 2)) << (4 * (j - xi));
            hleft = hv; hdiag = hup;
            if (++j > end) {
                dirrow[ii] = dirword;
                int* tmp = hb; hb = hc; hc = tmp;           // ssw.c:612 (see sw_trace.cuh for why a swap is equivalent)
                new_row = true;
                if (++ii >= sub_read) {
                    if (maxv >= score) {                                                      // ssw.c:614-615
                        BandRec br; br.dir_off = dir_off; br.bw = bw; br.kind = 2; recs[i] = br;
                        state = 0;
                    } else {
                        bw *= 2;
                        if (bw > NARROW_BW) { out[i].status = 7; BandRec br; br.dir_off = 0; br.bw = bw; br.kind = 0; recs[i] = br; flag_list[atomicAdd(nflag, 1)] = kcur; state = 0; }
                        else new_attempt = true;
                    }
                }
            }
        }
    }
}

// ---- traceback (ssw.c:618-697) from the packed direction words: one thread per pair
__global__ void __launch_bounds__(128)
sw_band_trace_kernel(const SwTask* __restrict__ order, int ntasks, const FwdResult* __restrict__ fr, const BandRec* __restrict__ recs, Arena scratch,
                     uint32_t* __restrict__ cig, unsigned long long cig_cap, unsigned long long* __restrict__ cig_used, FinalResult* __restrict__ out)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= ntasks) return;
    const int i = order[k].out;
    const BandRec br = recs[i];
    if (br.kind != 2) return;
    const FwdResult f = fr[i];
    FinalResult r = out[i];
    const int sub_ref = f.ref_end1 - r.ref_begin1 + 1, sub_read = f.read_end1 - r.read_begin1 + 1;
    const int bw = br.bw, width_d = 2 * bw + 1;
    const unsigned long long* dirrow = reinterpret_cast<const unsigned long long*>(scratch.base + br.dir_off);
    int l = 0;
    unsigned long long coff = 0;
    for (int pass = 0; pass < 2; ++pass) {
        int ti = sub_read - 1, tj = sub_ref - 1, state = 2, run = 0, cnt = 0;
        int op = 0, prev = 0;                               // BAM codes: 0 = M, 1 = I, 2 = D
        int cur_row = ti;
        unsigned long long cur = dirrow[ti], nxt = ti > 0 ? dirrow[ti - 1] : 0ull;      // row ti and, prefetched, row ti-1
        while (ti > 0) {
            // the reference indexes a flat array; cells left or right of the band alias into the neighbouring rows
            const int cpos = tj - band_x(ti, bw);
            unsigned long long word;
            int cc = cpos;
            if (cpos >= 0 && cpos < width_d) word = cur;
            else {
                const long long lin = (long long)width_d * ti + cpos;
                if (lin >= 0 && lin < (long long)width_d * sub_read) { word = dirrow[lin / width_d]; cc = (int)(lin % width_d); }
                else { word = 0; cc = 0; }
            }
            const int cell = (int)((word >> (4 * cc)) & 15ull);
            const int src = cell >> 2;
            int d;
            if (src == 0) d = 0;
            else if (state == 2) d = src == 1 ? 1 : (src == 2 ? ((cell & 1) ? 3 : 2) : ((cell & 2) ? 5 : 4));
            else if (state == 0) d = (cell & 1) ? 3 : 2;
            else d = (cell & 2) ? 5 : 4;
            switch (d) {
                case 1: --ti; --tj; state = 2; op = 0; break;
                case 2: --ti; state = 0; op = 1; break;
                case 3: --ti; state = 2; op = 1; break;
                case 4: --tj; state = 1; op = 2; break;
                case 5: --tj; state = 2; op = 2; break;
                default: r.status = 3; r.cigar_len = 0; out[i] = r; return;
            }
            if (ti != cur_row) { cur = nxt; cur_row = ti; nxt = ti > 0 ? dirrow[ti - 1] : 0ull; }
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
