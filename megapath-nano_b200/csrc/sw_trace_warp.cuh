// Warp-parallel banded traceback DP (sm_100a) for the pairs the narrow-band kernel hands back (band wider than 15 cells, or
// an alphabet with n > 8): one warp per pair, the band cells of a read row spread over the lanes, rows in sequence.
// Same recurrence, tie rules and band-coordinate quirks as banded_sw (ssw.c:583-612), evaluated row-parallel:
//   * lane L owns CPL consecutive band cells (CPL odd: conflict-free shared-memory stride); E and the diagonal term come from the
//     previous row (shared memory, band coordinates of that row);
//   * the left-to-right F chain  f(p) = max(H(p-1) - gapO, f(p-1) - gapE)  is a max-plus prefix scan: with
//     A(p) = max(E+(p), Hdiag + s) (the F-free part of H) and gm = min(gapO, gapE),
//     f(p) = max(-gm (p+1), max_{k<p} (A(k) - gapO + gm (k+1)) - gm p), exact in int32.  Each lane scans its own cells serially and
//     one warp scan of the lane aggregates supplies the carry; the open/extend flag of F is then recomputed from the actual
//     neighbours (ssw.c:596-599), so ties resolve exactly as in the scalar loop;
//   * direction bytes (layout of sw_trace_wide_kernel) go to a region owned by the warp and reused for every band attempt and
//     every pair, so ONT-scale pairs (10 kb rows x hundreds of band cells, several doubling attempts) cannot exhaust an arena;
//   * the traceback walk (ssw.c:618-697) is done by all lanes in lock-step from a shared-memory window that the lanes fill together,
//     one of the next 32 rows each, so that the DRAM latency of the direction bytes is paid once per 32 rows instead of once per step.
#pragma once
#include "sw_trace_narrow.cuh"

namespace mpn {

constexpr int WARPTR_MAXCPL = 17;                       // band cells <= 32 * 17 = 544  (bw <= 271)
constexpr int WARPTR_CELLS = 32 * WARPTR_MAXCPL;
constexpr int WARPTR_WARPS = 2;                         // warps per block
inline size_t warptr_smem_bytes(int n) { return (size_t)WARPTR_WARPS * 4 * (WARPTR_CELLS + 2) * sizeof(int) + (((size_t)n * n + 15) & ~(size_t)15); }
// bytes of direction storage one warp needs for reads up to max_rows rows, plus room behind them for the CIGAR words of the walk
// (at most one word per read row and per target column of the sub-rectangle: 3 words per row covers any band the kernel takes)
inline size_t warptr_region_bytes(int max_rows) { return (((size_t)(WARPTR_CELLS + 12) * (size_t)max_rows) + 4096 + 255) & ~(size_t)255; }

struct WarpRowCtx {
    int* Hp; int* Ep; int* Hc; int* Ec;
    const int8_t* smat; const int8_t* ref; const int8_t* read;
    int n, gapO, gapE, gm, bw, width, width_d, sub_ref, lane;
};

// one read row of the band, CPL cells per lane.  Returns the largest H of this lane's cells.
template <int CPL>
__device__ __forceinline__ int warp_band_row(const WarpRowCtx& c, int ii, uint8_t* __restrict__ line)
{
    const int NEG = -(1 << 29);
    const int end = min(c.sub_ref - 1, ii + c.bw);
    const int edge = min(end + 1, c.width - 1);
    const int xi = band_x(ii, c.bw), xp = band_x(ii - 1, c.bw);
    const int shift = xi - xp;
    const int8_t* mrow = c.smat + (int)c.read[ii];
    const int p0 = c.lane * CPL;
    int e1s[CPL], t2s[CPL], evs[CPL];
    unsigned de3s = 0;
    int ymax = NEG;                                                   // running max of y over this lane's cells
    // ---- phase 1: everything that does not depend on F
#pragma unroll
    for (int k = 0; k < CPL; ++k) {
        const int p = p0 + k, j = xi + p;
        const bool valid = p < c.width_d && j <= end;
        const int e_idx = min(p + shift + 1, WARPTR_CELLS + 1);        // previous-row band coordinate of (ii-1, j)
        int hup = c.Hp[e_idx], eup = c.Ep[e_idx];
        if (e_idx == edge) { hup = 0; eup = 0; }                       // the slot the reference zeroes before every row
        const int hdg = (e_idx - 1 == 0 || e_idx - 1 == edge) ? 0 : c.Hp[e_idx - 1];   // h_b[0] = 0, and the diagonal sees the zeroed slot too
        const int open = ii == 0 ? -c.gapO : hup - c.gapO, ext = ii == 0 ? -c.gapE : eup - c.gapE;
        const int ev = open > ext ? open : ext;
        de3s |= (open > ext ? 1u : 0u) << k;
        const int e1 = ev > 0 ? ev : 0;
        const int sc = valid ? (int)mrow[(int)c.ref[j] * c.n] : 0;
        const int t2 = hdg + sc;
        const int A = e1 > t2 ? e1 : t2;
        evs[k] = ev; e1s[k] = e1; t2s[k] = t2;
        const int y = valid ? A - c.gapO + c.gm * (p + 1) : NEG;
        ymax = max(ymax, y);
    }
    // ---- exclusive prefix max of the lane aggregates
    int inc = ymax;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int o2 = __shfl_up_sync(0xffffffffu, inc, off);
        if (c.lane >= off) inc = max(inc, o2);
    }
    int carry = __shfl_up_sync(0xffffffffu, inc, 1);
    if (c.lane == 0) carry = NEG;
    // ---- phase 2: F, H, directions.  The first cell of a lane needs H and f of the previous lane's last cell for its open/extend
    //      flag: that byte is completed after the loop.
    int run = carry, hprev = 0, fprev = 0, maxh = 0;
    int h_first = 0, f_first = 0; unsigned first_bits = 0; bool first_valid = false; bool first_needs_f = false;
#pragma unroll
    for (int k = 0; k < CPL; ++k) {
        const int p = p0 + k, j = xi + p;
        const bool valid = p < c.width_d && j <= end;
        const int fv = max(-c.gm * (p + 1), run - c.gm * p);
        const int f1 = fv > 0 ? fv : 0;
        const int e1 = e1s[k], t2 = t2s[k];
        const int t1 = e1 > f1 ? e1 : f1;
        const int hv = t1 > t2 ? t1 : t2;
        const unsigned de3 = (de3s >> k) & 1u;
        if (k == 0) {
            h_first = hv; f_first = fv; first_valid = valid;
            first_needs_f = !(t1 <= t2) && !(e1 > f1);
            unsigned dh = 0;
            if (t1 <= t2) dh = 1; else if (e1 > f1) dh = de3 ? 3 : 2;
            first_bits = de3 | (dh << 2);
        } else {
            const unsigned df5 = (hprev - c.gapO > fprev - c.gapE) ? 1u : 0u;
            unsigned dh;
            if (t1 <= t2) dh = 1; else dh = e1 > f1 ? (de3 ? 3u : 2u) : (df5 ? 5u : 4u);
            if (valid) line[p] = (uint8_t)(de3 | (df5 << 1) | (dh << 2));
        }
        if (valid && hv > maxh) maxh = hv;
        c.Hc[p + 1] = valid ? hv : 0;
        c.Ec[p + 1] = valid ? evs[k] : 0;
        const int y = valid ? max(e1, t2) - c.gapO + c.gm * (p + 1) : NEG;
        run = max(run, y);
        hprev = hv; fprev = fv;
    }
    // ---- first cell of every lane: neighbours are the previous lane's last cell (lane 0: H = 0, f = 0, ssw.c:580)
    int hl = __shfl_up_sync(0xffffffffu, hprev, 1), fl = __shfl_up_sync(0xffffffffu, fprev, 1);
    if (c.lane == 0) { hl = 0; fl = 0; }
    (void)h_first; (void)f_first;
    if (first_valid) {
        const unsigned df5 = (hl - c.gapO > fl - c.gapE) ? 1u : 0u;
        unsigned bits = first_bits | (df5 << 1);
        if (first_needs_f) bits |= (df5 ? 5u : 4u) << 2;
        line[p0] = (uint8_t)bits;
    }
    return maxh;
}

__global__ void __launch_bounds__(32 * WARPTR_WARPS)
sw_trace_warp_kernel(const SwTask* __restrict__ order, const int* __restrict__ flag_list, const int* __restrict__ nflag, int* __restrict__ cursor,
                     const int8_t* __restrict__ seq, const FwdResult* __restrict__ fr, TraceParams tp, uint8_t* __restrict__ dir_base, unsigned long long dir_stride,
                     uint32_t* __restrict__ cig, unsigned long long cig_cap, unsigned long long* __restrict__ cig_used, FinalResult* __restrict__ out,
                     const BandRec* __restrict__ recs)
{
    extern __shared__ int wsmem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    constexpr int ROW = WARPTR_CELLS + 2;               // index 0 = band coordinate 0 (left boundary), cell p lives at p + 1
    WarpRowCtx c;
    c.Hp = wsmem + wid * 4 * ROW; c.Ep = c.Hp + ROW; c.Hc = c.Ep + ROW; c.Ec = c.Hc + ROW;
    int8_t* smat = reinterpret_cast<int8_t*>(wsmem + WARPTR_WARPS * 4 * ROW);
    c.n = tp.n; c.gapO = tp.gapO; c.gapE = tp.gapE; c.gm = min(tp.gapO, tp.gapE); c.lane = lane; c.smat = smat;
    for (int q = threadIdx.x; q < c.n * c.n; q += blockDim.x) smat[q] = tp.mat[q];
    __syncthreads();
    uint8_t* const dir = dir_base + (unsigned long long)(blockIdx.x * WARPTR_WARPS + wid) * dir_stride;

    for (;;) {
        int li = 0;
        if (lane == 0) li = atomicAdd(cursor, 1);
        li = __shfl_sync(0xffffffffu, li, 0);
        if (li >= *nflag) break;
        const SwTask tk = order[flag_list[li]];
        const int i = tk.out;
        const FwdResult f = fr[i];
        FinalResult r = out[i];
        const int sub_ref = f.ref_end1 - r.ref_begin1 + 1, sub_read = f.read_end1 - r.read_begin1 + 1, score = f.score1;
        c.ref = seq + tk.rf_base + r.ref_begin1;
        c.read = seq + tk.rd_base + r.read_begin1;
        c.sub_ref = sub_ref;
        int bw = recs[i].bw;                            // first band of the doubling sequence that the narrow kernel did not try
        int maxv = 0, width_d = 0;
        bool fail = false;
        do {
            c.bw = bw; c.width = 2 * bw + 3; c.width_d = width_d = 2 * bw + 1;
            if (width_d > WARPTR_CELLS || (unsigned long long)width_d * (unsigned long long)sub_read > dir_stride) { r.status = 8; fail = true; break; }   // left to the generic kernel
            for (int q = lane; q < ROW; q += 32) { c.Hp[q] = 0; c.Ep[q] = 0; c.Hc[q] = 0; c.Ec[q] = 0; }
            __syncwarp();
            const int cpl = (width_d + 31) >> 5;
            for (int ii = 0; ii < sub_read; ++ii) {
                uint8_t* line = dir + (size_t)width_d * (size_t)ii;
                int mh;
                if (cpl <= 3) mh = warp_band_row<3>(c, ii, line);
                else if (cpl <= 5) mh = warp_band_row<5>(c, ii, line);
                else if (cpl <= 9) mh = warp_band_row<9>(c, ii, line);
                else mh = warp_band_row<WARPTR_MAXCPL>(c, ii, line);
                maxv = max(maxv, mh);
                __syncwarp();
                int* t = c.Hp; c.Hp = c.Hc; c.Hc = t; t = c.Ep; c.Ep = c.Ec; c.Ec = t;
            }
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) maxv = max(maxv, __shfl_xor_sync(0xffffffffu, maxv, off));
            bw *= 2;
        } while (maxv < score);
        bw /= 2;
        if (fail) { if (lane == 0) out[i] = r; continue; }
        __syncwarp();
        {
            // ---- traceback (ssw.c:618-697).  Every lane walks the same path.  By now the direction bytes of a long read have left the L2 (a
            //      batch writes gigabytes of them), so a walk that loads one byte per step waits for DRAM ten thousand times.  Instead the warp
            //      loads a WINDOW -- for each of the next 32 rows the 48 bytes around the column the path would reach on a pure diagonal, one row
            //      per lane, all in flight together -- into shared memory (the row buffers of the DP are free now) and walks from there; a cell
            //      outside its row's window (more than ~16 net indels within 32 rows, or the aliasing of ssw.c's flat index) is read from memory.
            //      The CIGAR comes out back to front and its length only at the end: the words go to the unused tail of the warp's direction
            //      region and are copied, reversed, into the arena once the length is known (one walk instead of count + write).
            const long long total = (long long)width_d * sub_read;
            const long long tail = (total + 15) & ~15ll;
            uint32_t* const tmp = reinterpret_cast<uint32_t*>(dir + tail);
            const long long tmp_cap = ((long long)dir_stride - tail) / 4;
            uint8_t* const win = reinterpret_cast<uint8_t*>(c.Hp < c.Hc ? c.Hp : c.Hc);        // 32 rows x 48 bytes
            long long* const wbase = reinterpret_cast<long long*>(win + 32 * 48);               // 32 window start indices
            int l = 0; unsigned long long coff = 0; bool bad = false;
            int ti = sub_read - 1, tj = sub_ref - 1, state = 2, run = 0, cnt = 0, op = 0, prev = 0;
            int top = -1;                                       // the window holds rows top, top - 1, ..., top - 31
            auto emit = [&](uint32_t word) {
                if ((long long)cnt >= tmp_cap) bad = true;
                else if (lane == 0) tmp[cnt] = word;
                ++cnt;
            };
            while (ti > 0 && !bad) {
                if (top < 0 || ti < top - 31) {
                    __syncwarp();
                    top = ti;
                    const int rr = ti - lane;
                    long long sb = -1;
                    if (rr >= 0) {
                        const int pc = tj - (ti - rr) - band_x(rr, bw);          // band coordinate if the path stays on the diagonal
                        long long a0 = (long long)width_d * rr + pc - 16;
                        a0 = a0 < 0 ? 0 : a0;
                        sb = a0 & ~15ll;
                        if (sb + 48 > (long long)dir_stride) sb = ((long long)dir_stride - 48) & ~15ll;
                        const uint4* src = reinterpret_cast<const uint4*>(dir + sb);
                        uint4* dst = reinterpret_cast<uint4*>(win + lane * 48);
                        const uint4 v0 = src[0], v1 = src[1], v2 = src[2];
                        dst[0] = v0; dst[1] = v1; dst[2] = v2;
                    }
                    wbase[lane] = sb;
                    __syncwarp();
                }
                const int slot = top - ti;
                const long long idx = (long long)width_d * ti + (tj - band_x(ti, bw));
                int cell = 0;
                if (idx >= 0 && idx < total) {
                    const long long o = idx - wbase[slot];
                    cell = (o >= 0 && o < 48) ? (int)win[slot * 48 + (int)o] : (int)dir[idx];
                }
                int d;
                if (state == 2) d = cell >> 2; else if (state == 0) d = (cell & 1) ? 3 : 2; else d = (cell & 2) ? 5 : 4;
                if ((cell >> 2) == 0) d = 0;
                if (d == 1) { --ti; --tj; state = 2; op = 0; }
                else if (d == 2) { --ti; state = 0; op = 1; }
                else if (d == 3) { --ti; state = 2; op = 1; }
                else if (d == 4) { --tj; state = 1; op = 2; }
                else if (d == 5) { --tj; state = 2; op = 2; }
                else { r.status = 3; r.cigar_len = 0; bad = true; break; }
                if (op == prev) ++run;
                else { emit(((uint32_t)run << 4) | (uint32_t)prev); prev = op; run = 1; }
            }
            if (!bad) {
                if (op == 0) emit((uint32_t)(run + 1) << 4);
                else { emit(((uint32_t)run << 4) | (uint32_t)op); emit(1u << 4); }
            }
            if (bad && r.status != 3) r.status = 8;             // no room for the words behind the direction bytes: left to the generic kernel
            if (!bad) {
                l = cnt;
                if (lane == 0) coff = atomicAdd(cig_used, (unsigned long long)l);
                coff = __shfl_sync(0xffffffffu, coff, 0);
                if (coff + (unsigned long long)l > cig_cap) { r.status = 6; bad = true; }
            }
            __syncwarp();
            if (!bad) {
                for (int k = lane; k < l; k += 32) cig[coff + (unsigned)k] = tmp[l - 1 - k];
                r.status = 0; r.cigar_len = l; r.cigar_off = (int64_t)coff;
            }
            if (lane == 0) out[i] = r;
        }
        __syncwarp();
    }
}

}  // namespace mpn
