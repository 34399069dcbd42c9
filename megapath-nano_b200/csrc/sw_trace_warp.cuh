// Warp-parallel banded traceback DP (sm_100a) for the pairs the narrow-band kernel hands back (band wider than 15 cells, or
// an alphabet with n > 8): one warp per pair, lanes across the band cells of a read row, rows in sequence.
// Same recurrence, tie rules and band-coordinate quirks as banded_sw (ssw.c:583-612), evaluated row-parallel:
//   * E and the diagonal term come from the previous row (shared memory, band coordinates of that row);
//   * the left-to-right F chain  f(p) = max(H(p-1) - gapO, f(p-1) - gapE)  is a max-plus prefix scan: with
//     A(p) = max(E+(p), Hdiag + s) (the F-free part of H) and gm = min(gapO, gapE),
//     f(p) = max(-gm (p+1), max_{k<p} (A(k) - gapO + gm (k+1)) - gm p), exact in int32; the open/extend flag of F is then
//     recomputed from the actual neighbours (ssw.c:596-599), so ties resolve exactly as in the scalar loop;
//   * direction bytes use the layout of sw_trace_wide_kernel, whose traceback code finishes the job (lane 0).
#pragma once
#include "sw_trace_narrow.cuh"

namespace mpn {

constexpr int WARPTR_CHUNKS = 16;                       // band cells <= 32 * 16 = 512  (bw <= 255)
constexpr int WARPTR_CELLS = 32 * WARPTR_CHUNKS;
constexpr int WARPTR_WARPS = 2;                         // warps per block
inline size_t warptr_smem_bytes(int n) { return (size_t)WARPTR_WARPS * 4 * (WARPTR_CELLS + 2) * sizeof(int) + (((size_t)n * n + 15) & ~(size_t)15); }

__global__ void __launch_bounds__(32 * WARPTR_WARPS)
sw_trace_warp_kernel(const SwTask* __restrict__ order, const int* __restrict__ flag_list, const int* __restrict__ nflag, int* __restrict__ cursor,
                     const int8_t* __restrict__ seq, const FwdResult* __restrict__ fr, TraceParams tp, Arena scratch, uint32_t* __restrict__ cig,
                     unsigned long long cig_cap, unsigned long long* __restrict__ cig_used, FinalResult* __restrict__ out, const BandRec* __restrict__ recs)
{
    extern __shared__ int wsmem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    constexpr int ROW = WARPTR_CELLS + 2;               // index 0 = band coordinate 0 (left boundary), cell p lives at p + 1
    int* Hp = wsmem + wid * 4 * ROW; int* Ep = Hp + ROW; int* Hc = Ep + ROW; int* Ec = Hc + ROW;
    int8_t* smat = reinterpret_cast<int8_t*>(wsmem + WARPTR_WARPS * 4 * ROW);
    const int n = tp.n, gapO = tp.gapO, gapE = tp.gapE, gm = min(gapO, gapE);
    for (int q = threadIdx.x; q < n * n; q += blockDim.x) smat[q] = tp.mat[q];
    __syncthreads();
    const int NEG = -(1 << 29);

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
        const int8_t* ref = seq + tk.rf_base + r.ref_begin1;
        const int8_t* read = seq + tk.rd_base + r.read_begin1;
        int bw = recs[i].bw;                            // first band of the doubling sequence that the narrow kernel did not try
        int maxv = 0, width = 0, width_d = 0;
        uint8_t* dir = nullptr;
        bool fail = false;
        do {
            width = 2 * bw + 3; width_d = 2 * bw + 1;
            if (width_d > WARPTR_CELLS) { r.status = 8; fail = true; break; }       // left to the generic kernel
            unsigned long long o = 0;
            const unsigned long long need = ((unsigned long long)width_d * (unsigned long long)sub_read + 15ull) & ~15ull;
            if (lane == 0) o = atomicAdd(scratch.used, need);
            o = __shfl_sync(0xffffffffu, o, 0);
            if (o + need > scratch.bytes) { r.status = 5; fail = true; break; }
            dir = scratch.base + o;
            for (int q = lane; q < ROW; q += 32) { Hp[q] = 0; Ep[q] = 0; Hc[q] = 0; Ec[q] = 0; }
            __syncwarp();
            const int chunks = (width_d + 31) >> 5;
            for (int ii = 0; ii < sub_read; ++ii) {
                const int end = min(sub_ref - 1, ii + bw);
                const int edge = min(end + 1, width - 1);
                const int xi = band_x(ii, bw), xp = band_x(ii - 1, bw);
                const int shift = xi - xp;
                const int8_t* mrow = smat + (int)read[ii];
                int pm_carry = NEG, hv_carry = 0, fv_carry = 0;                    // left neighbour of the first cell: H = 0, f = 0 (ssw.c:580)
                uint8_t* line = dir + (size_t)width_d * (size_t)ii;
                for (int c = 0; c < chunks; ++c) {
                    const int p = 32 * c + lane, j = xi + p;
                    const bool valid = p < width_d && j <= end;
                    const int e_idx = p + shift + 1;                               // previous-row band coordinate of (ii-1, j)
                    int hup = Hp[e_idx], eup = Ep[e_idx];
                    if (e_idx == edge) { hup = 0; eup = 0; }                       // the slot the reference zeroes before every row
                    const int hdg = (e_idx - 1 == 0) ? 0 : Hp[e_idx - 1];          // h_b[0] = 0
                    int open = ii == 0 ? -gapO : hup - gapO, ext = ii == 0 ? -gapE : eup - gapE;
                    const int ev = open > ext ? open : ext;
                    const int de3 = open > ext ? 1 : 0;
                    const int e1 = ev > 0 ? ev : 0;
                    const int sc = valid ? (int)mrow[(int)ref[j] * n] : 0;
                    const int t2 = hdg + sc;
                    const int A = e1 > t2 ? e1 : t2;
                    int y = valid ? A - gapO + gm * (p + 1) : NEG;
#pragma unroll
                    for (int off = 1; off < 32; off <<= 1) {
                        const int o2 = __shfl_up_sync(0xffffffffu, y, off);
                        if (lane >= off) y = max(y, o2);
                    }
                    y = max(y, pm_carry);                                          // inclusive prefix max over cells 0..p
                    int yex = __shfl_up_sync(0xffffffffu, y, 1);
                    if (lane == 0) yex = pm_carry;
                    const int fv = max(-gm * (p + 1), yex - gm * p);
                    const int f1 = fv > 0 ? fv : 0;
                    const int t1 = e1 > f1 ? e1 : f1;
                    const int hv = t1 > t2 ? t1 : t2;
                    int hl = __shfl_up_sync(0xffffffffu, hv, 1), fl = __shfl_up_sync(0xffffffffu, fv, 1);
                    if (lane == 0) { hl = hv_carry; fl = fv_carry; }
                    const int df5 = (hl - gapO > fl - gapE) ? 1 : 0;
                    int dh;
                    if (t1 <= t2) dh = 1; else dh = e1 > f1 ? (de3 ? 3 : 2) : (df5 ? 5 : 4);
                    if (valid) {
                        line[p] = (uint8_t)(de3 | (df5 << 1) | (dh << 2));
                        if (hv > maxv) maxv = hv;
                    }
                    Hc[p + 1] = valid ? hv : 0;
                    Ec[p + 1] = valid ? ev : 0;
                    pm_carry = __shfl_sync(0xffffffffu, y, 31);
                    hv_carry = __shfl_sync(0xffffffffu, hv, 31);
                    fv_carry = __shfl_sync(0xffffffffu, fv, 31);
                }
                __syncwarp();
                int* t = Hp; Hp = Hc; Hc = t; t = Ep; Ep = Ec; Ec = t;
            }
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) maxv = max(maxv, __shfl_xor_sync(0xffffffffu, maxv, off));
            bw *= 2;
        } while (maxv < score);
        bw /= 2;
        if (fail) { if (lane == 0) out[i] = r; continue; }
        __syncwarp();
        if (lane == 0) {
            // ---- traceback (ssw.c:618-697), pass 0 counts, pass 1 writes back to front
            const long long total = (long long)width_d * sub_read;
            int l = 0; unsigned long long coff = 0; bool bad = false;
            for (int pass = 0; pass < 2 && !bad; ++pass) {
                int ti = sub_read - 1, tj = sub_ref - 1, state = 2, run = 0, cnt = 0, op = 0, prev = 0;
                while (ti > 0) {
                    const long long idx = (long long)width_d * ti + (tj - band_x(ti, bw));
                    const int cell = (idx >= 0 && idx < total) ? (int)dir[idx] : 0;
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
                    else {
                        if (pass) cig[coff + (unsigned)(l - 1 - cnt)] = ((uint32_t)run << 4) | (uint32_t)prev;
                        ++cnt; prev = op; run = 1;
                    }
                }
                if (bad) break;
                if (op == 0) { if (pass) cig[coff + (unsigned)(l - 1 - cnt)] = ((uint32_t)(run + 1) << 4); ++cnt; }
                else {
                    if (pass) { cig[coff + (unsigned)(l - 1 - cnt)] = ((uint32_t)run << 4) | (uint32_t)op; cig[coff + (unsigned)(l - 2 - cnt)] = 1u << 4; }
                    cnt += 2;
                }
                if (!pass) {
                    l = cnt;
                    coff = atomicAdd(cig_used, (unsigned long long)l);
                    if (coff + (unsigned long long)l > cig_cap) { r.status = 6; bad = true; }
                }
            }
            if (!bad) { r.status = 0; r.cigar_len = l; r.cigar_off = (int64_t)coff; }
            out[i] = r;
        }
        __syncwarp();
    }
}

}  // namespace mpn
