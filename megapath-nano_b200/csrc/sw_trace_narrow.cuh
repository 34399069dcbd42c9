// Narrow-band traceback walk (sm_100a): the CIGAR half of banded_sw (ssw.c:618-718) for bands of up to 2 * NARROW_BW + 1 = 15
// cells per read row, which covers short-read realignment.  The banded DP that produces the
// direction words lives in sw_trace_rows.cuh (band row in registers, 4 bits per cell, one 64-bit word per read row); this file
// holds the shared constants, the per-pair hand-over record and the one-thread-per-pair walk over those words.
// Pairs whose band has to grow beyond NARROW_BW, whose reads are longer than NARROW_MAX_ROWS, or with n > 8 are flagged
// (status 7) and done by sw_trace_warp_kernel.
#pragma once
#include "sw_trace.cuh"

namespace mpn {

constexpr int NARROW_BW = 7;
constexpr int NARROW_MAX_ROWS = 1536;       // longer reads go to the warp-parallel kernel whatever their band

// bytes of one row's direction word: 4 bits per band cell, 2 bw + 1 cells -> 16 bits up to bw = 1, 32 bits up to bw = 3, else 64 bits
// (the path walk is bound by these reads: one thread per pair, 32-byte sectors of different pairs)
__host__ __device__ constexpr int dir_row_bytes(int bw) { return bw <= 1 ? 2 : (bw <= 3 ? 4 : 8); }

// per-pair hand-over from the DP kernel to the traceback kernel

struct BandRec {
    unsigned long long dir_off;   // byte offset of the pair's direction words in the scratch arena
    int32_t bw;                   // band half-width of the successful attempt
    int32_t kind;                 // 0: nothing to trace, 2: trace from dir_off
};

// ---- traceback (ssw.c:618-697) from the packed direction words: one thread per pair
__global__ void __launch_bounds__(128)
sw_band_trace_kernel(const SwTask* __restrict__ order, int ntasks, const FwdResult* __restrict__ fr, const BandRec* __restrict__ recs, Arena scratch,
                     uint32_t* __restrict__ cig, unsigned long long cig_cap, unsigned long long* __restrict__ cig_used, FinalResult* __restrict__ out)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = k < ntasks;
    const int i = order[live ? k : 0].out;
    BandRec br = recs[i];
    if (!live) br.kind = 0;
    // whole warps stay until the CIGAR arena has been claimed: one atomic per warp for the sum of its lengths
    const bool walk = br.kind == 2;
    const FwdResult f = fr[i];
    FinalResult r = out[i];
    const int sub_ref = f.ref_end1 - r.ref_begin1 + 1, sub_read = f.read_end1 - r.read_begin1 + 1;
    const int bw = br.bw, width_d = 2 * bw + 1;
    const uint8_t* const dirbase = scratch.base + br.dir_off;
    const int rowb = dir_row_bytes(bw);
    auto row_word = [&](int r) -> unsigned long long {
        if (rowb == 2) return reinterpret_cast<const uint16_t*>(dirbase)[r];
        if (rowb == 4) return reinterpret_cast<const uint32_t*>(dirbase)[r];
        return reinterpret_cast<const unsigned long long*>(dirbase)[r];
    };
    int l = 0;
    unsigned long long coff = 0;
    // The walk yields the CIGAR back to front and its length is only known at the end.  Pass 0 counts and keeps the first SHORT words
    // in a local buffer: a short-read CIGAR (a handful of runs) is then written from the buffer and the second walk is skipped.
    constexpr int SHORT = 24;
    uint32_t buf[SHORT];
    bool failed = false;
    for (int pass = 0; pass < 2; ++pass) {
        if (pass == 1 && (!walk || failed || l <= SHORT)) break;
        int ti = walk && !failed ? sub_read - 1 : 0, tj = sub_ref - 1, state = 2, run = 0, cnt = 0;
        int op = 0, prev = 0;                               // BAM codes: 0 = M, 1 = I, 2 = D
        int cur_row = ti;
        unsigned long long cur = walk ? row_word(ti) : 0ull, nxt = ti > 0 ? row_word(ti - 1) : 0ull;      // row ti and, prefetched, row ti-1
        auto emit = [&](uint32_t word) {
            if (pass) cig[coff + (unsigned)(l - 1 - cnt)] = word;
            else if (cnt < SHORT) buf[cnt] = word;
            ++cnt;
        };
        while (ti > 0) {
            // the reference indexes a flat array; cells left or right of the band alias into the neighbouring rows
            const int cpos = tj - band_x(ti, bw);
            unsigned long long word;
            int cc = cpos;
            if (cpos >= 0 && cpos < width_d) word = cur;
            else {
                const long long lin = (long long)width_d * ti + cpos;
                if (lin >= 0 && lin < (long long)width_d * sub_read) { word = row_word((int)(lin / width_d)); cc = (int)(lin % width_d); }
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
                default: failed = true; ti = 0; break;          // invalid direction: the reference returns NULL (ssw.c:657-665, 840-843)
            }
            if (failed) break;
            if (ti != cur_row) { cur = nxt; cur_row = ti; nxt = ti > 0 ? row_word(ti - 1) : 0ull; }
            if (op == prev) ++run;
            else { emit(((uint32_t)run << 4) | (uint32_t)prev); prev = op; run = 1; }
        }
        if (walk && !failed) {
            if (op == 0) emit((uint32_t)(run + 1) << 4);
            else { emit(((uint32_t)run << 4) | (uint32_t)op); emit(1u << 4); }
        }
        if (!pass) {
            l = walk && !failed ? cnt : 0;
            // warp-aggregated claim: inclusive scan of the lengths, one atomicAdd by the last lane, base broadcast
            const int lane = threadIdx.x & 31;
            int incl = l;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, off); if (lane >= off) incl += v; }
            unsigned long long base = 0;
            if (lane == 31 && incl > 0) base = atomicAdd(cig_used, (unsigned long long)incl);
            base = __shfl_sync(0xffffffffu, base, 31);
            coff = base + (unsigned long long)(incl - l);
            if (walk && !failed && coff + (unsigned long long)l > cig_cap) { r.status = 6; out[i] = r; failed = true; l = 0; }
            if (walk && !failed && l <= SHORT)
                for (int q = 0; q < l; ++q) cig[coff + (unsigned)(l - 1 - q)] = buf[q];
        }
    }
    if (!walk) return;
    if (failed) { if (r.status != 6) { r.status = 3; r.cigar_len = 0; out[i] = r; } return; }
    r.cigar_len = l;
    r.cigar_off = (int64_t)coff;
    out[i] = r;
}

}  // namespace mpn
