// Forward-pass epilogue (sm_100a): turns the per-column records of the score kernel into the second-best score exactly as
// ssw.c computes it, decides the 8-bit / 16-bit "mode" the reference would have ended in, and emits the reverse-pass task.
//
// Reference behaviour restated here:
//   * ssw_align runs the 8-bit kernel first and re-runs in 16 bit iff score + bias >= 255 (ssw.c:787-805, :271, :302).
//     The mode only matters through the number of zero-scoring pad rows the striped layout appends to the read:
//     W = 16 rows per segment vector in byte mode (ssw.c:95,108), W = 8 in word mode (ssw.c:335,346);
//     pad rows P = (W - readLen % W) % W.
//   * maxColumn[i] (ssw.c:280 / :482) is the column maximum INCLUDING those pad rows.  A pad row holds, by pure diagonal
//     moves, the bottom-row H of an earlier column, or that value eroded by one horizontal gap.  With B[i] = H of the
//     read's last row in column i the pad contribution to column i is
//         max( max_{1<=d<=P} B[i-d] ,  max_{d>P} B[i-d] - gapO - (d-P-1)*gapE )
//     (pinned by tests/test_pad_rows.py).  Every term belongs to one source column j, so the scans below are reductions over
//     (value, column) candidates: cm[j] at column j, B[j] at the first column of j+1 .. j+P inside the range, and the eroded
//     B[j] at the first column >= j+P+1 inside the range.
//   * second best = first strictly greater maxColumn over [0, max(end_ref-maskLen,0)) and then over
//     [min(end_ref+maskLen, refLen) + (byte ? 1 : 0), refLen)  (ssw.c:310-323 vs :512-525); nothing if maskLen < 15
//     (ssw.c:809-815).
//   * end_ref starts at -1 in byte mode and 0 in word mode (ssw.c:145 / :371), visible when score == 0.
#pragma once
#include "sw_common.cuh"
#include <climits>

#ifndef MPN_FINISH_LOADS
#define MPN_FINISH_LOADS 8          // column records a lane loads before it uses the first (A/B: 16)
#endif

namespace mpn {

struct PairArrays {               // device pointers, one entry per pair of the batch (lengths / offsets travel in the SwTask)
    const int32_t* masklen;
};

struct FinishParams {
    int32_t bias;                 // |min(mat)| as ssw_init computes it (ssw.c:741-745), 0 if no byte profile
    int32_t have_byte;            // score_size in {0, 2}
    int32_t have_word;            // score_size in {1, 2}
    int32_t gapO, gapE;
    int32_t flag, filters;        // ssw_align flag / score filter (ssw.c:817)
};

struct FwdResult {                // per pair, forward half of s_align (ssw.h:47-57)
    int32_t score1, score2, ref_end1, read_end1, ref_end2;
    int32_t word_mode;            // 1: reference ends in the 16-bit kernel
    int32_t status;               // 0 ok; 1: reference returns NULL (8-bit overflow without word profile, ssw.c:793-796); 2: no profile
    int32_t want_rev;             // reverse pass requested by the flag logic of ssw.c:817
};

// one warp per forward task
__global__ void __launch_bounds__(128)
sw_finish_kernel(const SwTask* __restrict__ fwd_tasks, int ntasks, const SwEnds* __restrict__ ends, const uint32_t* __restrict__ colrec,
                 PairArrays pa, FinishParams fp, FwdResult* __restrict__ res, SwTask* __restrict__ rev_tasks)
{
    const int lane = threadIdx.x & 31;
    const int k = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
    if (k >= ntasks) return;
    const SwTask tk = fwd_tasks[k];
    const int i = tk.out;
    const SwEnds e = ends[i];
    const int rd_len = tk.rd_len, rf_len = tk.rf_len, masklen = pa.masklen[i];

    FwdResult r;
    r.status = 0;
    int byte_mode = 0;
    if (fp.have_byte && e.score + fp.bias < 255) byte_mode = 1;
    else if (fp.have_byte && !fp.have_word) r.status = 1;
    else if (!fp.have_byte && !fp.have_word) r.status = 2;
    r.word_mode = byte_mode ? 0 : 1;
    r.score1 = e.score;
    r.ref_end1 = e.score > 0 ? e.col : (byte_mode ? -1 : 0);
    r.read_end1 = e.score > 0 ? e.row : 0;
    r.score2 = 0;
    r.ref_end2 = masklen >= 15 ? 0 : -1;

    if (masklen >= 15 && rf_len > 0 && r.status == 0 && tk.cm_off >= 0) {
        const int W = byte_mode ? 16 : 8;
        const int P = (W - rd_len % W) % W;
        const int end_ref = r.ref_end1;
        const int e1 = max(end_ref - masklen, 0);
        const int e2 = min(end_ref + masklen, rf_len) + (byte_mode ? 1 : 0);
        const uint32_t* rec = colrec + tk.cm_off;
        // Both scans are "largest value, smallest column" reductions, and every term of the pad formula is tied to ONE source column j:
        //   cm[j] lands on column j;  B[j] carried diagonally lands on columns j+1 .. j+P (first one counts);  B[j] eroded by a horizontal
        //   gap lands on columns >= j+P+1 with B[j] - gapO - (c-P-1-j)*gapE, largest at the first such column.
        // So each lane walks its columns once and offers at most three (value, column) candidates per range -- no prefix scans, no rings.
        int bestL = 0, idxL = 0, bestR = 0, idxR = 0;
        if (rf_len <= 0xffff) {
            // usual case: value (<= 65535) and column both fit 16 bits -> one 32-bit key, one IMNMX per candidate; values <= 0 clamp to 0,
            // which never beats the "strictly greater than 0" start of the scans
            uint32_t kL = 0u, kR = 0u;                    // (value << 16) | (0xffff - column)
            auto offer = [](uint32_t& k, int value, int col) { k = max(k, ((uint32_t)max(value, 0) << 16) | (uint32_t)(0xffff - col)); };
            auto column = [&](const int j, const uint32_t w) {
                const int cmj = (int)(w & 0xffffu), Bj = (int)(w >> 16);
                if (j < e1) offer(kL, cmj, j);
                else if (j >= e2) offer(kR, cmj, j);
                // pad candidates: skipped unless one of them can still beat (or tie, at a smaller column) what the warp holds.  Left range: Bj lands
                // on column j + 1.  Right range: Bj itself if the range starts within the P columns behind j, else only the eroded value, which
                // has lost gapO + (distance) * gapE by the time it gets there -- so columns far left of the range never qualify.
                bool may = false;
                if (P > 0 && Bj > 0) {
                    may = j + 1 < e1 && (uint32_t)Bj >= (kL >> 16);
                    int vR = Bj;
                    if (e2 > j + P) vR = Bj - fp.gapO - (e2 - P - 1 - j) * fp.gapE;
                    may = may || (e2 < rf_len && vR > 0 && (uint32_t)vR >= (kR >> 16));
                }
                if (may) {
                    if (j + 1 < e1) offer(kL, Bj, j + 1);
                    if (j + P + 1 < e1) offer(kL, Bj - fp.gapO, j + P + 1);
                    const int c2 = max(j + 1, e2);
                    if (c2 <= j + P && c2 < rf_len) offer(kR, Bj, c2);
                    const int c3 = max(j + P + 1, e2);
                    if (c3 < rf_len) offer(kR, Bj - fp.gapO - (c3 - P - 1 - j) * fp.gapE, c3);
                }
            };
            // FIN_Q x 32 columns per iteration: that many coalesced loads in flight per lane before any of them is used (the walk is latency bound otherwise)
            constexpr int FIN_Q = MPN_FINISH_LOADS;
            for (int c0 = 0; c0 < rf_len; c0 += 32 * FIN_Q) {
                uint32_t w8[FIN_Q];
#pragma unroll
                for (int q = 0; q < FIN_Q; ++q) { const int j = c0 + 32 * q + lane; w8[q] = j < rf_len ? rec[j] : 0u; }
#pragma unroll
                for (int q = 0; q < FIN_Q; ++q) { const int j = c0 + 32 * q + lane; if (j < rf_len) column(j, w8[q]); }
                // share the keys across the warp (the result is their maximum anyway): from the second block on, the test that skips the pad
                // candidates is then nearly warp-uniform and whole warps jump over that code
                if (c0 + 32 * FIN_Q < rf_len) {
#pragma unroll
                    for (int off = 16; off >= 1; off >>= 1) {
                        kL = max(kL, __shfl_xor_sync(0xffffffffu, kL, off));
                        kR = max(kR, __shfl_xor_sync(0xffffffffu, kR, off));
                    }
                }
            }
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) {
                kL = max(kL, __shfl_xor_sync(0xffffffffu, kL, off));
                kR = max(kR, __shfl_xor_sync(0xffffffffu, kR, off));
            }
            bestL = (int)(kL >> 16); idxL = 0xffff - (int)(kL & 0xffffu);
            bestR = (int)(kR >> 16); idxR = 0xffff - (int)(kR & 0xffffu);
        } else {
            unsigned long long kL = 0ull, kR = 0ull;      // (value << 32) | ~column
            auto offer = [](unsigned long long& k, int value, int col) {
                if (value > 0) {
                    const unsigned long long c = ((unsigned long long)(unsigned)value << 32) | (unsigned long long)(0xffffffffu - (unsigned)col);
                    k = c > k ? c : k;
                }
            };
            for (int j = lane; j < rf_len; j += 32) {
                const uint32_t w = rec[j];
                const int cmj = (int)(w & 0xffffu), Bj = (int)(w >> 16);
                if (j < e1) offer(kL, cmj, j);
                else if (j >= e2) offer(kR, cmj, j);
                if (P > 0 && Bj > 0) {
                    if (j + 1 < e1) offer(kL, Bj, j + 1);
                    if (j + P + 1 < e1) offer(kL, Bj - fp.gapO, j + P + 1);
                    const int c2 = max(j + 1, e2);
                    if (c2 <= j + P && c2 < rf_len) offer(kR, Bj, c2);
                    const int c3 = max(j + P + 1, e2);
                    if (c3 < rf_len) offer(kR, Bj - fp.gapO - (c3 - P - 1 - j) * fp.gapE, c3);
                }
            }
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) {
                const unsigned long long oL = __shfl_xor_sync(0xffffffffu, kL, off), oR = __shfl_xor_sync(0xffffffffu, kR, off);
                kL = oL > kL ? oL : kL; kR = oR > kR ? oR : kR;
            }
            bestL = (int)(kL >> 32); idxL = (int)(0xffffffffu - (unsigned)(kL & 0xffffffffull));
            bestR = (int)(kR >> 32); idxR = (int)(0xffffffffu - (unsigned)(kR & 0xffffffffull));
        }
        int s2 = 0, r2 = 0;
        if (bestL > s2) { s2 = bestL; r2 = idxL; }
        if (bestR > s2) { s2 = bestR; r2 = idxR; }
        r.score2 = s2;
        r.ref_end2 = r2;
    }

    // reverse pass wanted?  (ssw.c:817)
    r.want_rev = !(fp.flag == 0 || (fp.flag == 2 && r.score1 < fp.filters)) && r.status == 0;
    if (lane == 0) {
        res[i] = r;
        SwTask rt;
        rt.out = i; rt.cm_off = -1; rt.dir = -1; rt.stop = r.score1; rt.pad_ = 0;
        if (r.want_rev && r.score1 > 0) {
            rt.rd_len = r.read_end1 + 1; rt.rf_len = r.ref_end1 + 1;
            rt.rd_base = tk.rd_base + r.read_end1; rt.rf_base = tk.rf_base + r.ref_end1;
        } else {
            rt.rd_len = 0; rt.rf_len = 0; rt.rd_base = tk.rd_base; rt.rf_base = tk.rf_base;
        }
        rev_tasks[k] = rt;
    }
}

}  // namespace mpn
