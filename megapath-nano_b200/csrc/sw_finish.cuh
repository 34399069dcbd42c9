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
//     (pinned by tests/test_pad_rows.py).
//   * second best = first strictly greater maxColumn over [0, max(end_ref-maskLen,0)) and then over
//     [min(end_ref+maskLen, refLen) + (byte ? 1 : 0), refLen)  (ssw.c:310-323 vs :512-525); nothing if maskLen < 15
//     (ssw.c:809-815).
//   * end_ref starts at -1 in byte mode and 0 in word mode (ssw.c:145 / :371), visible when score == 0.
#pragma once
#include "sw_common.cuh"
#include <climits>

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

// value of `cur` (this chunk) or `prev` (previous chunk of 32 columns) at column (lane - k), 1 <= k <= 31
__device__ __forceinline__ int shift_back(int cur, int prev, int k, int lane)
{
    const int src = (lane - k) & 31;
    const int vc = __shfl_sync(0xffffffffu, cur, src);
    const int vp = __shfl_sync(0xffffffffu, prev, src);
    return lane >= k ? vc : vp;
}

// one warp per forward task
__global__ void __launch_bounds__(128)
sw_finish_kernel(const SwTask* __restrict__ fwd_tasks, int ntasks, const SwEnds* __restrict__ ends, const uint32_t* __restrict__ colrec,
                 PairArrays pa, FinishParams fp, FwdResult* __restrict__ res, SwTask* __restrict__ rev_tasks)
{
    __shared__ int fsm[4 * 128];                  // per warp: ring of B (64) + ring of prefix maxima (64)
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
        // Two 64-entry rings in shared memory (this warp's slice) hold B and the prefix maximum of the last two chunks, so "the value
        // d columns back" is one LDS instead of a two-shuffle select.  Entries of columns < 0 are the zero / -inf the formulas expect.
        int* ringB = fsm + (threadIdx.x >> 5) * 128;
        int* ringX = ringB + 64;
        const int NEGV = INT_MIN / 2;
        ringB[lane] = 0; ringB[32 + lane] = 0; ringX[lane] = NEGV; ringX[32 + lane] = NEGV;
        __syncwarp();
        int carryPM = NEGV;            // prefix max of (B[j] + j*gapE) over all earlier chunks
        int bestL = 0, idxL = 0, bestR = 0, idxR = 0;     // strict-greater-first maxima of the two ranges (per lane, merged at the end)
        bool anyL = false, anyR = false;
        for (int c0 = 0; c0 < rf_len; c0 += 32) {
            const int c = c0 + lane;
            const uint32_t w = c < rf_len ? rec[c] : 0u;
            const int cm = (int)(w & 0xffffu), B = (int)(w >> 16);
            int v = cm;
            if (P > 0) {
                ringB[c & 63] = B;
                // eroded contribution: PM[c-P-1] - gapO - (c-P-1)*gapE with PM the inclusive prefix max of B[j] + j*gapE
                int x = c < rf_len ? B + c * fp.gapE : NEGV;
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) {
                    const int o = __shfl_up_sync(0xffffffffu, x, off);
                    if (lane >= off) x = max(x, o);
                }
                x = max(x, carryPM);
                ringX[c & 63] = x;
                carryPM = __shfl_sync(0xffffffffu, x, 31);
                __syncwarp();
                // window maximum of B over the previous P (<= 15) columns
                int win = 0;
                for (int d = 1; d <= P; ++d) win = max(win, ringB[(c - d) & 63]);
                v = max(v, win);
                const int pm = ringX[(c - P - 1) & 63];
                v = max(v, pm - fp.gapO - (c - P - 1) * fp.gapE);       // columns < 0 hold -inf
                __syncwarp();
            }
            if (c < rf_len) {
                if (c < e1) { if (!anyL || v > bestL) { bestL = v; idxL = c; anyL = true; } }
                else if (c >= e2) { if (!anyR || v > bestR) { bestR = v; idxR = c; anyR = true; } }
            }
        }
        // merge lanes: larger value wins, ties -> smaller column
        if (!anyL) { bestL = -1; idxL = 0x7fffffff; }
        if (!anyR) { bestR = -1; idxR = 0x7fffffff; }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
            int ob = __shfl_xor_sync(0xffffffffu, bestL, off), oi = __shfl_xor_sync(0xffffffffu, idxL, off);
            if (ob > bestL || (ob == bestL && oi < idxL)) { bestL = ob; idxL = oi; }
            ob = __shfl_xor_sync(0xffffffffu, bestR, off); oi = __shfl_xor_sync(0xffffffffu, idxR, off);
            if (ob > bestR || (ob == bestR && oi < idxR)) { bestR = ob; idxR = oi; }
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
