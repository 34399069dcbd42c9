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

// one warp per forward task
__global__ void __launch_bounds__(128)
sw_finish_kernel(const SwTask* __restrict__ fwd_tasks, int ntasks, const SwEnds* __restrict__ ends, const uint32_t* __restrict__ colrec,
                 PairArrays pa, FinishParams fp, FwdResult* __restrict__ res, SwTask* __restrict__ rev_tasks)
{
    __shared__ int fsm[4 * 512];                  // per warp: ring of B (256) + ring of prefix maxima (256)
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
        // 128 columns per iteration, 4 consecutive columns per lane.  Two 256-entry rings in shared memory (this warp's slice) hold B and
        // the prefix maximum of the last two iterations, so "the value d columns back" is one LDS.  Entries of columns < 0 are the
        // zero / -inf the formulas expect.
        int* ringB = fsm + (threadIdx.x >> 5) * 512;
        int* ringX = ringB + 256;
        const int NEGV = INT_MIN / 2;
#pragma unroll
        for (int q = 0; q < 8; ++q) { ringB[q * 32 + lane] = 0; ringX[q * 32 + lane] = NEGV; }
        __syncwarp();
        int carryPM = NEGV;            // prefix max of (B[j] + j*gapE) over all earlier iterations
        int bestL = 0, idxL = 0, bestR = 0, idxR = 0;     // strict-greater-first maxima of the two ranges (per lane, merged at the end)
        bool anyL = false, anyR = false;
        for (int c0 = 0; c0 < rf_len; c0 += 128) {
            const int cb = c0 + 4 * lane;
            int cm[4], B[4], v[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint32_t w = cb + q < rf_len ? rec[cb + q] : 0u;
                cm[q] = (int)(w & 0xffffu); B[q] = (int)(w >> 16); v[q] = cm[q];
            }
            if (P > 0) {
                // inclusive prefix max of x = B[j] + j*gapE: serial inside the lane, one warp scan of the lane aggregates
                int x[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    x[q] = cb + q < rf_len ? B[q] + (cb + q) * fp.gapE : NEGV;
                    if (q) x[q] = max(x[q], x[q - 1]);
                    ringB[(cb + q) & 255] = B[q];
                }
                int inc = x[3];
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) {
                    const int o = __shfl_up_sync(0xffffffffu, inc, off);
                    if (lane >= off) inc = max(inc, o);
                }
                int pre = __shfl_up_sync(0xffffffffu, inc, 1);
                if (lane == 0) pre = NEGV;
                pre = max(pre, carryPM);
#pragma unroll
                for (int q = 0; q < 4; ++q) ringX[(cb + q) & 255] = max(x[q], pre);
                carryPM = max(carryPM, __shfl_sync(0xffffffffu, inc, 31));
                __syncwarp();
                // window maximum of B over the previous P (<= 15) columns: s0..s3 = running max over the d = 1.. columns before this
                // lane's block, captured at depths P, P-1, P-2, P-3 (what columns 0..3 of the block still see of it) ...
                int s0 = 0, s1 = 0, s2 = 0, s3 = 0;
                for (int d = 1; d <= P; ++d) { s3 = s2; s2 = s1; s1 = s0; s0 = max(s0, ringB[(cb - d) & 255]); }
                // ... plus the columns of the block itself that lie inside the window
                v[0] = max(v[0], s0);
                v[1] = max(v[1], max(s1, B[0]));
                v[2] = max(v[2], max(s2, P >= 2 ? max(B[1], B[0]) : B[1]));
                v[3] = max(v[3], max(s3, P >= 3 ? max(B[2], max(B[1], B[0])) : (P >= 2 ? max(B[2], B[1]) : B[2])));
                // eroded contribution: PM[c-P-1] - gapO - (c-P-1)*gapE with PM the inclusive prefix max of B[j] + j*gapE (-inf for columns < 0)
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int cc = cb + q - P - 1;
                    v[q] = max(v[q], ringX[cc & 255] - fp.gapO - cc * fp.gapE);
                }
                __syncwarp();
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int c = cb + q;
                if (c < rf_len) {
                    if (c < e1) { if (!anyL || v[q] > bestL) { bestL = v[q]; idxL = c; anyL = true; } }
                    else if (c >= e2) { if (!anyR || v[q] > bestR) { bestR = v[q]; idxR = c; anyR = true; } }
                }
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
