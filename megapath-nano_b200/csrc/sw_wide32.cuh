// 32-bit intra-task score kernel (sm_100a): one warp per pair, for everything the packed 16-bit kernel cannot take --
// ONT-scale pairs whose scores reach the int16 clamp of sw_sse2_word (ssw.c:425 `_mm_adds_epi16`, result fields are
// uint16_t), reads longer than the largest packed strip, reads with codes >= 4 that the packed kernel's N variants do not take (varying N column, codes above 4, long reads), and alphabets with n > 8.
//
// The read is cut into strips of 32 lanes x WIDE_KR rows (16, or 8 when all reads of the batch are short); inside a strip lane L runs WIDE_KR rows of column s-L at step s
// (anti-diagonal wavefront over the lanes).  Between strips the bottom row (H, F) and the running column maximum live in a
// per-warp global boundary buffer, streamed through registers 32 columns at a time.  Scores are int32 with the reference's
// clamp applied per cell: h = min(Hdiag + s, 32767) is one VIADDMNMX.  Substitution scores come from a copy of the matrix
// in shared memory, extended by one all-zero row and column used for dead rows (strip padding on top of the read) and for
// columns outside the target.
//
// Output contract identical to the packed kernel: SwEnds per task (+ the per-column records on forward passes).
#pragma once
#include "sw_common.cuh"

namespace mpn {

constexpr int WIDE_KR_LONG = 16;                 // rows per lane: 512-row strips for long reads ...
constexpr int WIDE_KR_SHORT = 8;                 // ... 256-row strips when every read of the batch is short (fewer dead lanes)
constexpr int WIDE_BLOCK = 128;

template <int WIDE_KR>
__global__ void __launch_bounds__(WIDE_BLOCK, 3)
sw_wide32_kernel(const SwTask* __restrict__ tasks, int ntasks, int* __restrict__ counter, const int8_t* __restrict__ seq,
                 const int8_t* __restrict__ mat, int n, int gapO, int gapE, uint32_t* __restrict__ colrec, SwEnds* __restrict__ out,
                 int* __restrict__ boundary, long long boundary_stride, int only_flagged)
{
    constexpr int WIDE_CAP = 32 * WIDE_KR;
    extern __shared__ int wsm[];
    const int n1 = n + 1;
    int* smat = wsm;                                              // (n+1) x (n+1), last row / column zero
    int* snap = wsm + ((n1 * n1 + 31) & ~31);                     // [WIDE_KR][WIDE_BLOCK]
    for (int i = threadIdx.x; i < n1 * n1; i += blockDim.x) {
        const int t = i / n1, q = i % n1;
        smat[i] = (t < n && q < n) ? (int)mat[t * n + q] : 0;
    }
    __syncthreads();

    const int tid = threadIdx.x, lane = tid & 31;
    const long long wslot = (long long)blockIdx.x * (WIDE_BLOCK / 32) + (tid >> 5);
    uint32_t* brec = reinterpret_cast<uint32_t*>(boundary + wslot * 2 * boundary_stride);   // (cm | Hbot << 16) per column
    int* bF = boundary + wslot * 2 * boundary_stride + boundary_stride;                      // bottom-row F per column
    const int mgapO = -gapO, mgapE = -gapE;

    // Task fetch.  Direct mode: one task per atomic.  Flagged-only mode (the re-run pass over ALL tasks of a batch, of which usually none is
    // flagged): a warp takes 32 tasks per atomic, every lane checks one flag, and only the flagged ones are processed -- 1/32 of the atomics.
    int chunk_base = 0;
    unsigned chunk_todo = 0u;
    for (;;) {
        int ti = 0;
        if (only_flagged) {
            while (chunk_todo == 0u) {
                if (lane == 0) chunk_base = atomicAdd(counter, 32);
                chunk_base = __shfl_sync(0xffffffffu, chunk_base, 0);
                if (chunk_base >= ntasks) break;
                const int mine = chunk_base + lane;
                chunk_todo = __ballot_sync(0xffffffffu, mine < ntasks && (out[tasks[mine].out].flags & SW_FLAG_NEEDS_WIDE) != 0);
            }
            if (chunk_todo == 0u) break;
            ti = chunk_base + __ffs((int)chunk_todo) - 1;
            chunk_todo &= chunk_todo - 1u;
        } else {
            if (lane == 0) ti = atomicAdd(counter, 1);
            ti = __shfl_sync(0xffffffffu, ti, 0);
            if (ti >= ntasks) break;
        }
        const SwTask tk = tasks[ti];
        const int rd_len = tk.rd_len, rf_len = tk.rf_len, dir = tk.dir;
        if (rd_len <= 0 || rf_len <= 0) {
            if (lane == 0) { SwEnds e; e.score = 0; e.col = -1; e.row = 0; e.flags = 0; out[tk.out] = e; }
            continue;
        }
        const int nstrips = (rd_len + WIDE_CAP - 1) / WIDE_CAP;
        const int dead = nstrips * WIDE_CAP - rd_len;
        unsigned long long runkey = 0;      // (score << 24 | 0xffffff - col): best over the strips done so far
        int runrow = 0;

        for (int strip = 0; strip < nstrips; ++strip) {
            const int row0 = strip * WIDE_CAP - dead + lane * WIDE_KR;   // read row of this lane's first row (negative: dead)
            int qo[WIDE_KR], H[WIDE_KR], E[WIDE_KR];
#pragma unroll
            for (int j = 0; j < WIDE_KR; ++j) {
                const int r = row0 + j;
                int q = n;
                if (r >= 0) { q = seq[tk.rd_base + (long long)dir * r]; if ((unsigned)q >= (unsigned)n) q = n; }
                qo[j] = q; H[j] = 0; E[j] = 0;
            }
            const bool last_strip = strip == nstrips - 1;
            int Ftop = 0, Hdtop = 0, cmin = 0, trow = n * n1;
            int best = 0, cv = 0, hb_last = 0;
            int tchunk = 0, rchunk = 0, fchunk = 0;
            const int nsteps = rf_len + 31;
            for (int s0 = 0; s0 < nsteps; s0 += 32) {
                {   // stage the next 32 columns: target codes and (for strips below the first) the boundary left by the strip above
                    const int c = s0 + lane;
                    tchunk = n; rchunk = 0; fchunk = 0;
                    if (c < rf_len) {
                        int t = seq[tk.rf_base + (long long)dir * c];
                        tchunk = (unsigned)t < (unsigned)n ? t : n;
                        if (strip > 0) { rchunk = (int)brec[c]; fchunk = bF[c]; }
                    }
                }
                const int umax = min(32, nsteps - s0);
                for (int u = 0; u < umax; ++u) {
                    const int s = s0 + u;
                    {
                        const int t0 = __shfl_sync(0xffffffffu, tchunk, u);
                        const int r0 = __shfl_sync(0xffffffffu, rchunk, u);
                        const int f0 = __shfl_sync(0xffffffffu, fchunk, u);
                        if (lane == 0) {
                            trow = t0 * n1;
                            Ftop = f0; cmin = r0 & 0xffff; Hdtop = hb_last; hb_last = (int)((unsigned)r0 >> 16);
                        }
                    }
                    int F = Ftop, m = 0, hd = Hdtop;
#pragma unroll
                    for (int j = 0; j < WIDE_KR; ++j) {
                        const int sc = smat[trow + qo[j]];
                        const int h = __viaddmin_s32(hd, sc, 32767);          // ssw.c:425 saturating add
                        hd = H[j];
                        const int Hn = __vimax3_s32_relu(h, E[j], F);
                        const int Hg = Hn + mgapO;
                        E[j] = __viaddmax_s32_relu(E[j], mgapE, Hg);
                        F = __viaddmax_s32_relu(F, mgapE, Hg);
                        H[j] = Hn;
                        if (j & 1) m = __vimax3_s32(m, H[j - 1], Hn);
                    }
                    if (m > best) {                                            // strict: first column wins (ssw.c:474)
                        best = m; cv = s;
#pragma unroll
                        for (int j = 0; j < WIDE_KR; ++j) snap[j * WIDE_BLOCK + tid] = H[j];
                    }
                    const int cmout = max(cmin, m);
                    if (lane == 31) {
                        const int c = s - 31;
                        if (c >= 0 && c < rf_len) {
                            const uint32_t rec = (uint32_t)cmout | ((uint32_t)H[WIDE_KR - 1] << 16);
                            if (!last_strip) { brec[c] = rec; bF[c] = F; }
                            else if (tk.cm_off >= 0) colrec[tk.cm_off + c] = rec;
                        }
                    }
                    Ftop = __shfl_up_sync(0xffffffffu, F, 1);
                    Hdtop = __shfl_up_sync(0xffffffffu, hd, 1);
                    cmin = __shfl_up_sync(0xffffffffu, cmout, 1);
                    trow = __shfl_up_sync(0xffffffffu, trow, 1);
                }
            }
            // ---- strip result: best score, first column, smallest row
            int col = best > 0 ? cv - lane : 0;
            unsigned long long key = ((unsigned long long)(unsigned)best << 32) | ((unsigned long long)(0xffffffu - (unsigned)col) << 8) | (unsigned)(255 - lane);
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) {
                unsigned long long o = __shfl_xor_sync(0xffffffffu, key, off);
                key = o > key ? o : key;
            }
            const int wlane = 255 - (int)(key & 0xffu);
            int row = 0;
            if (lane == wlane && best > 0) {
                row = -999;
                for (int j = WIDE_KR - 1; j >= 0; --j) if (snap[j * WIDE_BLOCK + tid] == best) row = row0 + j;
            }
            row = __shfl_sync(0xffffffffu, row, wlane);
            if ((key >> 8) > (runkey >> 8)) { runkey = key; runrow = row; }     // ties keep the earlier strip (smaller rows)
            __syncwarp();
        }
        if (lane == 0) {
            SwEnds e;
            e.score = (int)(runkey >> 32);
            e.col = e.score > 0 ? (int)(0xffffffu - (unsigned)((runkey >> 8) & 0xffffffu)) : -1;
            e.row = e.score > 0 ? runrow : 0;
            e.flags = 0;
            out[tk.out] = e;
        }
    }
}

inline size_t wide32_smem_bytes(int n, int kr) { return sizeof(int) * (size_t)((((n + 1) * (n + 1) + 31) & ~31) + kr * WIDE_BLOCK); }

// host launchers.  `boundary` must hold 2 * stride ints per warp of the grid.
inline void launch_wide32_impl(const SwTask* tasks, int ntasks, int* counter, const int8_t* seq, const int8_t* mat, int n, int gapO, int gapE,
                               uint32_t* colrec, SwEnds* ends, int* boundary, long long stride, int blocks, int only_flagged, bool short_reads, cudaStream_t st)
{
    if (ntasks <= 0) return;
    if (short_reads) sw_wide32_kernel<WIDE_KR_SHORT><<<blocks, WIDE_BLOCK, wide32_smem_bytes(n, WIDE_KR_SHORT), st>>>(tasks, ntasks, counter, seq, mat, n, gapO, gapE, colrec, ends, boundary, stride, only_flagged);
    else sw_wide32_kernel<WIDE_KR_LONG><<<blocks, WIDE_BLOCK, wide32_smem_bytes(n, WIDE_KR_LONG), st>>>(tasks, ntasks, counter, seq, mat, n, gapO, gapE, colrec, ends, boundary, stride, only_flagged);
}

}  // namespace mpn
