// Packed 16-bit score kernel for LONG reads and for scores that reach the int16 clamp (sm_100a): one warp per pair, the read
// cut into strips of 64 stages x KR rows that sweep the target one after the other.  Replaces sw_sse2_word (ssw.c:354-530)
// where the short-read kernel (sw_strip16.cuh) cannot go: reads longer than its largest strip, and pairs whose H can reach
// 32767, where `_mm_adds_epi16` (ssw.c:425) saturates and the reference's results depend on it (ONT-scale pairs, BASELINE
// configs[3] and [4]).
//
// Differences to sw_strip16.cuh (same systolic wavefront, same s16x2 packing of two stages per register):
//   * UNSIGNED arithmetic with a bias: every H / E / F is stored as value + LBIAS.  `VIADDMNMX.U16x2` then gives the clamp for free
//     -- h = min(Hdiag + s, 32767 + LBIAS) cannot wrap because Hdiag + s >= LBIAS - 128 > 0 -- where the signed form would overflow
//     (profiles/r01_ubench_sat.txt).  The floor max(., 0) becomes one extra VIMNMX.U16x2 against LBIAS.  E - gapE / F - gapE / H - gapO stay
//     >= LBIAS - 255 - 255 > 0, so nothing wraps below either.
//   * between strips the bottom row (H, F) and the running column maximum of every column go through a per-warp global buffer
//     (L2 resident: 8 B per column); the first stage of the next strip takes its boundary from there instead of zeros.
//   * reads are bottom-aligned in the LAST strip; the dead rows sit on top of the first strip.
// Per packed cell: PRMT, VIADDMNMX.U16x2 (diagonal + score, clamped), VIMNMX3.U16x2, VIMNMX.U16x2 (floor), 2x VIADDMNMX.U16x2 (E, F),
// 1/2 VIMNMX3.U16x2 (column maximum) on the alu pipe + one VIADD.16x2 on the fma pipe.
//
// NWP > 1 (small batches): NWP warps of a block share one pair.  Warp w takes strips w, w + NWP, ... and runs a few dozen columns behind
// the warp that owns the strip above: the producer publishes "32-column blocks flushed" in shared memory after a __threadfence, the
// consumer spins on it (bounded; a broken invariant traps instead of hanging) before it reads the boundary words with L2 loads.  The early
// end of reverse passes is not used in this mode (it changes no result: no column after the terminating one can hold a higher score).
//
// Output contract identical to the other score kernels: SwEnds per task (+ the per-column records on forward passes).
#pragma once
#include "sw_common.cuh"

namespace mpn {

constexpr int LONG_BLOCK = 128;                  // 4 warps = 4 pairs per block
// max(a, b) per half plus "a was already >= b" per half (unsigned twin of max2_track)
__device__ __forceinline__ uint32_t umax2_track(uint32_t a, uint32_t b, bool& a_ge_hi, bool& a_ge_lo)
{
    uint32_t val, ph, pl;
    asm("{.reg .pred pu, pv;\n\t"
        ".reg .u16 t0, t1, t2, t3;\n\t"
        "max.u16x2 %0, %3, %4;\n\t"
        "mov.b32 {t0, t1}, %0;\n\t"
        "mov.b32 {t2, t3}, %3;\n\t"
        "setp.eq.u16 pv, t0, t2;\n\t"
        "setp.eq.u16 pu, t1, t3;\n\t"
        "selp.b32 %1, 1, 0, pu;\n\t"
        "selp.b32 %2, 1, 0, pv;}"
        : "=&r"(val), "=&r"(ph), "=&r"(pl) : "r"(a), "r"(b));
    a_ge_hi = ph != 0; a_ge_lo = pl != 0;
    return val;
}

// shared memory: H-column snapshots [2 halves][ceil(KR/4)][LONG_BLOCK] uint4, then the per-warp boundary staging [2][32 steps][warps]
template <int KR>
__host__ __device__ constexpr size_t long16_smem_bytes() { return (size_t)2 * ((KR + 3) / 4) * LONG_BLOCK * sizeof(uint4) + (size_t)2 * 32 * (LONG_BLOCK / 32) * sizeof(uint32_t); }

template <int KR, int NWP>
__global__ void __launch_bounds__(LONG_BLOCK, 3)
sw_long16_kernel(const SwTask* __restrict__ tasks, int ntasks, int* __restrict__ counter, const int8_t* __restrict__ seq,
                 const Score16 sc, uint32_t* __restrict__ colrec, SwEnds* __restrict__ out, uint32_t* __restrict__ boundary, long long boundary_stride)
{
    constexpr int KRQ = (KR + 3) / 4;
    constexpr int G = 32;
    constexpr int CAP = 2 * G * KR;                 // rows per strip
    extern __shared__ uint4 lsnap[];               // [2 halves][KRQ][LONG_BLOCK]
    uint32_t* const stage = reinterpret_cast<uint32_t*>(lsnap + 2 * KRQ * LONG_BLOCK);   // [2 words][32 steps][warps per block]
    constexpr int NW = LONG_BLOCK / 32;

    static_assert(NWP == 1 || NWP == 2 || NWP == 4, "warps per pair");
    constexpr int PPB = NW / NWP;                   // pairs per block
    __shared__ int sh_ti[NW];                       // NWP > 1: task index of each pair of the block
    __shared__ int sh_progress[NW];                 // NWP > 1: per warp, strip << 16 | 32-column blocks whose boundary words are flushed
    __shared__ unsigned long long sh_key[NW];
    __shared__ int sh_row[NW], sh_wide[NW];
    __shared__ uint32_t smatrow[8];                 // matrix rows by target code (run-time index: shared memory, not the parameter struct)
    if (threadIdx.x < 8) smatrow[threadIdx.x] = sc.matrow[threadIdx.x];
    __syncthreads();
    const int tid = threadIdx.x, t = tid & 31, wib = tid >> 5;
    const int wip = wib % NWP, pib = wib / NWP;     // warp inside the pair, pair inside the block
    const long long wslot = NWP == 1 ? (long long)blockIdx.x * NW + wib : (long long)blockIdx.x * PPB + pib;
    uint32_t* const bA = boundary + wslot * 2 * boundary_stride;        // per column: F(bottom) << 16 | running column maximum
    uint32_t* const bB = bA + boundary_stride;                          // per column: H(bottom) << 16 (low half unused)
    const uint32_t mgapO2 = sc.mgapO2, mgapE2 = sc.mgapE2;
    // merge selectors (received word, own word): lo <- received.high, hi <- own.lo; thread 0 receives the boundary words instead
    const uint32_t selFH = 0x5432u;

    int prev_out = -1;                              // NWP > 1: output slot of the pair this warp group just finished (merged after the barrier)
    for (;;) {
        int ti = 0;
        if (NWP == 1) {
            if (t == 0) ti = atomicAdd(counter, 1);
            ti = __shfl_sync(0xffffffffu, ti, 0);
            if (ti >= ntasks) break;
        } else {
            __syncthreads();                        // every warp of the block has finished its strips of the previous pairs
            if (prev_out >= 0 && wip == 0 && t == 0) {
                // merge the warps of the pair: larger (score, -column) wins; on ties the smaller row (rows grow with the strip index)
                unsigned long long bk = 0; int br = 0, bw_ = 0;
                for (int q = 0; q < NWP; ++q) {
                    const unsigned long long k2 = sh_key[pib * NWP + q]; const int r2 = sh_row[pib * NWP + q];
                    bw_ |= sh_wide[pib * NWP + q];
                    if ((k2 >> 8) > (bk >> 8) || ((k2 >> 8) == (bk >> 8) && (k2 >> 40) != 0 && r2 < br)) { bk = k2; br = r2; }
                }
                SwEnds e;
                e.score = (int)(bk >> 40);
                e.col = e.score > 0 ? (int)(0xffffffu - (unsigned)((bk >> 8) & 0xffffffu)) : -1;
                e.row = e.score > 0 ? br : 0;
                e.flags = bw_ ? SW_FLAG_NEEDS_WIDE : 0;
                out[prev_out] = e;
            }
            prev_out = -1;
            __syncthreads();
            if (tid < PPB) sh_ti[tid] = atomicAdd(counter, 1);
            if (tid < NW) sh_progress[tid] = -1;
            __syncthreads();
            if (sh_ti[0] >= ntasks) break;          // task indices are handed out in order: the first one decides for the block
            ti = sh_ti[pib];
        }
        const bool have = ti < ntasks;
        SwTask tk;
        tk.rd_base = tk.rf_base = 0; tk.cm_off = -1; tk.rd_len = tk.rf_len = 0; tk.dir = 1; tk.out = 0; tk.stop = 0; tk.pad_ = 0;
        if (have) tk = tasks[ti];
        const int rd_len = tk.rd_len, tdir = tk.dir;
        int rf_len = tk.rf_len;
        if (rd_len <= 0 || rf_len <= 0) {
            if (have && t == 0 && wip == 0) { SwEnds e; e.score = 0; e.col = -1; e.row = 0; e.flags = 0; out[tk.out] = e; }
            if (NWP > 1 && t == 0) { sh_key[wib] = 0; sh_row[wib] = 0; sh_wide[wib] = 0; }
            continue;
        }
        const int nstrips = (rd_len + CAP - 1) / CAP;
        const int dead = nstrips * CAP - rd_len;
        const uint32_t stop2 = (NWP == 1 && tk.stop > 0) ? (((uint32_t)tk.stop + LBIAS) | (((uint32_t)tk.stop + LBIAS) << 16)) : 0u;
        const int nblocks_total = (rf_len + 2 * G - 1 + G - 1) / G;        // 32-step blocks of one strip (NWP > 1: no early end)
        unsigned long long runkey = 0;              // (score << 40 | 0xffffff - col << 8): best over the strips done so far
        int runrow = 0;
        int wide = 0;

        for (int strip = wip; strip < nstrips; strip += NWP) {
            const bool first = strip == 0, last = strip == nstrips - 1;
            // NWP > 1: wait until the warp that owns the strip above has flushed `need` blocks of boundary words
            auto wait_for = [&](int need) {
                if (NWP == 1 || first) return;
                const int want = ((strip - 1) << 16) | min(need, nblocks_total);
                const volatile int* pr = sh_progress + pib * NWP + (strip - 1) % NWP;
                unsigned spins = 0;
                while (*pr < want) {
                    __nanosleep(40);
                    if (++spins > (1u << 26)) __trap();                     // broken invariant: fail loudly instead of hanging the GPU
                }
                __threadfence();
            };
            uint32_t H[KR], E[KR], sel[KR];
            // selectors: low half = row of stage 2t, high half = row of stage 2t + 1 (sw_strip16.cuh)
#pragma unroll
            for (int j = 0; j < KR; ++j) {
                const int r_lo = strip * CAP + 2 * t * KR + j - dead, r_hi = r_lo + KR;
                uint32_t n_lo = 0x88u, n_hi = 0xccu;
                if (r_lo >= 0) {
                    const int q = seq[tk.rd_base + (int64_t)tdir * r_lo];
                    if ((unsigned)q < 4u) n_lo = (uint32_t)q | ((uint32_t)(q | 8) << 4); else wide = 1;
                }
                if (r_hi >= 0) {
                    const int q = seq[tk.rd_base + (int64_t)tdir * r_hi];
                    if ((unsigned)q < 4u) n_hi = (uint32_t)(q | 4) | ((uint32_t)(q | 12) << 4); else wide = 1;
                }
                sel[j] = n_lo | (n_hi << 8);
                H[j] = LBIAS2; E[j] = LBIAS2;
            }
            uint32_t Ftop = LBIAS2, Hdtop = LBIAS2, cmin = LBIAS2;
            uint32_t a = 0, b = 0;
            uint32_t best = LBIAS2, cvlo = 0, cvhi = 0;
            uint32_t hb_prev = LBIAS2;              // thread 0: bottom H of the strip above in the previous column (the diagonal), high half
            int nsteps = rf_len + 2 * G - 1;
            // chunk 0: matrix rows of the target bases and the boundary words of columns t
            uint32_t tnext = 0, anext = LBIAS2, bnext = LBIAS2;
            wait_for(3);                            // columns 0 .. 31 are flushed at the end of the producer's block 2
            if (t < rf_len) {
                tnext = smatrow[seq[tk.rf_base + (int64_t)tdir * t] & 7];
                if (!first) { anext = NWP == 1 ? bA[t] : __ldcg(bA + t); bnext = NWP == 1 ? bB[t] : __ldcg(bB + t); }
            }
            for (int s0 = 0; s0 < nsteps; s0 += G) {
                {   // reverse passes: the pass ends in the column whose maximum equals `stop` (ssw.c:483); a stage that has seen it
                    // bounds that column from above, and everything later is irrelevant for this and for the following strips
                    const uint32_t x = best ^ stop2;
                    const bool hit = stop2 != 0u && ((x & 0xffffu) == 0u || (x >> 16) == 0u);
                    if (__any_sync(0xffffffffu, hit)) nsteps = min(nsteps, s0 + 2 * G);
                }
                const uint32_t tchunk = tnext, achunk = anext, bchunk = bnext;
                {
                    const int idx = s0 + G + t;
                    tnext = 0; anext = LBIAS2; bnext = LBIAS2;
                    if (s0 + G < rf_len) wait_for(s0 / G + 4);      // columns of the next chunk: flushed at the end of the producer's block s0/G + 3
                    if (idx < rf_len) {
                        tnext = smatrow[seq[tk.rf_base + (int64_t)tdir * idx] & 7];
                        if (!first) { anext = NWP == 1 ? bA[idx] : __ldcg(bA + idx); bnext = NWP == 1 ? bB[idx] : __ldcg(bB + idx); }
                    }
                }
#pragma unroll 2
                for (int u = 0; u < G; ++u) {
                    const int s = s0 + u;
                    const uint32_t a0 = __shfl_sync(0xffffffffu, tchunk, u);
                    const uint32_t wa = __shfl_sync(0xffffffffu, achunk, u);
                    const uint32_t wb = __shfl_sync(0xffffffffu, bchunk, u);
                    if (t == 0) { a = a0; Ftop = prmt(wa, Ftop, 0x7632u); cmin = prmt(wa, cmin, 0x7610u); Hdtop = prmt(hb_prev, Hdtop, 0x7632u); hb_prev = wb; }
                    uint32_t F = Ftop, m = LBIAS2;
                    uint32_t h = uaddmin(Hdtop, prmt(a, b, sel[0]), LCAP2);
#pragma unroll
                    for (int j = 0; j < KR; ++j) {
                        uint32_t hnext = 0;
                        if (j + 1 < KR) hnext = uaddmin(H[j], prmt(a, b, sel[j + 1]), LCAP2);
                        else Hdtop = H[j];
                        const uint32_t Hn = umax2(umax3(h, E[j], F), LBIAS2);
                        const uint32_t Hg = add2(Hn, mgapO2);
                        E[j] = uaddmax(E[j], mgapE2, Hg);
                        F = uaddmax(F, mgapE2, Hg);
                        H[j] = Hn;
                        if (j & 1) m = umax3(m, H[j - 1], Hn);
                        else if (j == KR - 1) m = umax2(m, Hn);
                        h = hnext;
                    }
                    bool ge_hi, ge_lo;
                    best = umax2_track(best, m, ge_hi, ge_lo);
                    if (!ge_lo) {
                        cvlo = (uint32_t)s;
#pragma unroll
                        for (int k = 0; k < KRQ; ++k) lsnap[(size_t)k * LONG_BLOCK + tid] = make_uint4(H[4 * k], H[min(4 * k + 1, KR - 1)], H[min(4 * k + 2, KR - 1)], H[min(4 * k + 3, KR - 1)]);
                    }
                    if (!ge_hi) {
                        cvhi = (uint32_t)s;
#pragma unroll
                        for (int k = 0; k < KRQ; ++k) lsnap[(size_t)(KRQ + k) * LONG_BLOCK + tid] = make_uint4(H[4 * k], H[min(4 * k + 1, KR - 1)], H[min(4 * k + 2, KR - 1)], H[min(4 * k + 3, KR - 1)]);
                    }
                    const uint32_t cmout = umax2(cmin, m);
                    // the last stage (thread 31, high half) finishes column s - 63: stage its boundary words for the flush below
                    if (t == G - 1) {
                        stage[u * NW + wib] = prmt(cmout, F, 0x7632u);                 // F.hi << 16 | cm.hi
                        stage[(32 + u) * NW + wib] = H[KR - 1];                       // H(bottom).hi << 16 | (unused)
                    }
                    const uint32_t rF = __shfl_up_sync(0xffffffffu, F, 1);
                    const uint32_t rH = __shfl_up_sync(0xffffffffu, Hdtop, 1);
                    const uint32_t rC = __shfl_up_sync(0xffffffffu, cmout, 1);
                    const uint32_t rA = __shfl_up_sync(0xffffffffu, b, 1);
                    // thread 0: shfl_up returns its own word, whose high half is replaced by the boundary at the top of the next step
                    Ftop = prmt(rF, F, selFH);
                    Hdtop = prmt(rH, Hdtop, selFH);
                    cmin = prmt(rC, cmout, selFH);
                    b = a;
                    a = rA;
                }
                // ---- flush the 32 columns finished in this block: s0 - 63 + t
                __syncwarp();
                {
                    const int c = s0 + t - (2 * G - 1);
                    if (c >= 0 && c < rf_len) {
                        const uint32_t wA = stage[t * NW + wib], wB = stage[(32 + t) * NW + wib];
                        if (!last) { bA[c] = wA; bB[c] = wB; }
                        else if (tk.cm_off >= 0) colrec[tk.cm_off + c] = ((wA & 0xffffu) - LBIAS) | (((wB >> 16) - LBIAS) << 16);
                    }
                }
                if (NWP > 1 && !last) {             // publish: everything up to this block is visible to the consumer warp
                    __threadfence();
                    __syncwarp();
                    if (t == 0) *(volatile int*)(sh_progress + wib) = (strip << 16) | (s0 / G + 1);
                }
                __syncwarp();
            }
            // ---- strip result: reduce (score, first column, stage) over the 64 stages
            int sc_lo = (int)(best & 0xffffu) - (int)LBIAS, sc_hi = (int)(best >> 16) - (int)LBIAS;
            int col_lo = (int)cvlo - 2 * t, col_hi = (int)cvhi - 2 * t - 1;
            if (sc_lo <= 0) col_lo = 0;
            if (sc_hi <= 0) col_hi = 0;
            unsigned long long k_lo = ((unsigned long long)(unsigned)sc_lo << 40) | ((unsigned long long)(0xffffffu - (unsigned)col_lo) << 8) | (unsigned)(255 - 2 * t);
            unsigned long long k_hi = ((unsigned long long)(unsigned)sc_hi << 40) | ((unsigned long long)(0xffffffu - (unsigned)col_hi) << 8) | (unsigned)(254 - 2 * t);
            unsigned long long key = k_lo > k_hi ? k_lo : k_hi;
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) {
                unsigned long long o = __shfl_xor_sync(0xffffffffu, key, off);
                key = o > key ? o : key;
            }
            const int wscore = (int)(key >> 40);
            const int wstage = 255 - (int)(key & 0xffu);
            int row = 0;
            if (t == (wstage >> 1) && wscore > 0) {
                row = -999;
                const int half = wstage & 1;
                const uint4* sp = lsnap + (size_t)half * KRQ * LONG_BLOCK + tid;
                const int want = wscore + (int)LBIAS;
                for (int k = KRQ - 1; k >= 0; --k) {
                    uint4 v = sp[(size_t)k * LONG_BLOCK];
                    uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int q = 3; q >= 0; --q) {
                        const int hv = half ? (int)(w[q] >> 16) : (int)(w[q] & 0xffffu);
                        if (4 * k + q < KR && hv == want) row = strip * CAP + wstage * KR + 4 * k + q - dead;
                    }
                }
            }
            row = __shfl_sync(0xffffffffu, row, wstage >> 1);
            if ((key >> 8) > (runkey >> 8)) { runkey = key; runrow = row; }     // ties keep the earlier strip (smaller rows)
            if (NWP > 1 && !last && nblocks_total > 0 && t == 0) *(volatile int*)(sh_progress + wib) = (strip << 16) | 0xffff;   // strip complete
            // a strip that ended early bounds the terminating column for the strips below it
            // (columns up to nsteps - 64 have their boundary words written; the terminating column lies before that)
            if (nsteps < rf_len + 2 * G - 1) rf_len = min(rf_len, nsteps - (2 * G - 1));
            __syncwarp();
        }
        const unsigned anywide = __ballot_sync(0xffffffffu, wide != 0);
        if (NWP == 1) {
            if (t == 0) {
                SwEnds e;
                e.score = (int)(runkey >> 40);
                e.col = e.score > 0 ? (int)(0xffffffu - (unsigned)((runkey >> 8) & 0xffffffu)) : -1;
                e.row = e.score > 0 ? runrow : 0;
                e.flags = anywide ? SW_FLAG_NEEDS_WIDE : 0;
                out[tk.out] = e;
            }
        } else {
            // merge the warps of the pair: larger (score, -column) wins; on ties the smaller row (rows grow with the strip index)
            if (t == 0) { sh_key[wib] = runkey; sh_row[wib] = runrow; sh_wide[wib] = anywide ? 1 : 0; }
            prev_out = tk.out;
        }
    }
}

}  // namespace mpn
