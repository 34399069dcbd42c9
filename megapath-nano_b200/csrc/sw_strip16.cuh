// Packed 16-bit inter-task score kernel (sm_100a): replaces the score passes of ssw.c -- sw_sse2_byte (ssw.c:123-328),
// sw_sse2_word (ssw.c:354-530) and their query profiles qP_byte / qP_word (ssw.c:89-114, 330-352).
//
// Work decomposition (one pair per thread group, G threads, G | 32):
//   * the read is laid over 2*G "stages" of KR rows each; thread t owns stage 2t in the LOW 16-bit halves of its
//     registers and stage 2t+1 in the HIGH halves (s16x2 packing of two pipeline stages of the SAME pair);
//   * stage v processes target column c at step s = c + v (systolic wavefront).  Per step a thread updates KR packed
//     cells with H/E/F in registers; the bottom H/F, the running column maximum and the target's matrix row move to the
//     next stage by warp shuffle (between threads) or by a byte permute (low half -> high half of the same thread);
//   * reads shorter than the strip are aligned to its BOTTOM: the dead rows on top score <= 0 and therefore stay 0, which
//     is exactly the H[-1][*] = 0 boundary, and the last stage's bottom row is always the read's last row;
//   * substitution scores come from one PRMT per packed cell: the selector (per row, built once per task from the read)
//     picks mat[t_lo][q_lo] and mat[t_hi][q_hi] out of the two 4-byte matrix rows of the current target bases and
//     sign-extends them (selector bit 3).  This replaces the striped query profile.
//
// Per packed cell: 4.5 alu-pipe instructions (PRMT, VIMNMX3.S16x2.RELU, 2x VIADDMNMX.S16x2.RELU, 1/2 VIMNMX3.S16x2 for
// the column maximum) + 2 VIADD.16x2 on the fma pipe.
//
// What the kernel returns per task: the maximum score, the first column attaining it and the smallest row of that
// column attaining it (ssw.c:260-277, 284-293 tie rules), and -- for forward passes -- one record per target column
// holding (column maximum over the real rows, H of the read's last row).  The pad rows that the SSE2 layout adds
// (ssw.c:108, 346) are applied afterwards, analytically, by sw_finish_kernel (sw_finish.cuh) from these records.
//
// LONG = true (G = 32, one warp per pair): reads of ANY length and scores that reach the int16 clamp of ssw.c:425 (ONT-scale pairs,
// BASELINE configs[3] and [4]).  The read is cut into strips of 64 stages x KR rows that sweep the target one after the other; between
// strips the bottom row (H, F) and the running column maximum of every column go through a per-warp global buffer (8 B per column,
// L2 resident), which the first thread of the next strip takes as its boundary.  Arithmetic is UNSIGNED with a bias (every H / E / F is
// value + LBIAS): `VIADDMNMX.U16x2` then carries h = min(Hdiag + s, 32767 + LBIAS) -- `_mm_adds_epi16` -- without wrapping (the signed
// form overflows its 16-bit sum first, profiles/r01_ubench_sat.txt); the floor max(., 0) is one more VIMNMX.U16x2 against LBIAS.
#pragma once
#include "sw_common.cuh"
#include <cstdio>

namespace mpn {

constexpr int STRIP_BLOCK = 128;

#ifndef MPN_STRIP_UNROLL
#define MPN_STRIP_UNROLL 2          // wavefront steps unrolled inside a block of STRIP_CK steps (8: all of them)
#endif
constexpr int STRIP_UNROLL = MPN_STRIP_UNROLL;
#ifndef MPN_STRIP_MINB
#define MPN_STRIP_MINB 4
#endif

// Cell of the maximum (ssw.c:260-277: first column holding it; ssw.c:284-293: smallest row of that column): checkpoint + replay, in
// forward and reverse passes alike.  Every STRIP_CK steps a thread stores its state (H, E of its rows, the boundary entering its
//   stages) into a scratch slot, and every step the 2 x 16 bits it receives from the thread above into the slot's log; when the block of
//   steps ends and the thread's better stage improved in it, scratch and committed slot swap.  At the end of the pass the threads that hold
//   the maximum restore their committed slot and re-run at most STRIP_CK steps on their own (the log replaces the shuffles) to find their
//   first cell holding it.  This replaces the per-improvement H-column snapshots of round 1 (8 predicated STS.128 + 15 register moves per step).
__host__ __device__ constexpr int strip_ck(int G) { return G < 8 ? G : 8; }
template <int KR, int G>
__host__ __device__ constexpr size_t strip16_slot_bytes() { return ((size_t)2 * ((KR + 3) / 4) * 16 + 8 + (size_t)strip_ck(G) * 4) * STRIP_BLOCK; }
// shared memory: 2 checkpoint slots, the staging of the last stage's per-step output [G steps][groups of the block] (column records on
// forward passes; + the F leaving the strip in LONG mode), and, in the N variant only, the per-row score fix-up selectors [KR][STRIP_BLOCK]
template <int KR, int G, bool NM = false, bool REV = false, bool LONG = false>
__host__ __device__ constexpr size_t strip16_smem_bytes()
{
    return 2 * strip16_slot_bytes<KR, G>() + (size_t)((REV && !LONG ? 0 : STRIP_BLOCK * (LONG ? 2 : 1)) + (NM ? KR * STRIP_BLOCK : 0)) * sizeof(uint32_t);
}

// signed arithmetic with the floor at 0 (short reads, scores below the clamp) / unsigned arithmetic with a bias and the int16 clamp (LONG)
template <bool LONG> struct StripArith {
    static constexpr uint32_t Z = 0u;
    static __device__ __forceinline__ uint32_t hadd(uint32_t hd, uint32_t s) { return add2(hd, s); }
    static __device__ __forceinline__ uint32_t hmax(uint32_t h, uint32_t e, uint32_t f) { return max3_relu(h, e, f); }
    static __device__ __forceinline__ uint32_t ef(uint32_t x, uint32_t mge, uint32_t hg) { return addmax_relu(x, mge, hg); }
    static __device__ __forceinline__ uint32_t mx2(uint32_t a, uint32_t b) { return max2(a, b); }
    static __device__ __forceinline__ uint32_t mx3(uint32_t a, uint32_t b, uint32_t c) { return max3(a, b, c); }
    static __device__ __forceinline__ int lo(uint32_t x) { return (int)(int16_t)(x & 0xffffu); }
    static __device__ __forceinline__ int hi(uint32_t x) { return (int)(int16_t)(x >> 16); }
};
template <> struct StripArith<true> {
    static constexpr uint32_t Z = LBIAS2;
    static __device__ __forceinline__ uint32_t hadd(uint32_t hd, uint32_t s) { return uaddmin(hd, s, LCAP2); }                 // the clamp of ssw.c:425
    static __device__ __forceinline__ uint32_t hmax(uint32_t h, uint32_t e, uint32_t f) { return umax2(umax3(h, e, f), LBIAS2); }
    static __device__ __forceinline__ uint32_t ef(uint32_t x, uint32_t mge, uint32_t hg) { return uaddmax(x, mge, hg); }
    static __device__ __forceinline__ uint32_t mx2(uint32_t a, uint32_t b) { return umax2(a, b); }
    static __device__ __forceinline__ uint32_t mx3(uint32_t a, uint32_t b, uint32_t c) { return umax3(a, b, c); }
    static __device__ __forceinline__ int lo(uint32_t x) { return (int)(x & 0xffffu) - (int)LBIAS; }
    static __device__ __forceinline__ int hi(uint32_t x) { return (int)(x >> 16) - (int)LBIAS; }
};

// REV = false: forward passes (column records written, no early end).  REV = true: reverse passes (ssw.c:820-832): no column records, no
// column-maximum chain between the stages, and the pass ends once every stage is past the first column that reached the terminating score.
//
// NM = false: the kernel every pair goes through.  A read that contains N (code 4) cannot be scored by the one-PRMT lookup (the 4-byte
// matrix rows have no slot for a fifth read code): the pair is flagged and its task index appended to `relist` (relist[0] = count).
// NM = true: the re-run of those pairs.  When the matrix's N column is one constant c (both matrix builders of the reference:
// ssw_cpp.cpp:23-48, pyssw.py:61-79) every row gets a second PRMT that swaps its half of the looked-up score for c where the read has
// an N (selector in shared memory, identity elsewhere).  This variant walks `relist` instead of a task range (`tasks` is then the whole
// task array and `aux` the largest read length a smaller N variant already took); pairs it cannot take either (codes above 4) stay
// flagged for the int32 kernel.  Keeping the N code out of the main instantiation keeps its inner loop free of the second PRMT
// (a switchable version cost 3 % on reads without N).
//
// boundary / boundary_stride (LONG only): per warp of the grid two arrays of boundary_stride words -- per column (running column maximum |
// bottom H << 16) and (F leaving the bottom row << 16) of the strip above.
template <int KR, int G, bool REV, bool NM = false, bool LONG = false>
__global__ void __launch_bounds__(STRIP_BLOCK, MPN_STRIP_MINB)
sw_strip16_kernel(const SwTask* __restrict__ tasks, int ntasks, int* __restrict__ counter, const int8_t* __restrict__ seq,
                  const Score16 sc, uint32_t* __restrict__ colrec, SwEnds* __restrict__ out, int* __restrict__ relist, int aux,
                  uint32_t* __restrict__ boundary, long long boundary_stride)
{
    static_assert(KR >= 2, "KR too small");
    static_assert(G == 2 || G == 4 || G == 8 || G == 16 || G == 32, "G must divide 32");
    static_assert(!LONG || (G == 32 && !NM), "the multi-strip mode runs one warp per pair");
    using AR = StripArith<LONG>;
    constexpr uint32_t Z = AR::Z;
    constexpr int KRQ = (KR + 3) / 4;             // 16-byte words per checkpointed array
    constexpr int CAP = 2 * G * KR;               // rows covered by one strip
    constexpr int CK = strip_ck(G);               // steps between checkpoints
    constexpr int NG = STRIP_BLOCK / G;           // groups per block
    constexpr bool STAGED = !REV || LONG;         // the last stage's per-step output is staged and written out every G steps
    constexpr uint32_t SLOT = (uint32_t)strip16_slot_bytes<KR, G>();
    constexpr uint32_t SLOT_FH = 2 * KRQ * 16 * STRIP_BLOCK, SLOT_LOG = SLOT_FH + 8 * STRIP_BLOCK;     // byte offsets inside a slot
    extern __shared__ uint4 slots[];              // 2 checkpoint slots
    unsigned char* const smem0 = reinterpret_cast<unsigned char*>(slots);
    uint32_t* const crow = reinterpret_cast<uint32_t*>(smem0 + (size_t)2 * SLOT);         // [G steps][NG]: (column maximum | bottom H << 16) of the last stage
    uint32_t* const frow = crow + STRIP_BLOCK;                                            // LONG: [G steps][NG]: F leaving the last stage
    uint32_t* const nfix = crow + (STAGED ? STRIP_BLOCK * (LONG ? 2 : 1) : 0);           // NM only: [KR][STRIP_BLOCK] fix-up selectors

    // the 8 matrix rows are looked up by a run-time target code: shared memory (one LDS) instead of the by-value parameter struct
    // (which ptxas can only index with a chain of predicated constant loads)
    // entries 8..15 are 0: the code of a column past the end of the target (TB_NONE) looks up an all-zero row
    __shared__ uint32_t smatrow[16];
    if (threadIdx.x < 16) smatrow[threadIdx.x] = threadIdx.x < 8 ? sc.matrow[threadIdx.x] : 0u;
    __syncthreads();
    constexpr uint32_t TB_NONE = 8;
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int t = lane % G;                       // thread index inside the group
    const int gid = tid / G;                      // group index inside the block
    const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (lane - t));
    // merge selectors for (value received from thread t-1, own value): low half <- received.high (or .low), high half <- own.low.
    // Single strip: for the first thread of a group the low half must be the matrix boundary (0): select the sign byte of own byte 7
    // (all merged quantities are >= 0, so that byte replicates to 0x00).  LONG: the first thread receives the strip boundary instead.
    const uint32_t mergeHi = (t == 0 && !LONG) ? 0x54ffu : 0x5432u;       // received.high | own.low
    const uint32_t mergeLo = (t == 0 && !LONG) ? 0x54ffu : 0x5410u;       // received.low  | own.low
    const uint32_t mergeC = t == 0 ? (LONG ? 0x5410u : 0x54ffu) : 0x5432u; // column maximum: the boundary word carries it in its LOW half
    const uint32_t first01 = t == 0 ? 1u : 0u;
    uint32_t* const bR = LONG ? boundary + ((long long)blockIdx.x * NG + gid) * 2 * boundary_stride : nullptr;
    uint32_t* const bF = LONG ? bR + boundary_stride : nullptr;

    uint32_t H[KR], E[KR], sel[KR];
    uint32_t Ftop = Z, Hdtop = Z, cmin = Z;       // boundary values entering this thread's two stages at the next step
    uint32_t a = 0, b = 0;                        // matrix rows of the target bases under the low / high stage
    uint32_t best = Z, cvlo = 0, cvhi = 0;        // per-stage best score (packed) and the first step of the block in which it last improved
    uint32_t tchunk = 0, tnext = 0;               // matrix rows of target bases [kG + t] of the current / next chunk (tchunk rotates down the group by one lane per step)
    uint32_t tbyte = TB_NONE;                     // code of target base [(k+2)G + t]: loaded two chunks ahead, looked up one chunk ahead, so neither latency is waited for
    uint32_t rchunk = Z, fchunk = Z, rnext = Z, fnext = Z, hb_prev = Z;   // LONG: boundary words of columns [kG + 1 + t] (rotating like tchunk) / of the next chunk; thread 0: record of the previous column
    int s = 0, nsteps = 0, rowoff = 0;            // rowoff: read row of the strip's first row (negative: dead rows on top)
    int rf_len = 0, tdir = 1, tout = 0, wide = 0;
    int strip = 0, nstrips = 1, rd_len_cur = 0;
    uint32_t stop2 = 0;                           // reverse passes: score at which the pass may end, in both halves (0: never)
    int64_t rf_base = 0, rd_base = 0, cm_off = -1;
    uint32_t ck_off = 0, cstart = 0, best0 = Z;   // byte offset of the scratch slot (the committed one is the other), first step of the committed block, best at the last commit test
    int run_S = 0, run_row = 0;                   // best (score, column, row) over the strips done so far
    unsigned long long run_key = ~0ull;
    bool active = true;                           // group still has (or may fetch) a task
    const int relist_n = NM ? relist[0] : 0;      // NM: number of flagged pairs to walk

#pragma unroll
    for (int j = 0; j < KR; ++j) { H[j] = Z; E[j] = Z; sel[j] = 0x8888u | 0x4400u; }

    // One target column for this thread's two stages: KR packed cells.  aa / bb: matrix rows of the target bases under the low / high stage.
    // In: H, E (previous column), Ftop, Hdtop.  Out: H, E (this column), F = F leaving the bottom row, Hdtop = bottom H of the PREVIOUS
    // column (the diagonal for the stage below), m = maximum over the rows and m0.
    auto dp_column = [&](const uint32_t aa, const uint32_t bb, uint32_t& F, uint32_t& m, const uint32_t m0) {
        auto score = [&](const int j) -> uint32_t {
            const uint32_t v = prmt(aa, bb, sel[j]);
            return NM ? prmt(v, sc.ncol2, nfix[j * STRIP_BLOCK + tid]) : v;      // N variant: rows holding an N take the N column's constant
        };
        F = Ftop; m = m0;
        uint32_t h = AR::hadd(Hdtop, score(0));
#pragma unroll
        for (int j = 0; j < KR; ++j) {
            uint32_t hnext = 0;
            if (j + 1 < KR) hnext = AR::hadd(H[j], score(j + 1));        // uses H(j) of the previous column: diagonal of row j+1
            else Hdtop = H[j];                                            // bottom H of the previous column: diagonal for the next stage
            const uint32_t Hn = AR::hmax(h, E[j], F);
            const uint32_t Hg = add2(Hn, sc.mgapO2);
            E[j] = AR::ef(E[j], sc.mgapE2, Hg);
            F = AR::ef(F, sc.mgapE2, Hg);
            H[j] = Hn;
            if (j & 1) m = AR::mx3(m, H[j - 1], Hn);
            else if (j == KR - 1) m = AR::mx2(m, Hn);                     // odd KR: the last row has no partner
            h = hnext;
        }
    };

    // selectors of this thread's rows for the strip starting at read row `rowoff`; flags reads the variant cannot score
    auto build_selectors = [&]() {
        const int r0 = rowoff + 2 * t * KR;                              // this thread's first row; it owns 2*KR consecutive rows
        if (!NM && r0 >= 0) {
            // common case, no dead row in this thread: the 2*KR read bases are one contiguous span -> a few aligned 64-bit loads
            // instead of 2*KR byte loads.  Forward passes walk the read upwards (span starts at row r0), reverse passes downwards
            // (span starts at the LAST row of the thread; row k sits at byte 2*KR-1-k).
            constexpr int NB = 2 * KR, NWORD = (NB + 7) / 8;
            const int8_t* span = REV ? seq + rd_base - (int64_t)(r0 + NB - 1) : seq + rd_base + r0;
            unsigned long long wq[NWORD];
#pragma unroll
            for (int q = 0; q < NWORD; ++q) wq[q] = load8_aligned(span + 8 * q, NB - 8 * q);
            unsigned long long bad = 0;
#pragma unroll
            for (int q = 0; q < NWORD; ++q) bad |= wq[q] & 0xfcfcfcfcfcfcfcfcull;          // any code >= 4 (N): flagged, redone by the N variant or the 32-bit kernel
            if (bad != 0ull) wide = 1;
#pragma unroll
            for (int j = 0; j < KR; ++j) {
                const int b_lo = REV ? NB - 1 - j : j, b_hi = REV ? NB - 1 - (j + KR) : j + KR;
                const uint32_t q_lo = (uint32_t)(wq[b_lo >> 3] >> (8 * (b_lo & 7))) & 3u, q_hi = (uint32_t)(wq[b_hi >> 3] >> (8 * (b_hi & 7))) & 3u;
                sel[j] = (q_lo * 0x11u + 0x80u) | ((q_hi * 0x11u + 0xc4u) << 8);
                H[j] = Z; E[j] = Z;
            }
        } else {
#pragma unroll
            for (int j = 0; j < KR; ++j) {
                const int r_lo = r0 + j, r_hi = r_lo + KR;
                uint32_t n_lo = 0x88u, n_hi = 0xccu;                      // dead row: sign bytes only -> score 0 or -1
                uint32_t fix = 0x3210u;                                   // NM: identity = keep the looked-up score
                if (r_lo >= 0) {
                    const int q = seq[rd_base + (int64_t)tdir * r_lo];
                    if ((unsigned)q < 4u) n_lo = (uint32_t)q | ((uint32_t)(q | 8) << 4);
                    else if (NM && q == 4) fix = (fix & 0xff00u) | 0x54u;      // low half <- the N column's constant
                    else wide = 1;
                }
                if (r_hi >= 0) {
                    const int q = seq[rd_base + (int64_t)tdir * r_hi];
                    if ((unsigned)q < 4u) n_hi = (uint32_t)(q | 4) | ((uint32_t)(q | 12) << 4);
                    else if (NM && q == 4) fix = (fix & 0x00ffu) | 0x7600u;    // high half <- the N column's constant
                    else wide = 1;
                }
                sel[j] = n_lo | (n_hi << 8);
                H[j] = Z; E[j] = Z;
                if (NM) nfix[j * STRIP_BLOCK + tid] = fix;
            }
        }
    };

    // state of a strip before its first step (s = 0): target chunk 0 / 1, boundary of column 0 and chunk 0 of the strip boundary
    auto start_strip = [&]() {
        Ftop = Hdtop = cmin = Z; a = b = 0; best = best0 = Z; cvlo = cvhi = 0;
        s = 0;
        ck_off = 0; cstart = 0;
        uint32_t mr = 0;
        if (t < rf_len) mr = smatrow[seq[rf_base + (int64_t)tdir * t] & 7];
        tnext = mr;
        tbyte = (G + t < rf_len) ? (uint32_t)(uint8_t)seq[rf_base + (int64_t)tdir * (G + t)] : TB_NONE;
        if (LONG) {
            rnext = fnext = Z; hb_prev = Z;
            if (strip > 0 && rf_len > 0) {
                const uint32_t r0w = bR[0], f0w = bF[0];
                if (t == 0) {   // column 0 of the strip above: F and the column maximum enter directly, the diagonal H(., -1) is the matrix edge
                    Ftop = prmt(f0w, Z, 0x5432u); cmin = prmt(r0w, Z, 0x5410u); hb_prev = r0w;
                }
                if (1 + t < rf_len) { rnext = bR[1 + t]; fnext = bF[1 + t]; }
            }
        }
    };

    for (;;) {
        // ------------------------------------------------------------------ task boundary (every G steps) -------------
        if (REV) {   // early end of a reverse pass (ssw.c:281 / :483): once some stage has seen the terminating score in column c, every
            // stage has passed column c after at most 2G-1 further steps.  A stage knows the BLOCK of steps in which it reached the score
            // (cvlo / cvhi = first step of that block), so the bound below is up to STRIP_CK - 1 steps late -- harmless, the smallest
            // (column, stage) wins whatever else is computed.
            const uint32_t x = best ^ stop2;
            const bool hit_lo = stop2 != 0u && (x & 0xffffu) == 0u, hit_hi = stop2 != 0u && (x >> 16) == 0u;
            const unsigned hits = __ballot_sync(0xffffffffu, hit_lo || hit_hi) & gmask;
            if (hits != 0u) {
                int col = 0x3fffffff;
                if (hit_lo) col = (int)cvlo + CK - 1 - 2 * t;
                if (hit_hi) col = min(col, (int)cvhi + CK - 1 - 2 * t - 1);
#pragma unroll
                for (int off = G / 2; off >= 1; off >>= 1) col = min(col, __shfl_xor_sync(gmask, col, off));
                nsteps = min(nsteps, col + 2 * G);
            }
        }
        if (active && s >= nsteps) {
            if (nsteps > 0) {
                // ---- finalize the strip.  Forward: `best` of a stage is the running maximum over its columns of the column maximum over the
                //      stages (and strips) up to and including it; reverse: over its own rows only.  Every thread holding the maximum S replays
                //      its committed block (the log stands in for the thread above) and looks for the first step at which one of its OWN cells
                //      equals S; the smallest (column, stage) over the group is the cell of ssw.c:260-277, and the smallest row of that
                //      stage's column holding S the row of ssw.c:284-293.
                const int sc_lo = AR::lo(best), sc_hi = AR::hi(best);
                int S = max(sc_lo, sc_hi);
#pragma unroll
                for (int off = G / 2; off >= 1; off >>= 1) S = max(S, __shfl_xor_sync(gmask, S, off));
                const unsigned anywide = __ballot_sync(gmask, wide != 0);
                constexpr unsigned long long NOKEY = ~0ull;
                unsigned long long key = NOKEY;          // (column << 8 | stage) of this thread's first own cell equal to S
                int row = 0;
                if (S > 0 && S >= run_S && (sc_lo == S || sc_hi == S)) {
                    const unsigned char* const cs = smem0 + (SLOT - ck_off);
                    const uint4* const cq = reinterpret_cast<const uint4*>(cs) + tid;
#pragma unroll
                    for (int k = 0; k < KRQ; ++k) {
                        const uint4 hq = cq[(size_t)k * STRIP_BLOCK], eq = cq[(size_t)(KRQ + k) * STRIP_BLOCK];
                        const uint32_t hw[4] = {hq.x, hq.y, hq.z, hq.w}, ew[4] = {eq.x, eq.y, eq.z, eq.w};
#pragma unroll
                        for (int q = 0; q < 4; ++q) if (4 * k + q < KR) { H[4 * k + q] = hw[q]; E[4 * k + q] = ew[q]; }
                    }
                    {
                        const uint2 fh = reinterpret_cast<const uint2*>(cs + SLOT_FH)[tid];
                        Ftop = fh.x; Hdtop = fh.y;
                    }
                    const uint32_t* const lg = reinterpret_cast<const uint32_t*>(cs + SLOT_LOG) + tid;
                    // the first hit is the best one of this thread: later steps mean larger columns (a hit in both stages at once: the high
                    // stage is one column behind, so it wins)
                    for (int sr = (int)cstart, k = 0; k < CK; ++sr, ++k) {
                        const int ia = sr - 2 * t, ib = ia - 1;
                        const uint32_t ra = (ia >= 0 && ia < rf_len) ? smatrow[seq[rf_base + (int64_t)tdir * ia] & 7] : 0u;
                        const uint32_t rb = (ib >= 0 && ib < rf_len) ? smatrow[seq[rf_base + (int64_t)tdir * ib] & 7] : 0u;
                        uint32_t F, m;
                        dp_column(ra, rb, F, m, Z);
                        const bool hit_lo = AR::lo(m) == S, hit_hi = AR::hi(m) == S;
                        if (hit_lo || hit_hi) {
                            key = hit_hi ? (((unsigned long long)(unsigned)ib << 8) | (unsigned)(2 * t + 1)) : (((unsigned long long)(unsigned)ia << 8) | (unsigned)(2 * t));
#pragma unroll
                            for (int j = KR - 1; j >= 0; --j) {
                                const int hv = hit_hi ? AR::hi(H[j]) : AR::lo(H[j]);
                                if (hv == S) row = rowoff + (2 * t + (hit_hi ? 1 : 0)) * KR + j;
                            }
                            break;
                        }
                        const uint32_t lw = lg[(size_t)k * STRIP_BLOCK];
                        Ftop = prmt(lw, F, mergeLo);          // low half <- F the thread above handed down, high half <- own low stage
                        Hdtop = prmt(lw, Hdtop, mergeHi);     // same for the diagonal H
                    }
                }
                unsigned long long wkey = key;
#pragma unroll
                for (int off = G / 2; off >= 1; off >>= 1) {
                    const unsigned long long o = __shfl_xor_sync(gmask, wkey, off);
                    wkey = o < wkey ? o : wkey;
                }
                // the strip's cell against the strips done so far: larger score, then smaller column; ties keep the earlier strip (smaller rows)
                int wrow = 0;
                if (wkey != NOKEY) {
                    const int wt = (int)(wkey & 0xffu) >> 1;
                    wrow = __shfl_sync(gmask, row, (lane - t) + wt);
                }
                const bool lost = S > run_S && wkey == NOKEY;      // guard: a new maximum whose cell was not found in the committed blocks -> redo the pair in the 32-bit kernel
                if (wkey != NOKEY && (S > run_S || (S == run_S && (wkey >> 8) < (run_key >> 8)))) { run_S = S; run_key = wkey; run_row = wrow; }
                if (anywide || lost) wide = 1;               // (group-uniform: every thread of the group carries the verdict to the last strip)
                if (LONG && REV && nsteps < rf_len + 2 * G - 1) rf_len = min(rf_len, nsteps - (2 * G - 1));   // a strip that ended early bounds the terminating column for the strips below it
                if (strip + 1 >= nstrips && t == 0) {
                    SwEnds e;
                    e.score = run_S;
                    e.col = run_S > 0 ? (int)(run_key >> 8) : -1;
                    e.row = run_S > 0 ? run_row : 0;
                    e.flags = wide ? SW_FLAG_NEEDS_WIDE : 0;
                    // a reverse pass that did not reach its terminating score (cannot happen for a symmetric recurrence; kept as a guard): same
                    if (REV && stop2 != 0u && (uint32_t)(run_S + (LONG ? (int)LBIAS : 0)) != (stop2 & 0xffffu)) e.flags = SW_FLAG_NEEDS_WIDE;
                    out[tout] = e;
                }
            }
            if (LONG && nsteps > 0 && strip + 1 < nstrips) {
                // ---- next strip of the same pair
                ++strip;
                rowoff += CAP;
                build_selectors();
                nsteps = rf_len > 0 ? rf_len + 2 * G - 1 : 1;
                start_strip();
            } else {
                // ---- fetch the next task of this group
                int ti = 0;
                bool have = false;
                if (NM) {
                    // walk the list of flagged pairs: take those of this variant's length class that are still flagged
                    for (;;) {
                        int k = 0;
                        if (t == 0) k = atomicAdd(counter, 1);
                        k = __shfl_sync(gmask, k, lane - t);
                        if (k >= relist_n) break;
                        ti = relist[1 + k];
                        const int len = tasks[ti].rd_len;
                        if (len > aux && len <= CAP && (out[tasks[ti].out].flags & SW_FLAG_NEEDS_WIDE)) { have = true; break; }
                    }
                } else {
                    if (t == 0) ti = atomicAdd(counter, 1);
                    ti = __shfl_sync(gmask, ti, lane - t);
                    have = ti < ntasks;
                }
                if (!have) {
                    active = false;
                    rf_len = 0; nsteps = 0; cm_off = -1; stop2 = 0; strip = 0; nstrips = 1;
#pragma unroll
                    for (int j = 0; j < KR; ++j) { H[j] = Z; E[j] = Z; sel[j] = 0x8888u | 0x4400u; }
                    Ftop = Hdtop = cmin = Z; a = b = 0; best = best0 = Z; tnext = 0; tchunk = 0; tbyte = TB_NONE;
                    rchunk = fchunk = rnext = fnext = hb_prev = Z;
                } else {
                    const SwTask tk = tasks[ti];
                    rd_len_cur = tk.rd_len;
                    rf_len = tk.rf_len; tdir = tk.dir; tout = tk.out; rf_base = tk.rf_base; rd_base = tk.rd_base; cm_off = tk.cm_off;
                    const uint32_t stopv = (uint32_t)tk.stop + (LONG ? LBIAS : 0u);
                    stop2 = tk.stop > 0 ? (stopv | (stopv << 16)) : 0u;
                    strip = 0;
                    nstrips = LONG ? max(1, (rd_len_cur + CAP - 1) / CAP) : 1;
                    rowoff = rd_len_cur - nstrips * CAP;                          // minus the dead rows on top of the first strip
                    wide = 0;
                    run_S = 0; run_key = ~0ull; run_row = 0;
                    build_selectors();
                    nsteps = (rf_len > 0 && rd_len_cur > 0) ? rf_len + 2 * G - 1 : 1;
                    if (!LONG && rd_len_cur > CAP) { wide = 1; nsteps = 1; rf_len = 0; }      // host scheduling error: never index out of the strip
                    {   // a pair this variant cannot score is redone completely elsewhere: stop after one block of steps, and (main variant,
                        // constant N column) put it on the list the N variants walk
                        const unsigned refused = __ballot_sync(gmask, wide != 0);
                        if (refused != 0u) {
                            nsteps = 1; rf_len = 0; nstrips = 1;
                            if (!NM && relist != nullptr && lane == __ffs((int)refused) - 1) relist[1 + atomicAdd(&relist[0], 1)] = aux + ti;
                        }
                    }
                    start_strip();
                }
            }
        }
        if (!__any_sync(0xffffffffu, active)) break;

        // rotate the target chunk: tchunk <- chunk s/G, look up chunk s/G + 1, load the codes of chunk s/G + 2
        //      (the loaded byte is not touched until the next rotation: nothing here waits for the load)
        tchunk = tnext;
        tnext = smatrow[tbyte & 15u];
        {
            const int idx = s + 2 * G + t;
            tbyte = TB_NONE;
            if (idx < rf_len) tbyte = (uint32_t)(uint8_t)seq[rf_base + (int64_t)tdir * idx];
        }
        if (LONG) {     // ... and the strip boundary: words of columns s + 1 + t now, those of the next chunk requested a whole block ahead
            rchunk = rnext; fchunk = fnext;
            rnext = fnext = Z;
            const int idx = s + G + 1 + t;
            if (strip > 0 && idx < rf_len) { rnext = bR[idx]; fnext = bF[idx]; }
        }

        // ------------------------------------------------------------------ G wavefront steps ---------------------------
        auto step = [&](const int u) {
            // the first stage takes the next target base from the chunk (which moves down the group by one lane per step, so the base is
            // always in the thread's own register), the others got theirs by shuffle last step
            a = blend_first(a, tchunk, first01);
            tchunk = __shfl_down_sync(0xffffffffu, tchunk, 1, G);
            uint32_t F, m;
            dp_column(a, b, F, m, REV ? Z : cmin);
            // forward: m already holds the column maximum over the stages (and strips) up to this one (dp_column starts from cmin); the
            // best of a stage is tracked on that, without positions -- the replay finds them (see finalize).  Reverse: own rows only.
            best = AR::mx2(best, m);
            const uint32_t cmout = REV ? Z : m;
            if (STAGED) {
                // ---- (column maximum, bottom-row H) of this step, from the last stage only: a finished column (s - (2G-1)); the group
                //      writes G of them to global memory after the loop.  LONG: also the F that leaves the strip.
                if (t == G - 1) {
                    crow[u * NG + gid] = prmt(cmout, H[KR - 1], 0x7632u);
                    if (LONG) frow[u * NG + gid] = F;
                }
            }
            // ---- hand the boundary to the next stage
            //      only the high halves (the thread's second stage) leave the thread: F and the diagonal H travel in one word
            uint32_t rX = __shfl_up_sync(0xffffffffu, prmt(F, Hdtop, 0x7632u), 1, G);
            const uint32_t rA = __shfl_up_sync(0xffffffffu, b, 1, G);
            uint32_t rC = 0;
            if (!REV) rC = __shfl_up_sync(0xffffffffu, cmout, 1, G);
            if (LONG) {
                // the first thread takes the strip above instead: F of the next column, bottom H of this column (the diagonal), and the
                // running column maximum of the next column (low half of the record)
                rX = blend_first(rX, prmt(fchunk, hb_prev, 0x7632u), first01);
                if (!REV) rC = blend_first(rC, rchunk, first01);
                hb_prev = rchunk;
                rchunk = __shfl_down_sync(0xffffffffu, rchunk, 1, G);
                fchunk = __shfl_down_sync(0xffffffffu, fchunk, 1, G);
            }
            // log what came from above: the replay of this block runs without shuffles
            reinterpret_cast<uint32_t*>(smem0 + ck_off + SLOT_LOG)[(u % CK) * STRIP_BLOCK + tid] = rX;
            Ftop = prmt(rX, F, mergeLo);
            Hdtop = prmt(rX, Hdtop, mergeHi);
            if (!REV) cmin = prmt(rC, cmout, mergeC);
            b = a;
            a = rA;
        };
        for (int u0 = 0; u0 < G; u0 += CK) {
            {
                // ---- checkpoint into the scratch slot: H / E of this thread's rows and the boundary entering its stages at step s
                uint4* const cq = reinterpret_cast<uint4*>(smem0 + ck_off) + tid;
#pragma unroll
                for (int k = 0; k < KRQ; ++k) {
                    cq[(size_t)k * STRIP_BLOCK] = make_uint4(H[4 * k], H[min(4 * k + 1, KR - 1)], H[min(4 * k + 2, KR - 1)], H[min(4 * k + 3, KR - 1)]);
                    cq[(size_t)(KRQ + k) * STRIP_BLOCK] = make_uint4(E[4 * k], E[min(4 * k + 1, KR - 1)], E[min(4 * k + 2, KR - 1)], E[min(4 * k + 3, KR - 1)]);
                }
                reinterpret_cast<uint2*>(smem0 + ck_off + SLOT_FH)[tid] = make_uint2(Ftop, Hdtop);
            }
#pragma unroll STRIP_UNROLL
            for (int uu = 0; uu < CK; ++uu, ++s) step(u0 + uu);
            {
                // ---- commit: the block just done becomes the committed one iff the better of this thread's two stages (score, then
                //      earlier column, then the low stage) made its last improvement inside it
                //      (cvlo / cvhi hold the first step of the block in which the stage last improved; best0 = best at the last look)
                const uint32_t imp = best ^ best0;
                if (imp != 0u) {
                    const uint32_t s0 = (uint32_t)s - CK;
                    if (imp & 0xffffu) cvlo = s0;
                    if (imp >> 16) cvhi = s0;
                    const int sc_lo = AR::lo(best), sc_hi = AR::hi(best);
                    const bool lo_wins = sc_lo > sc_hi || (sc_lo == sc_hi && cvlo <= cvhi);          // same block: the replay looks at both stages
                    if ((lo_wins ? cvlo : cvhi) == s0) { ck_off = SLOT - ck_off; cstart = s0; }
                    best0 = best;
                }
            }
        }
        // ---- the G columns finished in these steps: thread t of the group writes the one of step s - G + t (one 4*G-byte run per group):
        //      the column record (forward, last strip) or the boundary words for the strip below (LONG)
        if (STAGED) {
            __syncwarp();
            const int c = s - G + t - (2 * G - 1);
            if (c >= 0 && c < rf_len) {
                const uint32_t rec = crow[t * NG + gid];
                if (LONG && strip + 1 < nstrips) { bR[c] = rec; bF[c] = frow[t * NG + gid]; }
                else if (!REV && cm_off >= 0) colrec[cm_off + c] = LONG ? (((rec & 0xffffu) - LBIAS) | (((rec >> 16) - LBIAS) << 16)) : rec;
            }
            __syncwarp();
        }
    }
}

}  // namespace mpn
