// Packed 16-bit inter-task score kernel (sm_100a): replaces the score passes of ssw.c -- sw_sse2_byte (ssw.c:123-328),
// sw_sse2_word (ssw.c:354-530) and their query profiles qP_byte / qP_word (ssw.c:89-114, 330-352) -- for pairs whose
// scores cannot reach the int16 clamp.
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
//   steps ends and the thread's better stage improved in it, scratch and committed slot swap.  At the end of the pass the thread that owns
//   the winning stage restores the committed slot and re-runs at most STRIP_CK steps on its own (the log replaces the shuffles) to get the
//   column again.  This replaces the per-improvement H-column snapshots of round 1 (8 predicated STS.128 + 15 register moves per step).
__host__ __device__ constexpr int strip_ck(int G) { return G < 8 ? G : 8; }
template <int KR, int G>
__host__ __device__ constexpr size_t strip16_slot_bytes() { return ((size_t)2 * ((KR + 3) / 4) * 16 + 8 + (size_t)strip_ck(G) * 4) * STRIP_BLOCK; }
// shared memory: 2 checkpoint slots, the column-record staging [G][STRIP_BLOCK] words (forward passes only) and, in the N variant only, the
// per-row score fix-up selectors [KR][STRIP_BLOCK]
template <int KR, int G, bool NM = false, bool REV = false>
__host__ __device__ constexpr size_t strip16_smem_bytes()
{
    return 2 * strip16_slot_bytes<KR, G>() + (size_t)((REV ? 0 : G) + (NM ? KR : 0)) * STRIP_BLOCK * sizeof(uint32_t);
}

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
template <int KR, int G, bool REV, bool NM = false>
__global__ void __launch_bounds__(STRIP_BLOCK, MPN_STRIP_MINB)
sw_strip16_kernel(const SwTask* __restrict__ tasks, int ntasks, int* __restrict__ counter, const int8_t* __restrict__ seq,
                  const Score16 sc, uint32_t* __restrict__ colrec, SwEnds* __restrict__ out, int* __restrict__ relist, int aux)
{
    static_assert(KR >= 2, "KR too small");
    constexpr int KRQ = (KR + 3) / 4;             // 16-byte words per checkpointed array
    static_assert(G == 2 || G == 4 || G == 8 || G == 16 || G == 32, "G must divide 32");
    constexpr int CAP = 2 * G * KR;               // rows covered by one strip
    constexpr int CK = strip_ck(G);               // forward: steps between checkpoints
    constexpr uint32_t SLOT = (uint32_t)strip16_slot_bytes<KR, G>();
    constexpr uint32_t SLOT_FH = 2 * KRQ * 16 * STRIP_BLOCK, SLOT_LOG = SLOT_FH + 8 * STRIP_BLOCK;     // byte offsets inside a slot
    extern __shared__ uint4 slots[];              // 2 checkpoint slots
    unsigned char* const smem0 = reinterpret_cast<unsigned char*>(slots);
    uint32_t* const crow = reinterpret_cast<uint32_t*>(smem0 + (size_t)2 * SLOT);         // forward only: [G][STRIP_BLOCK] column records of the last G steps
    uint32_t* const nfix = crow + (REV ? 0 : G) * STRIP_BLOCK;                            // NM only: [KR][STRIP_BLOCK] fix-up selectors

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
    const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (lane - t));
    // merge selector for (value received from thread t-1, own value): low half <- received.high, high half <- own.low.
    // For the first thread of a group the low half must be the matrix boundary (0): select the sign byte of own byte 7
    // (all merged quantities are >= 0, so that byte replicates to 0x00).
    const uint32_t mergeHi = (t == 0) ? 0x54ffu : 0x5432u;       // received.high | own.low
    const uint32_t mergeLo = (t == 0) ? 0x54ffu : 0x5410u;       // received.low  | own.low
    const uint32_t first01 = t == 0 ? 1u : 0u;

    uint32_t H[KR], E[KR], sel[KR];
    uint32_t Ftop = 0, Hdtop = 0, cmin = 0;       // boundary values entering this thread's two stages at the next step
    uint32_t a = 0, b = 0;                        // matrix rows of the target bases under the low / high stage
    uint32_t best = 0, cvlo = 0, cvhi = 0;        // per-stage best score (packed) and the step at which it was first reached
    uint32_t tchunk = 0, tnext = 0;               // matrix rows of target bases [kG + t] of the current / next chunk (tchunk rotates down the group by one lane per step)
    uint32_t tbyte = TB_NONE;                     // code of target base [(k+2)G + t]: loaded two chunks ahead, looked up one chunk ahead, so neither latency is waited for
    int s = 0, nsteps = 0, dead = 0;
    int rf_len = 0, tdir = 1, tout = 0, wide = 0;
    uint32_t stop2 = 0;                           // reverse passes: score at which the pass may end, in both halves (0: never)
    int64_t rf_base = 0, cm_off = -1;
    uint32_t ck_off = 0, cstart = 0, best0 = 0;   // forward: byte offset of the scratch slot (the committed one is the other), first step of the committed block, best at the last commit test
    bool active = true;                           // group still has (or may fetch) a task
    const int relist_n = NM ? relist[0] : 0;      // NM: number of flagged pairs to walk

#pragma unroll
    for (int j = 0; j < KR; ++j) { H[j] = 0; E[j] = 0; sel[j] = 0x8888u | 0x4400u; }

    // One target column for this thread's two stages: KR packed cells.  aa / bb: matrix rows of the target bases under the low / high stage.
    // In: H, E (previous column), Ftop, Hdtop.  Out: H, E (this column), F = F leaving the bottom row, Hdtop = bottom H of the PREVIOUS
    // column (the diagonal for the stage below), m = maximum over the rows.
    auto dp_column = [&](const uint32_t aa, const uint32_t bb, uint32_t& F, uint32_t& m, const uint32_t m0 = 0u) {
        auto score = [&](const int j) -> uint32_t {
            const uint32_t v = prmt(aa, bb, sel[j]);
            return NM ? prmt(v, sc.ncol2, nfix[j * STRIP_BLOCK + tid]) : v;      // N variant: rows holding an N take the N column's constant
        };
        F = Ftop; m = m0;
        uint32_t h = add2(Hdtop, score(0));
#pragma unroll
        for (int j = 0; j < KR; ++j) {
            uint32_t hnext = 0;
            if (j + 1 < KR) hnext = add2(H[j], score(j + 1));   // uses H(j) of the previous column: diagonal of row j+1
            else Hdtop = H[j];                                            // bottom H of the previous column: diagonal for the next stage
            const uint32_t Hn = max3_relu(h, E[j], F);
            const uint32_t Hg = add2(Hn, sc.mgapO2);
            E[j] = addmax_relu(E[j], sc.mgapE2, Hg);
            F = addmax_relu(F, sc.mgapE2, Hg);
            H[j] = Hn;
            if (j & 1) m = max3(m, H[j - 1], Hn);
            else if (j == KR - 1) m = max2(m, Hn);                        // odd KR: the last row has no partner
            h = hnext;
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
                // ---- finalize.  Forward: `best` of a stage is the running maximum over its columns of the column maximum over the stages up
                //      to and including it (reverse: over its own rows only).  Every thread holding the global maximum S replays its
                //      committed block (the log stands in for the thread above) and looks for the first step at which one of its OWN cells
                //      equals S; the smallest (column, stage) over the group is the cell of ssw.c:260-277, and the smallest row of that
                //      stage's column holding S the row of ssw.c:284-293.
                const int sc_lo = (int)(int16_t)(best & 0xffffu), sc_hi = (int)(int16_t)(best >> 16);
                int S = max(sc_lo, sc_hi);
#pragma unroll
                for (int off = G / 2; off >= 1; off >>= 1) S = max(S, __shfl_xor_sync(gmask, S, off));
                const unsigned anywide = __ballot_sync(gmask, wide != 0);
                constexpr unsigned long long NOKEY = ~0ull;
                unsigned long long key = NOKEY;          // (column << 8 | stage) of this thread's first own cell equal to S
                int row = 0;
                if (S > 0 && (sc_lo == S || sc_hi == S)) {
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
                        dp_column(ra, rb, F, m);
                        const bool hit_lo = (int)(int16_t)(m & 0xffffu) == S, hit_hi = (int)(int16_t)(m >> 16) == S;
                        if (hit_lo || hit_hi) {
                            key = hit_hi ? (((unsigned long long)(unsigned)ib << 8) | (unsigned)(2 * t + 1)) : (((unsigned long long)(unsigned)ia << 8) | (unsigned)(2 * t));
#pragma unroll
                            for (int j = KR - 1; j >= 0; --j) {
                                const int hv = hit_hi ? (int)(int16_t)(H[j] >> 16) : (int)(int16_t)(H[j] & 0xffffu);
                                if (hv == S) row = (2 * t + (hit_hi ? 1 : 0)) * KR + j - dead;
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
                if (S <= 0 ? t == 0 : (wkey == NOKEY ? t == 0 : key == wkey)) {
                    SwEnds e;
                    e.score = S;
                    e.col = S > 0 ? (int)(wkey >> 8) : -1;
                    e.row = S > 0 ? row : 0;
                    e.flags = anywide ? SW_FLAG_NEEDS_WIDE : 0;
                    if (S > 0 && wkey == NOKEY) e.flags = SW_FLAG_NEEDS_WIDE;      // guard: no cell found in the committed blocks -> redo the pair in the 32-bit kernel
                    // a reverse pass that did not reach its terminating score (cannot happen for a symmetric recurrence; kept as a guard): same
                    if (REV && stop2 != 0u && (uint32_t)S != (stop2 & 0xffffu)) e.flags = SW_FLAG_NEEDS_WIDE;
                    out[tout] = e;
                }
            }
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
                rf_len = 0; nsteps = 0; cm_off = -1; stop2 = 0;
#pragma unroll
                for (int j = 0; j < KR; ++j) { H[j] = 0; E[j] = 0; sel[j] = 0x8888u | 0x4400u; }
                Ftop = Hdtop = cmin = a = b = best = 0; tnext = 0; tchunk = 0; tbyte = TB_NONE;
            } else {
                const SwTask tk = tasks[ti];
                const int rd_len = tk.rd_len;
                rf_len = tk.rf_len; tdir = tk.dir; tout = tk.out; rf_base = tk.rf_base; cm_off = tk.cm_off;
                stop2 = tk.stop > 0 ? ((uint32_t)tk.stop | ((uint32_t)tk.stop << 16)) : 0u;
                dead = CAP - rd_len;
                wide = 0;
                // selectors: low half = row (2t)*KR + j, high half = row (2t+1)*KR + j, both minus the dead rows on top
                const int r0 = 2 * t * KR - dead;                                // this thread's first row; it owns 2*KR consecutive rows
                if (!NM && r0 >= 0 && rd_len <= CAP) {
                    // common case, no dead row in this thread: the 2*KR read bases are one contiguous span -> a few aligned 64-bit loads
                    // instead of 2*KR byte loads.  Forward passes walk the read upwards (span starts at row r0), reverse passes downwards
                    // (span starts at the LAST row of the thread; row k sits at byte 2*KR-1-k).
                    constexpr int NB = 2 * KR, NWORD = (NB + 7) / 8;
                    const int8_t* span = REV ? seq + tk.rd_base - (int64_t)(r0 + NB - 1) : seq + tk.rd_base + r0;
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
                        H[j] = 0; E[j] = 0;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < KR; ++j) {
                        const int r_lo = 2 * t * KR + j - dead, r_hi = r_lo + KR;
                        uint32_t n_lo = 0x88u, n_hi = 0xccu;                      // dead row: sign bytes only -> score 0 or -1
                        uint32_t fix = 0x3210u;                                   // NM: identity = keep the looked-up score
                        if (r_lo >= 0) {
                            const int q = seq[tk.rd_base + (int64_t)tdir * r_lo];
                            if ((unsigned)q < 4u) n_lo = (uint32_t)q | ((uint32_t)(q | 8) << 4);
                            else if (NM && q == 4) fix = (fix & 0xff00u) | 0x54u;      // low half <- the N column's constant
                            else wide = 1;
                        }
                        if (r_hi >= 0) {
                            const int q = seq[tk.rd_base + (int64_t)tdir * r_hi];
                            if ((unsigned)q < 4u) n_hi = (uint32_t)(q | 4) | ((uint32_t)(q | 12) << 4);
                            else if (NM && q == 4) fix = (fix & 0x00ffu) | 0x7600u;    // high half <- the N column's constant
                            else wide = 1;
                        }
                        sel[j] = n_lo | (n_hi << 8);
                        H[j] = 0; E[j] = 0;
                        if (NM) nfix[j * STRIP_BLOCK + tid] = fix;
                    }
                }
                Ftop = Hdtop = cmin = a = b = best = 0; cvlo = cvhi = 0;
                s = 0;
                nsteps = (rf_len > 0 && rd_len > 0) ? rf_len + 2 * G - 1 : 1;
                if (rd_len > CAP) { wide = 1; nsteps = 1; rf_len = 0; }      // host scheduling error: never index out of the strip
                {   // a pair this variant cannot score is redone completely elsewhere: stop after one block of steps, and (main variant,
                    // constant N column) put it on the list the N variants walk
                    const unsigned refused = __ballot_sync(gmask, wide != 0);
                    if (refused != 0u) {
                        nsteps = 1; rf_len = 0;
                        if (!NM && relist != nullptr && lane == __ffs((int)refused) - 1) relist[1 + atomicAdd(&relist[0], 1)] = aux + ti;
                    }
                }
                {   // matrix rows of target chunk 0, code of chunk 1
                    uint32_t mr = 0;
                    if (t < rf_len) mr = smatrow[seq[rf_base + (int64_t)tdir * t] & 7];
                    tnext = mr;
                    tbyte = (G + t < rf_len) ? (uint32_t)(uint8_t)seq[rf_base + (int64_t)tdir * (G + t)] : TB_NONE;
                }
                ck_off = 0; cstart = 0; best0 = 0;
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

        // ------------------------------------------------------------------ G wavefront steps ---------------------------
        auto step = [&](const int u) {
            // the first stage takes the next target base from the chunk (which moves down the group by one lane per step, so the base is
            // always in the thread's own register), the others got theirs by shuffle last step
            a = blend_first(a, tchunk, first01);
            tchunk = __shfl_down_sync(0xffffffffu, tchunk, 1, G);
            uint32_t F, m;
            dp_column(a, b, F, m, REV ? 0u : cmin);
            if (REV) best = max2(best, m);      // own rows only; positions come from the replay
            uint32_t cmout = 0;
            if (!REV) {
                // forward: m already holds the column maximum over the stages up to this one (dp_column starts from cmin); the best of a
                // stage is tracked on that, without positions -- the replay finds them (see finalize)
                cmout = m;
                best = max2(best, m);
                // ---- (column maximum, bottom-row H) of this step: staged in shared memory by every thread, only the last stage's
                //      entry is a finished column (s - (2G-1)); the group writes G of them to global memory after the loop
                crow[u * STRIP_BLOCK + tid] = prmt(cmout, H[KR - 1], 0x7632u);
            }
            // ---- hand the boundary to the next stage
            //      only the high halves (the thread's second stage) leave the thread: F and the diagonal H travel in one word
            const uint32_t rX = __shfl_up_sync(0xffffffffu, prmt(F, Hdtop, 0x7632u), 1, G);
            const uint32_t rA = __shfl_up_sync(0xffffffffu, b, 1, G);
            // log what came from the thread above: the replay of this block runs without shuffles
            reinterpret_cast<uint32_t*>(smem0 + ck_off + SLOT_LOG)[(u % CK) * STRIP_BLOCK + tid] = rX;
            Ftop = prmt(rX, F, mergeLo);
            Hdtop = prmt(rX, Hdtop, mergeHi);
            if (!REV) {
                const uint32_t rC = __shfl_up_sync(0xffffffffu, cmout, 1, G);
                cmin = prmt(rC, cmout, mergeHi);
            }
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
                    const int sc_lo = (int)(int16_t)(best & 0xffffu), sc_hi = (int)(int16_t)(best >> 16);
                    const bool lo_wins = sc_lo > sc_hi || (sc_lo == sc_hi && cvlo <= cvhi);          // same block: the replay looks at both stages
                    if ((lo_wins ? cvlo : cvhi) == s0) { ck_off = SLOT - ck_off; cstart = s0; }
                    best0 = best;
                }
            }
        }
        // ---- column records of the G steps just done: thread t of the group stores the one of step s - G + t (one 4*G-byte run per group)
        if (!REV) {
            __syncwarp();
            if (cm_off >= 0) {
                const int c = s - G + t - (2 * G - 1);
                const uint32_t v = crow[t * STRIP_BLOCK + tid - t + (G - 1)];
                if (c >= 0 && c < rf_len) colrec[cm_off + c] = v;
            }
            __syncwarp();
        }
    }
}

}  // namespace mpn
