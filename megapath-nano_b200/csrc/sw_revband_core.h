// Banded reverse pass, one pair per lane (sm_100a; the per-lane routine also compiles for the host so that tests/test_revband_core.py
// can run it against the compiled reference without a GPU -- test infrastructure only, the product calls it from sw_revband.cuh).
//
// What it replaces.  ssw_align finds the begin of the best alignment by a second, full score pass over the reversed prefixes
// read[0..read_end1] x ref[0..ref_end1] that stops at the first column whose maximum equals score1 (ssw.c:820-832; sw_sse2_byte /
// sw_sse2_word with `terminate`, ssw.c:281 / :483), and takes the smallest row of that column holding the maximum (ssw.c:284-293).
//
// Why a band is exact.  Let S = score1, L = read_end1 + 1, and let the substitution matrix be "match mt > 0 on the diagonal, mm < 0
// elsewhere".  (1) ref_end1 is the FIRST forward column whose maximum is S and read_end1 the smallest row of it holding S, so every
// alignment of score S inside the sub-rectangle ends in its corner: in reversed coordinates (i = rows from read_end1 down, j = columns
// from ref_end1 down) it starts in cell (0, 0).  (2) Such a path makes at most L diagonal moves, each worth at most mt.  If it ever
// is D cells below the main diagonal (i - j = -D: D more target bases than read bases) it has paid for >= D gap bases:
// S <= mt * L - gapO - (D - 1) * gapE.  If it is D cells above (i - j = +D) the D extra read bases are rows that cannot be diagonal
// moves as well: S <= mt * (L - D) - gapO - (D - 1) * gapE.  With X = mt * L - S (requires gapO >= gapE >= 1):
//        D_below <= wd = (X - gapO + gapE) / gapE,        D_above <= wi = (X - gapO + gapE) / (mt + gapE).
// (3) A DP restricted to the band -wd <= i - j <= wi (cells outside read as 0) computes, in every cell, the maximum over a SUBSET of
// the alignments ending there: values never exceed the true H, and every cell of a score-S path keeps its true value because the whole
// path lies in the band.  S is the global maximum, so the cells equal to S are exactly the reference's, and so are "first column" and
// "smallest row".  Checked pair by pair against the compiled reference in tests/test_revband_core.py (degenerate alphabets included).
//
// Layout.  The band is walked by anti-diagonals k = i + j, two per step m (k = 2m, 2m + 1).  A lane keeps T = 2 NW band slots, two per
// register (s16x2): on even diagonals slot t is the cell with i - j = 2t - 2 h0, on odd diagonals 2t - 2 h0 + 1 (h0 = wd / 2 rounded
// up).  The diagonal neighbour (k - 2) is the same slot; of the two gap neighbours (k - 1) one is the same slot and the other the next
// slot down / up (one PRMT per register).  Scores come from one PRMT table look-up: the read bases of the T slots and the target
// bases of the T slots sit in two byte arrays (4 slots per register) that shift by one byte per step in opposite directions; a byte
// holds the base code in both nibbles, the target's with bit 7 set, so that their XOR is the selector pair (x, x | 8) PRMT needs to
// fetch the low byte of score x = r ^ f and its sign extension.  Cells outside the matrix read a sentinel base that mismatches
// everything: they stay at 0 before the matrix and can never reach S after it.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define MPN_HD __host__ __device__ __forceinline__
#else
#define MPN_HD inline
#endif

namespace mpn {
namespace rb {

struct Score {
    uint32_t tlo, thi;       // PRMT table: byte 0 = mt, bytes 1..7 = mm
    uint32_t mgo2, mge2;     // -gapO, -gapE in both halves
    int32_t mt, gapO, gapE;
    int32_t n_is_mismatch;   // every matrix entry involving code 4 equals mm: N needs no special case.  Otherwise a pair with an N bails out
};

constexpr uint32_t READ_SENT = 0x44444444u;      // code 4 in both nibbles
constexpr uint32_t REF_SENT = 0xD5D5D5D5u;       // code 5 in both nibbles, bit 7 set

// ---- packed arithmetic: single SASS instructions on sm_100a, plain C++ on the host
MPN_HD int16_t lo16(uint32_t v) { return (int16_t)(v & 0xffffu); }
MPN_HD int16_t hi16(uint32_t v) { return (int16_t)(v >> 16); }
MPN_HD uint32_t pack16(int lo, int hi) { return ((uint32_t)lo & 0xffffu) | ((uint32_t)hi << 16); }
MPN_HD int imax(int a, int b) { return a > b ? a : b; }
MPN_HD int imin(int a, int b) { return a < b ? a : b; }

MPN_HD uint32_t rprmt(uint32_t a, uint32_t b, uint32_t sel)
{
#if defined(__CUDA_ARCH__)
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
#else
    const uint64_t src = ((uint64_t)b << 32) | a;
    uint32_t d = 0;
    for (int q = 0; q < 4; ++q) {
        const uint32_t s = (sel >> (4 * q)) & 15u;
        uint32_t byte = (uint32_t)(src >> (8 * (s & 7u))) & 0xffu;
        if (s & 8u) byte = (byte & 0x80u) ? 0xffu : 0u;
        d |= byte << (8 * q);
    }
    return d;
#endif
}
MPN_HD uint32_t radd2(uint32_t a, uint32_t b)
{
#if defined(__CUDA_ARCH__)
    return __vadd2(a, b);
#else
    return pack16((int16_t)(lo16(a) + lo16(b)), (int16_t)(hi16(a) + hi16(b)));
#endif
}
MPN_HD uint32_t rmax3_relu(uint32_t a, uint32_t b, uint32_t c)
{
#if defined(__CUDA_ARCH__)
    // max(max(b, c), a) with the relu on the outer max: ptxas keeps the producer of `a` a VIADD.16x2 (fma pipe) and emits one
    // VIMNMX3.S16x2.RELU, instead of folding the add into a VIADDMNMX + VIMNMX pair (two alu-pipe instructions) -- see sw_common.cuh
    uint32_t t, d;
    asm("max.s16x2 %0, %1, %2;" : "=r"(t) : "r"(b), "r"(c));
    asm("max.s16x2.relu %0, %1, %2;" : "=r"(d) : "r"(t), "r"(a));
    return d;
#else
    return pack16(imax(0, imax(lo16(a), imax(lo16(b), lo16(c)))), imax(0, imax(hi16(a), imax(hi16(b), hi16(c)))));
#endif
}
MPN_HD uint32_t rmax3(uint32_t a, uint32_t b, uint32_t c)
{
#if defined(__CUDA_ARCH__)
    return __vimax3_s16x2(a, b, c);
#else
    return pack16(imax(lo16(a), imax(lo16(b), lo16(c))), imax(hi16(a), imax(hi16(b), hi16(c))));
#endif
}
// max(a + b, c) per half
MPN_HD uint32_t raddmax(uint32_t a, uint32_t b, uint32_t c)
{
#if defined(__CUDA_ARCH__)
    return __viaddmax_s16x2(a, b, c);
#else
    return pack16(imax((int16_t)(lo16(a) + lo16(b)), lo16(c)), imax((int16_t)(hi16(a) + hi16(b)), hi16(c)));
#endif
}
MPN_HD uint32_t rfunnel_r8(uint32_t lo, uint32_t hi)      // (hi:lo) >> 8
{
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, 8);
#else
    return (lo >> 8) | (hi << 24);
#endif
}
MPN_HD uint32_t rfunnel_l8(uint32_t lo, uint32_t hi)      // high word of (hi:lo) << 8
{
#if defined(__CUDA_ARCH__)
    return __funnelshift_l(lo, hi, 8);
#else
    return (hi << 8) | (lo >> 24);
#endif
}
MPN_HD uint64_t rbswap64(uint64_t v)
{
#if defined(__CUDA_ARCH__)
    const uint32_t lo = (uint32_t)v, hi = (uint32_t)(v >> 32);
    return ((uint64_t)__byte_perm(lo, 0, 0x0123) << 32) | __byte_perm(hi, 0, 0x0123);
#else
    return __builtin_bswap64(v);
#endif
}

// up to eight consecutive arena bytes starting at p (bytes beyond `left` are zero): two aligned 64-bit loads + a funnel shift.
// The sequence arena starts 256-byte aligned and is padded by 16 bytes (engine.cu), so the aligned words around any base are readable.
MPN_HD uint64_t rload8(const int8_t* p, int left)
{
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint64_t* w = reinterpret_cast<const uint64_t*>(a & ~(uintptr_t)7);
    const unsigned sh = (unsigned)(a & 7u) * 8u;
    const uint64_t lo = w[0], hi = w[1];
    uint64_t v = sh ? ((lo >> sh) | (hi << (64u - sh))) : lo;
    if (left < 8) v &= (1ull << (8 * left)) - 1ull;
    return v;
}

// Selector bytes of the stream positions idx0 .. idx0 + 7 of one sequence.  Position p is arena byte base - p (reverse passes walk
// downwards); positions outside [0, len) read the sentinel.  `bail` is set if a base with code >= 4 shows up and N is not a plain mismatch.
MPN_HD uint64_t chunk(const int8_t* seq, int64_t base, int idx0, int len, bool is_ref, int n_is_mismatch, bool& bail)
{
    const uint64_t sent = is_ref ? (((uint64_t)REF_SENT << 32) | REF_SENT) : (((uint64_t)READ_SENT << 32) | READ_SENT);
    const int q_lo = imax(0, -idx0), q_hi = imin(7, len - 1 - idx0);
    if (q_lo > q_hi) return sent;
    const int cnt = q_hi - q_lo + 1;
    uint64_t v = rload8(seq + (base - idx0 - q_hi), cnt);                 // byte c <-> position idx0 + q_hi - c
    v = rbswap64(v) >> (8 * (7 - q_hi));                                    // byte q <-> position idx0 + q
    const uint64_t vm = ((~0ull) >> (8 * (8 - cnt))) << (8 * q_lo);
    const uint64_t nmask = v & 0x0404040404040404ull;
    v &= 0x0303030303030303ull;
    if (nmask) {
        if (!n_is_mismatch) bail = true;
        v |= nmask | (is_ref ? (nmask >> 2) : 0ull);                        // read N = 4, target N = 5
    }
    const uint32_t lo = (uint32_t)v * 0x11u, hi = (uint32_t)(v >> 32) * 0x11u;
    uint64_t enc = ((uint64_t)hi << 32) | lo;
    if (is_ref) enc |= 0x8080808080808080ull;
    return (enc & vm) | (sent & ~vm);
}

// One pair.  rd_base / rf_base: arena index of read_end1 / ref_end1; L, C: rows and columns of the reversed sub-rectangle; S: score1;
// h0: band slots below the main diagonal / 2.  Returns 0 and (col, row) of the reference's terminating cell in reversed coordinates,
// or 1 if the pair has to be done by the full-matrix kernel (an N the table cannot express, or -- guard, cannot happen -- S not found).
template <int NW>
MPN_HD int lane(const int8_t* seq, int64_t rd_base, int64_t rf_base, int L, int C, int S, int h0, const Score& sc, int& col, int& row)
{
    constexpr int T = 2 * NW, NB = NW / 2;
    static_assert(NW % 4 == 0, "slots come in octets");
    uint32_t He[NW], Ho[NW], E[NW], F[NW], Ra[NB], Rb[NB];
#pragma unroll
    for (int u = 0; u < NW; ++u) He[u] = Ho[u] = E[u] = F[u] = 0u;
#pragma unroll
    for (int v = 0; v < NB; ++v) { Ra[v] = READ_SENT; Rb[v] = REF_SENT; }
    const uint32_t S2 = pack16(S, S);
    const int dmax = 2 * T - 1 - 2 * h0;              // largest i - j of a band slot
    bool bail = false;
    int ia = -h0, jb = h0 - T + 1;                    // next stream positions: Ra holds a[m - h0 + t], Rb holds b[m + h0 - t]

    auto push_a = [&](uint32_t word, int q) {         // shift the read bytes down one slot, byte q of `word` enters at the top
#pragma unroll
        for (int v = 0; v + 1 < NB; ++v) Ra[v] = rfunnel_r8(Ra[v], Ra[v + 1]);
        Ra[NB - 1] = rprmt(Ra[NB - 1], word, 0x0321u | ((uint32_t)(4 + q) << 12));
    };
    auto push_b = [&](uint32_t word, int q) {         // shift the target bytes up one slot, byte q of `word` enters at the bottom
#pragma unroll
        for (int v = NB - 1; v >= 1; --v) Rb[v] = rfunnel_l8(Rb[v - 1], Rb[v]);
        Rb[0] = rprmt(Rb[0], word, 0x2100u | (uint32_t)(4 + q));
    };

    // fill: T pushes each
    for (int o = 0; o < T / 8; ++o) {
        const uint64_t ca = chunk(seq, rd_base, ia, L, false, sc.n_is_mismatch, bail);
        const uint64_t cb = chunk(seq, rf_base, jb, C, true, sc.n_is_mismatch, bail);
        ia += 8; jb += 8;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            push_a(q < 4 ? (uint32_t)ca : (uint32_t)(ca >> 32), q & 3);
            push_b(q < 4 ? (uint32_t)cb : (uint32_t)(cb >> 32), q & 3);
        }
    }

    const int m_end = (L + C) / 2 + 1;                // diagonals 0 .. L + C - 2
    int m_stop = m_end;
    int best = 0x7fffffff;
    uint64_t ca = 0, cb = 0;
    uint64_t na = chunk(seq, rd_base, ia, L, false, sc.n_is_mismatch, bail);
    uint64_t nb = chunk(seq, rf_base, jb, C, true, sc.n_is_mismatch, bail);
    // one step = two anti-diagonals.  The loop is NOT unrolled: its body is 100-330 instructions per class, and an unrolled-by-8 version
    // (static byte selectors for the pushes) ran out of instruction cache -- `no_instruction` was the top stall of the three wider classes.
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int m = 0; m < m_stop && !bail; ++m) {
        if ((m & 7) == 0) {
            ca = na; cb = nb;
            ia += 8; jb += 8;
            na = chunk(seq, rd_base, ia, L, false, sc.n_is_mismatch, bail);       // one octet ahead: nothing below waits for these loads
            nb = chunk(seq, rf_base, jb, C, true, sc.n_is_mismatch, bail);
        }
        {
            uint32_t acc = 0u;
            // ---- even diagonal k = 2m: E side (i, j-1) is the same slot of the odd diagonal before, F side (i-1, j) the slot below
            {
                uint32_t fprev = 0u;
#pragma unroll
                for (int u = 0; u < NW; ++u) {
                    const uint32_t xw = Ra[u >> 1] ^ Rb[u >> 1];
                    const uint32_t s = rprmt(sc.tlo, sc.thi, (u & 1) ? (xw >> 16) : xw);
                    const uint32_t fo = F[u];
                    const uint32_t f_in = rprmt(fprev, fo, 0x5432u);
                    fprev = fo;
                    const uint32_t e_in = E[u];
                    const uint32_t h = rmax3_relu(radd2(He[u], s), e_in, f_in);
                    const uint32_t hgo = radd2(h, sc.mgo2);
                    E[u] = raddmax(e_in, sc.mge2, hgo);
                    F[u] = raddmax(f_in, sc.mge2, hgo);
                    He[u] = h;
                    if (u & 1) acc = rmax3(acc, He[u - 1], h);
                }
            }
            push_a((uint32_t)ca, 0);
            ca >>= 8;
            // ---- odd diagonal k = 2m + 1: F side is the same slot of the even diagonal before, E side the slot above
            {
#pragma unroll
                for (int u = 0; u < NW; ++u) {
                    const uint32_t xw = Ra[u >> 1] ^ Rb[u >> 1];
                    const uint32_t s = rprmt(sc.tlo, sc.thi, (u & 1) ? (xw >> 16) : xw);
                    const uint32_t e_in = rprmt(E[u], u + 1 < NW ? E[u + 1] : 0u, 0x5432u);
                    const uint32_t f_in = F[u];
                    const uint32_t h = rmax3_relu(radd2(Ho[u], s), e_in, f_in);
                    const uint32_t hgo = radd2(h, sc.mgo2);
                    E[u] = raddmax(e_in, sc.mge2, hgo);
                    F[u] = raddmax(f_in, sc.mge2, hgo);
                    Ho[u] = h;
                    if (u & 1) acc = rmax3(acc, Ho[u - 1], h);
                }
            }
            push_b((uint32_t)cb, 0);
            cb >>= 8;
            // ---- S reached on one of the two diagonals?  (no cell exceeds S, so the maximum equals S iff a cell does)
            const uint32_t x = acc ^ S2;
            if (((x - 0x00010001u) & ~x & 0x80008000u) != 0u) {
                // per diagonal the smallest column is the highest slot; (column, row) keys, smallest wins
                int key = 0x7fffffff;
#pragma unroll
                for (int t = T - 1; t >= 0; --t) {
                    const int v = (t & 1) ? hi16(He[t >> 1]) : lo16(He[t >> 1]);
                    const int i = m - h0 + t, j = m + h0 - t;
                    if (v == S && i >= 0 && j >= 0 && i < L && j < C) { key = (j << 16) | i; break; }
                }
#pragma unroll
                for (int t = T - 1; t >= 0; --t) {
                    const int v = (t & 1) ? hi16(Ho[t >> 1]) : lo16(Ho[t >> 1]);
                    const int i = m + 1 - h0 + t, j = m + h0 - t;
                    if (v == S && i >= 0 && j >= 0 && i < L && j < C) { key = imin(key, (j << 16) | i); break; }
                }
                if (key < best) {
                    best = key;
                    // a cell of a smaller column j' < j lies on a diagonal k' <= 2 j' + dmax
                    const int kmax = 2 * ((key >> 16) - 1) + dmax;
                    m_stop = imin(m_stop, kmax < 0 ? 0 : kmax / 2 + 1);
                }
            }
        }
    }
    if (bail || best == 0x7fffffff) return 1;
    col = best >> 16;
    row = best & 0xffff;
    return 0;
}

// band of a pair: 0 = not eligible, else NW (4, 8, 12, 16, 20) and h0
MPN_HD int classify(int L, int C, int S, const Score& sc, int& h0)
{
    h0 = 0;
    if (S <= 0 || L <= 0 || C <= 0 || L > 32767 || C > 32767 || S + sc.mt > 32767) return 0;
    const int X = sc.mt * L - S;
    if (X < 0) return 0;
    const int wd = X >= sc.gapO ? (X - sc.gapO + sc.gapE) / sc.gapE : 0;
    const int wi = X >= sc.gapO + sc.mt ? (X - sc.gapO + sc.gapE) / (sc.mt + sc.gapE) : 0;
    const int wde = (wd + 1) & ~1;
    const int W = wde + wi + 1;                        // values of i - j to cover: -wde .. wi
    h0 = wde / 2;
    if (W <= 16) return 4;
    if (W <= 32) return 8;
    if (W <= 48) return 12;
    if (W <= 64) return 16;
    if (W <= 80) return 20;
    return 0;
}

}  // namespace rb
}  // namespace mpn
