// Shared device-side definitions for the batched Smith-Waterman kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mpn {

// One score pass over one (read, target) pair in "processing order": forward passes walk both sequences upwards,
// reverse passes (begin-position search, ssw.c:820-832) walk them downwards from (read_end1, ref_end1).
struct SwTask {
    int64_t rd_base;   // index into the sequence arena of the first read base in processing order
    int64_t rf_base;   // index into the sequence arena of the first target base in processing order
    int64_t cm_off;    // word offset into the column-record arena (one uint32 per target column), or -1 for "do not record"
    int32_t rd_len;
    int32_t rf_len;
    int32_t dir;       // +1 forward, -1 reverse
    int32_t out;       // slot in the SwEnds array
    int32_t stop;      // > 0: the pass may end once a column maximum equals this score (ssw.c:281 / :483 `terminate`)
    int32_t pad_;
};

// Result of a score pass.  col/row are in processing order (for a reverse pass: distance from the end).
struct SwEnds {
    int32_t score;     // maximum H over the matrix (pad rows never exceed it)
    int32_t col;       // first column (processing order) whose maximum equals score; -1 if score == 0
    int32_t row;       // smallest row of that column holding score; 0 if score == 0
    int32_t flags;     // SW_FLAG_*
};
enum { SW_FLAG_NEEDS_WIDE = 1 };   // pair refused by the packed 16-bit kernel (read code >= 4): redone by its N variant or by the 32-bit kernel

// Scoring for the packed kernel: row t of the substitution matrix, entries for read codes 0..3, one byte each.
struct Score16 {
    uint32_t matrow[8];
    uint32_t mgapO2;   // (-gapO) in both 16-bit halves
    uint32_t mgapE2;   // (-gapE) in both 16-bit halves
    uint32_t ncol2;    // score of read code 4 (N) against every target code, in both halves -- valid if ncol_ok
    uint32_t ncol_ok;  // n == 5 and mat[t][4] is the same for all t: reads containing N can stay in the short-read packed kernel
};

// ---- packed s16x2 helpers.  All single SASS instructions on sm_100a (profiles/r01_ubench_cell_pipes.md):
//      VIADD.16x2 runs on the fmaheavy pipe, everything else here on the alu pipe (64 lanes/clk/SM).
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel)
{
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}
__device__ __forceinline__ uint32_t add2(uint32_t a, uint32_t b) { return __vadd2(a, b); }            // VIADD.16x2
__device__ __forceinline__ uint32_t max2(uint32_t a, uint32_t b)
{
    uint32_t d;
    asm("max.s16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
// relu(max(h, e, f)).  Written as max(max(e, f), h) so that ptxas emits VIADD.16x2 (fma pipe) for the producer of h and one
// VIMNMX3.S16x2.RELU here, instead of folding the add into a VIADDMNMX + VIMNMX pair (one more alu-pipe instruction).
__device__ __forceinline__ uint32_t max3_relu(uint32_t h, uint32_t e, uint32_t f)
{
    uint32_t t, d;
    asm("max.s16x2 %0, %1, %2;" : "=r"(t) : "r"(e), "r"(f));
    asm("max.s16x2.relu %0, %1, %2;" : "=r"(d) : "r"(t), "r"(h));
    return d;
}
// f01 ? x : y for f01 in {0, 1} as two multiply-adds (fma pipe) instead of a SEL (alu pipe, the busy one): y + f01 * (x - y)
__device__ __forceinline__ uint32_t blend_first(uint32_t y, uint32_t x, uint32_t f01)
{
    uint32_t d, r;
    asm("mad.lo.u32 %0, %1, 0xffffffff, %2;" : "=r"(d) : "r"(y), "r"(x));      // x - y
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(d), "r"(f01), "r"(y));
    return r;
}
__device__ __forceinline__ uint32_t max3(uint32_t a, uint32_t b, uint32_t c) { return __vimax3_s16x2(a, b, c); }             // VIMNMX3.S16x2
__device__ __forceinline__ uint32_t addmax_relu(uint32_t a, uint32_t b, uint32_t c) { return __viaddmax_s16x2_relu(a, b, c); } // VIADDMNMX.S16x2.RELU

// ---- unsigned twins with a bias, for the kernels that carry the int16 clamp of ssw.c:425 (every H / E / F stored as value + LBIAS)
constexpr uint32_t LBIAS = 512u;                 // > 2 * 255 (largest gapO + gapE of the uint8_t ABI)
constexpr uint32_t LBIAS2 = LBIAS | (LBIAS << 16);
constexpr uint32_t LCAP2 = (32767u + LBIAS) | ((32767u + LBIAS) << 16);
__device__ __forceinline__ uint32_t umax2(uint32_t a, uint32_t b) { uint32_t d; asm("max.u16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ uint32_t umax3(uint32_t a, uint32_t b, uint32_t c) { return __vimax3_u16x2(a, b, c); }
__device__ __forceinline__ uint32_t uaddmin(uint32_t a, uint32_t b, uint32_t c) { return __viaddmin_u16x2(a, b, c); }
__device__ __forceinline__ uint32_t uaddmax(uint32_t a, uint32_t b, uint32_t c) { return __viaddmax_u16x2(a, b, c); }

// up to eight consecutive arena bytes starting at p (bytes beyond `left` are zero): two aligned 64-bit loads + a funnel shift.
// The sequence arena starts 256-byte aligned and is padded by 16 bytes (engine.cu), so the aligned words around any base are readable.
__device__ __forceinline__ unsigned long long load8_aligned(const int8_t* __restrict__ p, int left)
{
    if (left <= 0) return 0ull;
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const unsigned long long* w = reinterpret_cast<const unsigned long long*>(a & ~(uintptr_t)7);
    const unsigned sh = (unsigned)(a & 7u) * 8u;
    const unsigned long long lo = w[0], hi = w[1];
    unsigned long long v = sh ? ((lo >> sh) | (hi << (64u - sh))) : lo;
    if (left < 8) v &= (1ull << (8 * left)) - 1ull;
    return v;
}

}  // namespace mpn
