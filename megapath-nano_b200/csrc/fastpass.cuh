// fastpass_kernel -- the region realigner's k-mer fast pass (reference realigner.cpp:170-253, :429-451) without a k-mer index.
//
// The reference hashes every 32-mer of the reads, walks the haplotype, and at every shared 32-mer compares the read ungapped at
// start = max(0, hap position - read offset) (<= 2 mismatches, N matches anything).  What that computes, per (haplotype, read):
//   * a placement start s is EVALUATED iff some diagonal d with max(0, d) == s holds 32 consecutive exactly-equal characters
//     (a "run"); it is first evaluated at scan time t = d + o of the first such run (o = its read offset);
//   * the read's placement is the passing start of highest score, ties to the earliest (t, o);
//   * a base p of the haplotype counts as covered at scan time i iff a passing start with s <= p < s + len was evaluated at a time <= i;
//   * a non-reference haplotype is dropped iff some position i inside the window has a 32-mer shared with ANY read (a run start
//     at hap position i on any diagonal) and is uncovered at time i.
// Here: every diagonal of every (haplotype, read) is compared bit-parallel.  Sequences are three bit planes (2 code bits + "is N";
// padding around the haplotype is the 8th symbol, equal to nothing).  One block per haplotype keeps, in shared memory, the
// haplotype's planes pre-shifted by every bit offset 0..31, so lane l owns the diagonals d == l (mod 32) and reads whole 32-base
// words of "its" alignment with one conflict-free LDS per plane and no shifts.  One warp per read, read planes in registers.
// Integer / byte work, bound by shared-memory loads and LOP3 issue -- no tensor-core shape here.
#pragma once
#include <cstdint>
#include <climits>

namespace mpn {

constexpr int FP_K = 32;
constexpr int FP_PAD = 256;                 // pad symbols in front of the haplotype (>= longest read - 32, multiple of 32)
constexpr int FP_MAX_READ = 256;            // 8 words per plane in registers
constexpr int FP_READ_WORDS = FP_MAX_READ / 32;
constexpr int FP_MAX_HAP = 2816;
constexpr int FP_BLOCK = 256;
constexpr int FP_MATCH = 4, FP_MISMATCH = 6, FP_MAX_MISMATCHES = 2;        // ReAligner::set_options, realigner.cpp:62-72

struct FpHap { long long text; int len; int region; int local; int is_ref; };
struct FpRegion { long long place_first; int hap_first, nhap, read_first, nread, prefix, suffix; };

__host__ __device__ inline int fp_hap_words(int hap_len) { return ((FP_PAD + hap_len) >> 5) + FP_READ_WORDS + 1; }
// dynamic shared memory: base planes [3][nhw + 1], shifted planes [3][nhw][32], cov[hap_len] (int), occupied bits
__host__ __device__ inline size_t fp_smem_bytes(int hap_len)
{
    const size_t nhw = (size_t)fp_hap_words(hap_len);
    return sizeof(uint32_t) * (3 * (nhw + 1) + 3 * nhw * 32 + (size_t)hap_len + (size_t)((hap_len + 31) / 32 + 2));
}

// symbol -> 3 bits: A 0, C 1, G 2, T 3, N 4; 7 = padding; 8 = any other character (flags the region)
__device__ __forceinline__ int fp_symbol(char c)
{
    switch (c) { case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3; case 'N': return 4; default: return 8; }
}

__global__ void __launch_bounds__(FP_BLOCK)
fastpass_kernel(const char* __restrict__ text, const FpHap* __restrict__ haps, const FpRegion* __restrict__ regions,
                const long long* __restrict__ read_start, const int* __restrict__ read_len,
                int2* __restrict__ places, int* __restrict__ hap_score, int* __restrict__ region_flag)
{
    extern __shared__ uint32_t fp_smem[];
    __shared__ int s_acc, s_drop;
    const FpHap hp = haps[blockIdx.x];
    const FpRegion rg = regions[hp.region];
    const int L = hp.len;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = FP_BLOCK / 32;
    int2* const my_places = places + rg.place_first + (long long)hp.local * rg.nread;

    if (L < FP_K) {                          // realigner.cpp:176 loops over an unsigned bound that wraps; nothing can be placed
        for (int r = tid; r < rg.nread; r += FP_BLOCK) my_places[r] = make_int2(0, -1);
        if (tid == 0) hap_score[rg.hap_first + hp.local] = 0;
        return;
    }
    const int nhw = fp_hap_words(L);
    uint32_t* const base = fp_smem;                              // [3][nhw + 1]
    uint32_t* const shifted = base + 3 * (nhw + 1);              // [3][nhw][32]
    int* const cov = reinterpret_cast<int*>(shifted + 3 * nhw * 32);   // [L] scan time at which the base was first covered
    uint32_t* const occ = reinterpret_cast<uint32_t*>(cov + L);  // bit i: some read shares the 32-mer at hap position i

    // ---- haplotype planes (padded coordinates: bit FP_PAD + p is hap base p)
    bool other = false;
    for (int j = warp; j <= nhw; j += nwarps) {
        const int p = 32 * j + lane - FP_PAD;
        int sym = 7;
        if (p >= 0 && p < L) { sym = fp_symbol(text[hp.text + p]); if (sym == 8) { other = true; sym = 7; } }
        const uint32_t b0 = __ballot_sync(0xffffffffu, sym & 1), b1 = __ballot_sync(0xffffffffu, sym & 2), b2 = __ballot_sync(0xffffffffu, sym & 4);
        if (lane == 0) { base[j] = b0; base[(nhw + 1) + j] = b1; base[2 * (nhw + 1) + j] = b2; }
    }
    for (int p = tid; p < L; p += FP_BLOCK) cov[p] = INT_MAX;
    for (int p = tid; p < (L + 31) / 32 + 2; p += FP_BLOCK) occ[p] = 0u;
    if (tid == 0) { s_acc = 0; s_drop = 0; }
    __syncthreads();
    for (int idx = tid; idx < 3 * nhw * 32; idx += FP_BLOCK) {
        const int pl = idx / (nhw * 32), rem = idx - pl * nhw * 32, j = rem >> 5, sh = rem & 31;
        const uint32_t* bp = base + pl * (nhw + 1);
        shifted[idx] = __funnelshift_r(bp[j], bp[j + 1], sh);    // bit x = padded hap position 32 j + sh + x
    }
    __syncthreads();
    const uint32_t* const hs0 = shifted + lane, * const hs1 = shifted + nhw * 32 + lane, * const hs2 = shifted + 2 * nhw * 32 + lane;

    // ---- reads of the region: one warp each
    for (int r = warp; r < rg.nread; r += nwarps) {
        const int n = read_len[rg.read_first + r];
        if (n <= FP_K || n > FP_MAX_READ) {                      // AddReadToIndex, realigner.cpp:436-438: reads of at most k bases are not indexed
            if (lane == 0) my_places[r] = make_int2(0, -1);      // (longer than FP_MAX_READ: refused on the host before the launch)
            continue;
        }
        const int nw = (n + 31) >> 5;
        const long long rs = read_start[rg.read_first + r];
        uint32_t r0[FP_READ_WORDS], r1[FP_READ_WORDS], r2[FP_READ_WORDS];
#pragma unroll
        for (int w = 0; w < FP_READ_WORDS; ++w) {
            int sym = 0;
            if (w < nw) {
                const int x = 32 * w + lane;
                if (x < n) { sym = fp_symbol(text[rs + x]); if (sym == 8) { other = true; sym = 4; } }
            }
            r0[w] = __ballot_sync(0xffffffffu, sym & 1); r1[w] = __ballot_sync(0xffffffffu, sym & 2); r2[w] = __ballot_sync(0xffffffffu, sym & 4);
        }
        const uint32_t tail = (n & 31) ? ((1u << (n & 31)) - 1u) : 0xffffffffu;     // valid bits of the last word

        const int d_lo = FP_PAD - (n - FP_K), d_hi = FP_PAD + L - FP_K;             // padded diagonals that can hold a run
        unsigned long long best = 0ull;                          // (score, -t, -o) of this lane's best passing start
        uint32_t zero_trig = 0xffffffffu;                        // earliest (t << 12 | o) at which start 0 is triggered (diagonals <= 0)
        int zero_mism = -1;                                      // mismatches of the start-0 comparison (lane that owns diagonal 0)
        for (int q = d_lo >> 5; q <= (d_hi >> 5); ++q) {
            const int D = 32 * q + lane, d = D - FP_PAD;
            int mism = 0, first_o = -1;
            // Filter: a 32-base run contains a whole aligned 16-base block, so on a diagonal with a run some half word of
            // (plane 0 differs | plane 1 differs) is zero.  Two loads and five integer instructions per word decide that for almost every
            // diagonal; one VIMNMX3 per two words tracks the smallest half word; only the few diagonals that pass (and diagonal 0, whose mismatch count the start-0 rule needs) take the exact path below.
            uint32_t lowest = 0xffffffffu;                       // per half word: minimum over the words of the diagonal (VIMNMX3.U16x2)
#pragma unroll
            for (int w = 0; w < FP_READ_WORDS; w += 2) {
                if (w < nw) {                                    // warp-uniform
                    const uint32_t xa = (r0[w] ^ hs0[(q + w) * 32]) | (r1[w] ^ hs1[(q + w) * 32]);
                    uint32_t xb = 0xffffffffu;
                    if (w + 1 < nw) xb = (r0[w + 1] ^ hs0[(q + w + 1) * 32]) | (r1[w + 1] ^ hs1[(q + w + 1) * 32]);
                    lowest = __vimin3_u16x2(lowest, xa, xb);
                }
            }
            const bool maybe = (lowest & 0xffffu) == 0u || (lowest >> 16) == 0u;
            if (maybe || d == 0) {
                uint32_t eprev = 0u;
#pragma unroll
                for (int w = 0; w <= FP_READ_WORDS; ++w) {
                    if (w <= nw) {                               // warp-uniform
                        uint32_t e = 0u;
                        if (w < nw) {
                            const uint32_t h0 = hs0[(q + w) * 32], h1 = hs1[(q + w) * 32], h2 = hs2[(q + w) * 32];
                            const uint32_t x0 = r0[w < FP_READ_WORDS ? w : 0] ^ h0, x1 = r1[w < FP_READ_WORDS ? w : 0] ^ h1, x2 = r2[w < FP_READ_WORDS ? w : 0] ^ h2;
                            e = ~(x0 | x1 | x2);
                            uint32_t m = (x0 | x1) & ~(r2[w < FP_READ_WORDS ? w : 0] | h2);
                            if (w == nw - 1) { e &= tail; m &= tail; }
                            mism += __popc(m);
                        }
                        if (w > 0 && eprev != 0u) {
                            // 32-base windows that start at bit b of word w-1: its top 32-b bits and the low b bits of word w are all equal
                            const int hi = __clz((int)~eprev), lo = (e == 0xffffffffu) ? 32 : (__ffs((int)~e) - 1);
                            const int b_min = 32 - hi, b_max = min(31, lo);
                            if (hi > 0 && b_min <= b_max) {
                                const int o_a = 32 * (w - 1) + b_min, o_b = 32 * (w - 1) + b_max;
                                if (first_o < 0) first_o = o_a;
                                const int i_a = d + o_a, cnt = o_b - o_a + 1;                // hap positions i_a .. i_a + cnt - 1 share a 32-mer with this read
                                const unsigned long long bits = ((1ull << cnt) - 1ull) << (i_a & 31);
                                atomicOr(&occ[i_a >> 5], (uint32_t)bits);
                                if ((uint32_t)(bits >> 32)) atomicOr(&occ[(i_a >> 5) + 1], (uint32_t)(bits >> 32));
                            }
                        }
                        eprev = e;
                    }
                }
            }
            if (d == 0) zero_mism = mism;
            bool pass = false;
            int t = 0;
            if (first_o >= 0 && D >= d_lo && D <= d_hi) {
                t = d + first_o;
                if (d <= 0) zero_trig = min(zero_trig, ((uint32_t)t << 12) | (uint32_t)first_o);
                else if (d + n <= L && mism <= FP_MAX_MISMATCHES) {
                    pass = true;
                    const unsigned long long key = ((unsigned long long)((n - mism) * FP_MATCH - mism * FP_MISMATCH) << 40) |
                                                   ((unsigned long long)(0xfffffu - (uint32_t)t) << 20) | (unsigned long long)(0xfffffu - (uint32_t)first_o);
                    best = key > best ? key : best;
                }
            }
            // passing starts cover their bases from time t on
            for (unsigned todo = __ballot_sync(0xffffffffu, pass); todo; todo &= todo - 1) {
                const int src = __ffs((int)todo) - 1;
                const int s = __shfl_sync(0xffffffffu, d, src), tt = __shfl_sync(0xffffffffu, t, src);
                for (int p = s + lane; p < s + n; p += 32) atomicMin(&cov[p], tt);
            }
        }
        // start 0: triggered by every run on a diagonal <= 0, evaluated once, at the earliest of those
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
            zero_trig = min(zero_trig, __shfl_xor_sync(0xffffffffu, zero_trig, off));
            zero_mism = max(zero_mism, __shfl_xor_sync(0xffffffffu, zero_mism, off));
        }
        if (zero_trig != 0xffffffffu && n <= L && zero_mism >= 0 && zero_mism <= FP_MAX_MISMATCHES) {
            const int t = (int)(zero_trig >> 12), o = (int)(zero_trig & 0xfffu);
            for (int p = lane; p < n; p += 32) atomicMin(&cov[p], t);
            const unsigned long long key = ((unsigned long long)((n - zero_mism) * FP_MATCH - zero_mism * FP_MISMATCH) << 40) |
                                           ((unsigned long long)(0xfffffu - (uint32_t)t) << 20) | (unsigned long long)(0xfffffu - (uint32_t)o);
            best = key > best ? key : best;
        }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
            const unsigned long long o = __shfl_xor_sync(0xffffffffu, best, off);
            best = o > best ? o : best;
        }
        if (lane == 0) {
            if (best == 0ull) my_places[r] = make_int2(0, -1);
            else {
                const int score = (int)(best >> 40);
                const int t = (int)(0xfffffu - (uint32_t)((best >> 20) & 0xfffffu)), o = (int)(0xfffffu - (uint32_t)(best & 0xfffffu));
                my_places[r] = make_int2(score, max(0, t - o));
                atomicAdd(&s_acc, score);
            }
        }
    }
    __syncthreads();
    // ---- realigner.cpp:248-252: a non-reference haplotype with a window base that is uncovered when the scan reaches it is dropped
    if (!hp.is_ref) {
        const unsigned long long limit = (unsigned long long)L - (unsigned long long)(long long)rg.suffix;     // size_t arithmetic of the reference
        for (int i = tid; i <= L - FP_K; i += FP_BLOCK)
            if (i >= rg.prefix && (unsigned long long)i < limit && ((occ[i >> 5] >> (i & 31)) & 1u) && cov[i] > i) s_drop = 1;
    }
    __syncthreads();
    if (tid == 0) hap_score[rg.hap_first + hp.local] = s_drop ? 0 : s_acc;
    if (other) atomicOr(&region_flag[hp.region], 1);
}

}  // namespace mpn
