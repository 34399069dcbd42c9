// Narrow-band traceback DP with the band row held in registers (sm_100a): replaces banded_sw (ssw.c:532-616) for bands of up to
// 15 cells on reads of up to NARROW_MAX_ROWS rows (fewer when the batch holds only a few long reads) -- every pair of BASELINE configs[1].
//
//   sw_band_setup_kernel     one thread per pair: begin positions from the reverse pass (ssw.c:820-832), the CIGAR filters of
//                            ssw.c:833, the trivial "1M" case, and the first band |sub_ref - sub_read| + 1 (ssw.c:838); pairs are
//                            appended to the work queue of their band half-width (1..7) or handed to the warp kernel (wider, longer).
//   sw_band_rows_kernel<BW>  one lane per pair, ONE band half-width per instantiation: a warp takes 32 consecutive pairs of the queue and
//                            does a whole read row per loop iteration, the 2*BW+1 cells unrolled with H / E of the previous row in registers
//                            (updated in place, band coordinates of the reference), the target window in a 128-bit register pair and
//                            the 4-bit direction cells of the row packed into one 64-bit store.  A failed attempt (banded maximum
//                            below the score, ssw.c:614-616) re-queues the pair in the queue of the doubled band, which is launched later.
// Band-coordinate quirks (`edge` zeroing, ssw.c:580) and tie rules as in sw_trace_warp.cuh; direction words as sw_band_trace_kernel reads them.
#pragma once
#include <type_traits>
#include "sw_trace_narrow.cuh"

namespace mpn {

constexpr int ROWS_BLOCK = 128;
constexpr int ROWS_MAXBW = NARROW_BW;                 // 7
constexpr int ROWS_CLASSES = ROWS_MAXBW;              // one queue (and one kernel instantiation) per band half-width 1 .. 7
__host__ __device__ constexpr int rows_class_of(int bw) { return bw - 1; }

// One queued banded-DP attempt: everything a lane needs to run it, in two 16-byte words, so that taking a pair costs ONE dependent load
// after the queue claim (round 1 chased queue -> task -> forward record -> begin positions: four round trips per pair and lane).
struct __align__(16) BandItem {
    long long rf_off, rd_off;   // arena index of the first target / read base of the sub-rectangle (ssw.c:836-839)
    int32_t out_i, k;           // result slot, index in the sorted task list
    uint32_t dims;              // sub_ref | sub_read << 16   (both <= NARROW_MAX_ROWS + NARROW_BW)
    uint32_t score_bw;          // score1 | band half-width << 16
};
static_assert(sizeof(BandItem) == 32, "two uint4");

struct BandQueues {
    BandItem* items;         // [ROWS_CLASSES][capacity]
    int* count;              // [ROWS_CLASSES] filled entries per queue
    int* head;               // [ROWS_CLASSES] consumed entries per queue
    int capacity;
};

__device__ __forceinline__ void band_enqueue(const BandQueues& q, BandItem it, int bw)
{
    const int cls = rows_class_of(bw);
    it.score_bw = (it.score_bw & 0xffffu) | ((uint32_t)bw << 16);
    q.items[(size_t)cls * q.capacity + atomicAdd(q.count + cls, 1)] = it;
}

__global__ void __launch_bounds__(128)
sw_band_setup_kernel(const SwTask* __restrict__ order, int ntasks, const FwdResult* __restrict__ fr, const SwEnds* __restrict__ rev, TraceParams tp,
                     uint32_t* __restrict__ cig, unsigned long long cig_cap, unsigned long long* __restrict__ cig_used, FinalResult* __restrict__ out,
                     BandRec* __restrict__ recs, int* __restrict__ flag_list, int* __restrict__ nflag, BandQueues q)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = k < ntasks;                      // whole warps stay: the queue appends below are warp-aggregated (one atomic per warp and queue)
    const SwTask tk = order[live ? k : 0];
    const int i = tk.out;
    const FwdResult f = fr[i];
    FinalResult r;
    r.ref_begin1 = -1; r.read_begin1 = -1; r.cigar_len = 0; r.status = 0; r.cigar_off = 0;
    BandRec br; br.dir_off = 0; br.bw = 0; br.kind = 0;
    int dest = -1;                                     // -1 nothing, 0..3 band queue of that class, 4 warp-kernel list
    BandItem it;
    it.rf_off = 0; it.rd_off = 0; it.out_i = i; it.k = k; it.dims = 0; it.score_bw = 0;
    if (live && f.want_rev) {
        if (f.score1 > 0) {
            const SwEnds e = rev[i];
            r.ref_begin1 = f.ref_end1 - e.col;
            r.read_begin1 = f.read_end1 - e.row;
        } else {
            r.ref_begin1 = f.word_mode ? 0 : -1;       // empty / 1x1 reverse matrix (ssw.c:820-831)
            r.read_begin1 = 0;
        }
        const bool no_cigar = (7 & tp.flag) == 0 || ((2 & tp.flag) != 0 && f.score1 < tp.filters) ||
            ((4 & tp.flag) != 0 && (f.ref_end1 - r.ref_begin1 > tp.filterd || f.read_end1 - r.read_begin1 > tp.filterd));   // ssw.c:833
        if (!no_cigar) {
            const int sub_ref = f.ref_end1 - r.ref_begin1 + 1, sub_read = f.read_end1 - r.read_begin1 + 1;
            const int bw = abs(sub_ref - sub_read) + 1;
            if (f.score1 <= 0) {
                unsigned long long o = atomicAdd(cig_used, 1ull);          // "1M" (ssw.c:625,680-687)
                if (o + 1 > cig_cap) r.status = 6;
                else { cig[o] = 1u << 4; r.cigar_off = (int64_t)o; r.cigar_len = 1; }
            } else if (tp.n > 8 || bw > ROWS_MAXBW || sub_read > tp.lane_max_rows || sub_ref < 2 * bw + 2) {
                // (the last condition keeps `edge` of ssw.c:580 at its steady value 2 bw + 2 in every row past the band's head; a
                //  sub-rectangle that small is a handful of cells)
                // wide bands, and reads so long that one lane would serialise millions of cells: one warp per pair instead
                r.status = 7; br.bw = bw; dest = ROWS_CLASSES;
            } else {
                it.rf_off = tk.rf_base + r.ref_begin1; it.rd_off = tk.rd_base + r.read_begin1;
                it.dims = (uint32_t)sub_ref | ((uint32_t)sub_read << 16);
                it.score_bw = (uint32_t)f.score1 | ((uint32_t)bw << 16);
                dest = rows_class_of(bw);
            }
        }
    }
    // one atomic per warp and destination instead of one per pair (a million same-address atomics cost as much as the DP of a band class)
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 0; d <= ROWS_CLASSES; ++d) {
        const unsigned m = __ballot_sync(0xffffffffu, dest == d);
        if (m == 0u) continue;
        int base = 0;
        if (lane == __ffs((int)m) - 1) base = atomicAdd(d < ROWS_CLASSES ? q.count + d : nflag, __popc(m));
        base = __shfl_sync(0xffffffffu, base, __ffs((int)m) - 1);
        if (dest == d) {
            const int at = base + __popc(m & ((1u << lane) - 1u));
            if (d < ROWS_CLASSES) q.items[(size_t)d * q.capacity + at] = it;
            else flag_list[at] = k;
        }
    }
    if (live) { out[i] = r; recs[i] = br; }
}

// eight consecutive sequence bytes starting at p[from], zero beyond `len`.  Two aligned 64-bit loads + a funnel shift instead of eight
// byte loads: the arena starts 256-byte aligned and is padded by 16 bytes (engine.cu), so the aligned words around any base are readable.
__device__ __forceinline__ unsigned long long load8(const uint8_t* __restrict__ p, int from, int len)
{
    const int left = len - from;
    if (left <= 0) return 0ull;
    const uintptr_t a = reinterpret_cast<uintptr_t>(p + from);
    const unsigned long long* w = reinterpret_cast<const unsigned long long*>(a & ~(uintptr_t)7);
    const unsigned sh = (unsigned)(a & 7u) * 8u;
    const unsigned long long lo = w[0], hi = w[1];
    unsigned long long v = sh ? ((lo >> sh) | (hi << (64u - sh))) : lo;
    if (left < 8) v &= (1ull << (8 * left)) - 1ull;
    return v;
}

// One kernel instantiation per band half-width BW (1..7): W = 2 BW + 1 band cells per read row, every one of them a real cell in the
// "steady" rows (BW < ii <= sub_ref - 1 - BW: the band neither touches the left edge of the matrix nor is clipped by its right edge).
// A warp takes 32 CONSECUTIVE queue items (the queue is filled in task order, i.e. by read-length bin: the 32 pairs have about the same
// number of rows) and walks them row-synchronously: the first BW + 1 rows with the general row code (band start at the matrix edge, row 0),
// every later row with the band row code, in which the band shift and the zeroed slot of ssw.c:580 are constants and only the clipping at the
// right edge of the matrix is left as a per-cell test (a lane whose pair is finished idles with every cell clipped).  Round 1 ran the
// general code on every row (~100 instructions per cell); the band row code is ~32.
template <int BW>
__device__ __forceinline__ void band_rows_body(const unsigned long long* __restrict__ srow, const int8_t* __restrict__ seq, const TraceParams& tp, const Arena& scratch,
                                               FinalResult* __restrict__ out, BandRec* __restrict__ recs, int* __restrict__ flag_list, int* __restrict__ nflag, const BandQueues& q)
{
    constexpr int CLS = BW - 1;
    constexpr int W = 2 * BW + 1;
    const int gapO = tp.gapO, gapE = tp.gapE;
    const int total = q.count[CLS];                    // complete: every producer of this queue ran in an earlier launch
    const uint4* items = reinterpret_cast<const uint4*>(q.items + (size_t)CLS * q.capacity);
    const uint8_t* useq = reinterpret_cast<const uint8_t*>(seq);
    const int lane = threadIdx.x & 31;

    for (;;) {
        int base = 0;
        if (lane == 0) base = atomicAdd(q.head + CLS, 32);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= total) break;
        const bool live = base + lane < total;
        int kcur = 0, i = 0, sub_ref = 1, sub_read = 0, score = 0;
        const uint8_t* ref = useq; const uint8_t* read = useq;
        if (live) {
            const uint4 nd0 = items[2 * (size_t)(base + lane)], nd1 = items[2 * (size_t)(base + lane) + 1];
            ref = useq + (long long)(((unsigned long long)nd0.y << 32) | nd0.x);
            read = useq + (long long)(((unsigned long long)nd0.w << 32) | nd0.z);
            i = (int)nd1.x; kcur = (int)nd1.y;
            sub_ref = (int)(nd1.z & 0xffffu); sub_read = (int)(nd1.z >> 16); score = (int)(nd1.w & 0xffffu);
        }
        // direction words: 2, 4 or 8 bytes per read row (dir_row_bytes); one arena claim per warp
        unsigned long long dir_off = 0;
        {
            constexpr unsigned long long RB = (unsigned long long)dir_row_bytes(BW);       // bytes per row of direction bits
            const unsigned long long need = ((unsigned long long)sub_read * RB + 15ull) & ~15ull;
            unsigned long long incl = need;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) { const unsigned long long v = __shfl_up_sync(0xffffffffu, incl, off); if (lane >= off) incl += v; }
            unsigned long long wbase_off = 0;
            if (lane == 31) wbase_off = atomicAdd(scratch.used, incl);
            wbase_off = __shfl_sync(0xffffffffu, wbase_off, 31);
            dir_off = wbase_off + incl - need;
        }
        bool run = live;
        if (run && dir_off + (unsigned long long)sub_read * (unsigned long long)dir_row_bytes(BW) > scratch.bytes) { out[i].status = 5; run = false; }
        uint8_t* const dirbase = scratch.base + (run ? dir_off : 0ull);
        auto store_row = [&](int r, unsigned long long word) {
            if (dir_row_bytes(BW) == 2) reinterpret_cast<uint16_t*>(dirbase)[r] = (uint16_t)word;
            else if (dir_row_bytes(BW) == 4) reinterpret_cast<uint32_t*>(dirbase)[r] = (uint32_t)word;
            else reinterpret_cast<unsigned long long*>(dirbase)[r] = word;
        };
        const int rows = run ? sub_read : 0;
        int rows_max = rows;
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) rows_max = max(rows_max, __shfl_xor_sync(0xffffffffu, rows_max, off));

        int H[W + 3], E[W + 3];                        // previous row in band coordinates: index 0 = left boundary, cell p lives at p + 1
#pragma unroll
        for (int x = 0; x < W + 3; ++x) { H[x] = 0; E[x] = 0; }
        unsigned long long win_lo = load8(ref, 0, sub_ref), win_hi = load8(ref, 8, sub_ref);            // target bases wbase .. wbase + 15
        unsigned long long tq_cur = load8(ref, 16, sub_ref), tq_nxt = load8(ref, 24, sub_ref);          // the 8 + 8 target bases that enter the window next
        unsigned long long rq_cur = load8(read, 0, sub_read), rq_nxt = load8(read, 8, sub_read);        // read bases of rows ii .. (octet end) / next octet
        int wbase = 0, maxv = 0;

        // ---- the general row (any ii): band start / validity / zeroed slot / row 0 decided per cell
        auto general_row = [&](const int ii) {
            const int xi = max(ii - BW, 0);
            const int end = min(sub_ref - 1, ii + BW);
            const int edge = min(end + 1, 2 * BW + 2);     // `width - 1` of ssw.c:557,580
            const bool sh = ii > BW;                       // band start moved by one column against the previous row
            if (xi > wbase) {                              // slide the target window by one base
                win_lo = (win_lo >> 8) | (win_hi << 56);
                win_hi = (win_hi >> 8) | (tq_cur << 56);
                tq_cur >>= 8;
                wbase = xi;
                if ((wbase & 7) == 0) { tq_cur = tq_nxt; tq_nxt = load8(ref, wbase + 24, sub_ref); }      // 8 slides consumed one word
            }
            const unsigned long long rscore = srow[(unsigned)rq_cur & 7u];
            rq_cur >>= 8;
            if ((ii & 7) == 7) { rq_cur = rq_nxt; rq_nxt = load8(read, ii + 9, sub_read); }
            const bool row0 = ii == 0;
            unsigned long long dirword = 0;
            int hleft = 0, fv = 0;
            int hdg = sh ? H[1] : 0;                       // H(ii-1, xi-1); band coordinate 0 is the matrix edge
#pragma unroll
            for (int p = 0; p < W; ++p) {
                const bool valid = xi + p <= end;
                const int e_idx = p + 1 + (sh ? 1 : 0);    // previous-row band coordinate of (ii-1, xi+p)
                int hup = sh ? H[p + 2] : H[p + 1];
                int eup = sh ? E[p + 2] : E[p + 1];
                if (e_idx == edge) { hup = 0; eup = 0; }   // the slot the reference zeroes before every row (ssw.c:580)
                const unsigned code = (unsigned)((p < 8 ? win_lo >> (8 * p) : win_hi >> (8 * (p - 8))) & 7ull);
                const int sc = (int)(int8_t)(rscore >> (8 * code));
                int open = row0 ? -gapO : hup - gapO;
                int ext = row0 ? -gapE : eup - gapE;
                const int ev = open > ext ? open : ext;
                const unsigned de3 = open > ext ? 1u : 0u;                                        // ties extend (ssw.c:593-594)
                open = hleft - gapO; ext = fv - gapE;
                fv = open > ext ? open : ext;
                const unsigned df5 = open > ext ? 1u : 0u;                                        // ssw.c:596-599
                const int e1 = ev > 0 ? ev : 0, f1 = fv > 0 ? fv : 0;
                const int t1 = e1 > f1 ? e1 : f1;
                const int t2 = hdg + sc;
                const int hv = t1 > t2 ? t1 : t2;
                const unsigned src = t1 <= t2 ? 1u : (e1 > f1 ? 2u : 3u);                         // 1 diagonal, 2 from E, 3 from F (ssw.c:609-610)
                hdg = hup;                                 // the next cell's diagonal (zeroed together with the slot)
                H[p + 1] = valid ? hv : 0;
                E[p + 1] = valid ? ev : 0;
                if (valid) {
                    maxv = max(maxv, hv);
                    dirword |= (unsigned long long)(de3 | (df5 << 1) | (src << 2)) << (4 * p);
                }
                hleft = hv;
            }
            store_row(ii, dirword);
        };
        // ---- a row past the band's head (ii > BW): the band start moves by one per row, so the previous row's cell above band cell p sits at
        //      p + 2; slot 2 BW + 2 (above the last cell: outside the previous row's band) is never written and reads 0 -- which is also what
        //      ssw.c:580 zeroes there (sub_ref >= 2 BW + 2, see the setup kernel).  Cells right of the matrix (p > lim) do not exist: they are
        //      stored as 0 and leave no direction bits.  A finished lane runs with lim = -1.
        auto band_row = [&](const int ii, auto clip_tag) {
            constexpr bool CLIP = decltype(clip_tag)::value;       // false: every lane of the warp has all W cells inside its matrix (no per-cell test)
            win_lo = (win_lo >> 8) | (win_hi << 56);
            win_hi = (win_hi >> 8) | (tq_cur << 56);
            tq_cur >>= 8;
            ++wbase;
            if ((wbase & 7) == 0) { tq_cur = tq_nxt; tq_nxt = load8(ref, wbase + 24, sub_ref); }
            const unsigned long long rscore = srow[(unsigned)rq_cur & 7u];
            const uint32_t rs_lo = (uint32_t)rscore, rs_hi = (uint32_t)(rscore >> 32);
            rq_cur >>= 8;
            if ((ii & 7) == 7) { rq_cur = rq_nxt; rq_nxt = load8(read, ii + 9, sub_read); }
            const int lim = ii < rows ? sub_ref - 1 - (ii - BW) : -1;                          // last existing band cell of this row
            uint32_t dlo = 0, dhi = 0;
            int hleft = 0, fv = 0, hdg = H[1];
#pragma unroll
            for (int p = 0; p < W; ++p) {
                const bool valid = !CLIP || p <= lim;
                const int hup = H[p + 2], eup = E[p + 2];
                const uint32_t code = (uint32_t)((p < 8 ? win_lo >> (8 * p) : win_hi >> (8 * (p - 8)))) & 7u;
                const int sc = (int)prmt(rs_lo, rs_hi, code * 0x1111u + 0x8880u);              // byte `code` of the 8 scores, sign-extended
                int open = hup - gapO, ext = eup - gapE;
                const int ev = max(open, ext);
                const uint32_t de3 = open > ext ? 1u : 0u;
                open = hleft - gapO; ext = fv - gapE;
                fv = max(open, ext);
                const uint32_t df5 = open > ext ? 2u : 0u;
                const int f1 = max(fv, 0);
                const int t1 = max(max(ev, f1), 0);                                             // max(e1, f1)
                const int t2 = hdg + sc;
                const int hv = max(t1, t2);
                const uint32_t src = t1 <= t2 ? 4u : (ev > f1 ? 8u : 12u);                      // e1 > f1  <=>  ev > f1  (f1 >= 0)
                hdg = hup;
                H[p + 1] = valid ? hv : 0; E[p + 1] = valid ? ev : 0;
                maxv = max(maxv, H[p + 1]);
                const uint32_t bits = valid ? (de3 | df5 | src) : 0u;
                if (p < 8) dlo |= bits << (4 * p); else dhi |= bits << (4 * (p - 8));
                hleft = hv;
            }
            if (!CLIP || ii < rows) store_row(ii, ((unsigned long long)dhi << 32) | dlo);
        };

        const int head = min(BW + 1, rows_max);
        for (int ii = 0; ii < head; ++ii) if (ii < rows) general_row(ii);
        for (int ii = head; ii < rows_max; ++ii) {
            // most rows: every lane still has rows left and its band lies inside the matrix -> the row code without the clipping tests
            const bool inside = ii < rows && sub_ref - 1 - (ii - BW) >= W - 1;
            if (__all_sync(0xffffffffu, inside)) band_row(ii, std::false_type{});
            else band_row(ii, std::true_type{});
        }

        if (run) {
            if (maxv >= score) {                                                              // ssw.c:614-615
                BandRec br; br.dir_off = dir_off; br.bw = BW; br.kind = 2; recs[i] = br;
            } else if (2 * BW <= ROWS_MAXBW && sub_ref >= 4 * BW + 2) {
                BandItem it;
                it.rf_off = (long long)(ref - useq); it.rd_off = (long long)(read - useq); it.out_i = i; it.k = kcur;
                it.dims = (uint32_t)sub_ref | ((uint32_t)sub_read << 16); it.score_bw = (uint32_t)score;
                band_enqueue(q, it, 2 * BW);               // always a later queue: the launches go through the widths in increasing order
            } else {
                out[i].status = 7; BandRec br; br.dir_off = 0; br.bw = 2 * BW; br.kind = 0; recs[i] = br;
                flag_list[atomicAdd(nflag, 1)] = kcur;
            }
        }
    }
}

// Up to four band half-widths in one launch (blockIdx.y picks the instantiation; 0 = none): widths whose queues do not feed each other run
// side by side -- {1, 3, 5, 7}, then {2, 6} (fed by 1 and 3), then {4} (fed by 2) -- instead of seven launches in a row, most of which hold a
// few thousand pairs and cannot fill the GPU on their own.
template <int B0, int B1, int B2, int B3>
__global__ void __launch_bounds__(ROWS_BLOCK)
sw_band_rows_kernel(const int8_t* __restrict__ seq, TraceParams tp, Arena scratch,
                    FinalResult* __restrict__ out, BandRec* __restrict__ recs, int* __restrict__ flag_list, int* __restrict__ nflag, BandQueues q)
{
    __shared__ unsigned long long srow[8];             // srow[q] = bytes mat[t*n+q], t = 0..7: scores of read code q against the target codes
    const int n = tp.n;
    if (threadIdx.x < 8) {
        unsigned long long v = 0;
        if ((int)threadIdx.x < n && n <= 8)
            for (int t = 0; t < n; ++t) v |= (unsigned long long)(uint8_t)tp.mat[t * n + threadIdx.x] << (8 * t);
        srow[threadIdx.x] = v;
    }
    __syncthreads();
    const int y = blockIdx.y;
    if (y == 0) { if (B0 > 0) band_rows_body<(B0 > 0 ? B0 : 1)>(srow, seq, tp, scratch, out, recs, flag_list, nflag, q); }
    else if (y == 1) { if (B1 > 0) band_rows_body<(B1 > 0 ? B1 : 1)>(srow, seq, tp, scratch, out, recs, flag_list, nflag, q); }
    else if (y == 2) { if (B2 > 0) band_rows_body<(B2 > 0 ? B2 : 1)>(srow, seq, tp, scratch, out, recs, flag_list, nflag, q); }
    else { if (B3 > 0) band_rows_body<(B3 > 0 ? B3 : 1)>(srow, seq, tp, scratch, out, recs, flag_list, nflag, q); }
}

}  // namespace mpn
