// Internals of the batched engine that the multi-device pool (pool.cu) shares with engine.cu.  Not part of the C ABI.
#pragma once
#include "../../include/mpn_ssw_batch.h"
#include <cuda_runtime.h>
#include <cstdint>
#include <functional>

namespace mpn {

// Where the sequences of a batch come from.  Both forms end up as one device arena plus (start, length) spans per pair.
//   CsrPairs   reads / refs as two CSR arrays (mpn_batch_upload): arena = [reads | refs], copied straight from the caller's buffers
//   SpanPairs  one caller arena + explicit spans (mpn_batch_upload_spans): pairs may share sequences (a haplotype aligned to many reads)
struct CsrPairs {
    const int8_t* reads; const int64_t* read_off; const int8_t* refs; const int64_t* ref_off; int64_t npairs;
    int64_t reads_total() const { return npairs ? read_off[npairs] - read_off[0] : 0; }
    int64_t refs_total() const { return npairs ? ref_off[npairs] - ref_off[0] : 0; }
    int64_t rl(int64_t i) const { return read_off[i + 1] - read_off[i]; }
    int64_t fl(int64_t i) const { return ref_off[i + 1] - ref_off[i]; }
    int64_t rd_base(int64_t i) const { return read_off[i] - read_off[0]; }
    int64_t rf_base(int64_t i) const { return reads_total() + (ref_off[i] - ref_off[0]); }
    bool valid() const { return npairs == 0 || (reads && read_off && refs && ref_off); }
    bool span_ok(int64_t) const { return true; }
    size_t arena_bytes() const { return (size_t)(reads_total() + refs_total()); }
    size_t staging_bytes() const { return 0; }
    size_t h2d_bytes() const { return arena_bytes(); }
    void copy_arena(int8_t* dst, uint8_t*, cudaStream_t st) const {
        if (!npairs) return;
        if (reads_total()) cudaMemcpyAsync(dst, reads + read_off[0], (size_t)reads_total(), cudaMemcpyHostToDevice, st);
        if (refs_total()) cudaMemcpyAsync(dst + reads_total(), refs + ref_off[0], (size_t)refs_total(), cudaMemcpyHostToDevice, st);
    }
    int64_t read_bases() const { return reads_total(); }
    int8_t* shared_arena() const { return nullptr; }
    cudaEvent_t shared_ready() const { return nullptr; }
    CsrPairs slice(int64_t first, int64_t count) const { return CsrPairs{reads, read_off + first, refs, ref_off + first, count}; }
};
// CSR pairs whose bases are NIBBLE-PACKED on the host (mpn_align_batch_packed4): base i of a stream is the low (i even) or high (i odd)
// nibble of byte i / 2; offsets stay in bases.  Half the host->device bytes; a small kernel expands the nibbles into the int8 arena.
void launch_unpack4(const uint8_t* src, int64_t first_base, int64_t nbases, int8_t* dst, cudaStream_t st);
struct Csr4Pairs {
    const uint8_t* reads4; const int64_t* read_off; const uint8_t* refs4; const int64_t* ref_off; int64_t npairs;
    int64_t reads_total() const { return npairs ? read_off[npairs] - read_off[0] : 0; }
    int64_t refs_total() const { return npairs ? ref_off[npairs] - ref_off[0] : 0; }
    int64_t rl(int64_t i) const { return read_off[i + 1] - read_off[i]; }
    int64_t fl(int64_t i) const { return ref_off[i + 1] - ref_off[i]; }
    int64_t rd_base(int64_t i) const { return read_off[i] - read_off[0]; }
    int64_t rf_base(int64_t i) const { return reads_total() + (ref_off[i] - ref_off[0]); }
    bool valid() const { return npairs == 0 || (reads4 && read_off && refs4 && ref_off); }
    bool span_ok(int64_t) const { return true; }
    size_t arena_bytes() const { return (size_t)(reads_total() + refs_total()); }
    // packed bytes covering bases [lo, hi)
    static int64_t pk_first(int64_t lo) { return lo >> 1; }
    static size_t pk_bytes(int64_t lo, int64_t hi) { return hi > lo ? (size_t)(((hi + 1) >> 1) - (lo >> 1)) : 0; }
    size_t staging_bytes() const { return npairs ? pk_bytes(read_off[0], read_off[npairs]) + 32 + pk_bytes(ref_off[0], ref_off[npairs]) + 32 : 0; }
    size_t h2d_bytes() const { return npairs ? pk_bytes(read_off[0], read_off[npairs]) + pk_bytes(ref_off[0], ref_off[npairs]) : 0; }
    void copy_arena(int8_t* dst, uint8_t* staging, cudaStream_t st) const {
        if (!npairs) return;
        const size_t rb = pk_bytes(read_off[0], read_off[npairs]), fb = pk_bytes(ref_off[0], ref_off[npairs]);
        uint8_t* sr = staging; uint8_t* sf = staging + ((rb + 31) & ~(size_t)15);
        if (rb) { cudaMemcpyAsync(sr, reads4 + pk_first(read_off[0]), rb, cudaMemcpyHostToDevice, st); launch_unpack4(sr, read_off[0] & 1, reads_total(), dst, st); }
        if (fb) { cudaMemcpyAsync(sf, refs4 + pk_first(ref_off[0]), fb, cudaMemcpyHostToDevice, st); launch_unpack4(sf, ref_off[0] & 1, refs_total(), dst + reads_total(), st); }
    }
    int64_t read_bases() const { return reads_total(); }
    int8_t* shared_arena() const { return nullptr; }
    cudaEvent_t shared_ready() const { return nullptr; }
    Csr4Pairs slice(int64_t first, int64_t count) const { return Csr4Pairs{reads4, read_off + first, refs4, ref_off + first, count}; }
};
// CSR pairs whose bases are packed FOUR PER BYTE (mpn_align_batch_packed2): base i of a stream is bits 2 (i & 3) .. 2 (i & 3) + 1 of byte i / 4.
// Codes above 3 (N) do not fit: they are stored as 0 and listed, sorted by position, as (position << 4 | code) in an exception array per
// stream; a second small kernel patches them into the int8 arena after the expansion.  A quarter of the host->device bytes.
void launch_unpack2(const uint8_t* src, int64_t first_base, int64_t nbases, int8_t* dst, cudaStream_t st);
void launch_patch_exceptions(const int64_t* exc_dev, int64_t n, int64_t base_pos, int8_t* dst, cudaStream_t st);
struct Csr2Pairs {
    const uint8_t* reads2; const int64_t* read_off; const uint8_t* refs2; const int64_t* ref_off; int64_t npairs;
    const int64_t* read_exc; int64_t n_read_exc; const int64_t* ref_exc; int64_t n_ref_exc;
    int64_t reads_total() const { return npairs ? read_off[npairs] - read_off[0] : 0; }
    int64_t refs_total() const { return npairs ? ref_off[npairs] - ref_off[0] : 0; }
    int64_t rl(int64_t i) const { return read_off[i + 1] - read_off[i]; }
    int64_t fl(int64_t i) const { return ref_off[i + 1] - ref_off[i]; }
    int64_t rd_base(int64_t i) const { return read_off[i] - read_off[0]; }
    int64_t rf_base(int64_t i) const { return reads_total() + (ref_off[i] - ref_off[0]); }
    bool valid() const { return npairs == 0 || (reads2 && read_off && refs2 && ref_off && (n_read_exc == 0 || read_exc) && (n_ref_exc == 0 || ref_exc)); }
    bool span_ok(int64_t) const { return true; }
    size_t arena_bytes() const { return (size_t)(reads_total() + refs_total()); }
    static int64_t pk_first(int64_t lo) { return lo >> 2; }
    static size_t pk_bytes(int64_t lo, int64_t hi) { return hi > lo ? (size_t)(((hi + 3) >> 2) - (lo >> 2)) : 0; }
    // exceptions with lo <= position < hi: [first, last) of a sorted list
    static void exc_range(const int64_t* exc, int64_t n, int64_t lo, int64_t hi, int64_t& first, int64_t& last) {
        int64_t a = 0, b = n;
        while (a < b) { const int64_t m = (a + b) >> 1; if ((exc[m] >> 4) < lo) a = m + 1; else b = m; }
        first = a; b = n;
        while (a < b) { const int64_t m = (a + b) >> 1; if ((exc[m] >> 4) < hi) a = m + 1; else b = m; }
        last = a;
    }
    size_t exc_bytes() const {
        if (!npairs) return 0;
        int64_t a, b, c, d;
        exc_range(read_exc, n_read_exc, read_off[0], read_off[npairs], a, b);
        exc_range(ref_exc, n_ref_exc, ref_off[0], ref_off[npairs], c, d);
        return (size_t)((b - a) + (d - c)) * sizeof(int64_t);
    }
    size_t staging_bytes() const { return npairs ? pk_bytes(read_off[0], read_off[npairs]) + 48 + pk_bytes(ref_off[0], ref_off[npairs]) + 48 + exc_bytes() + 16 : 0; }
    size_t h2d_bytes() const { return npairs ? pk_bytes(read_off[0], read_off[npairs]) + pk_bytes(ref_off[0], ref_off[npairs]) + exc_bytes() : 0; }
    void copy_arena(int8_t* dst, uint8_t* staging, cudaStream_t st) const {
        if (!npairs) return;
        const size_t rb = pk_bytes(read_off[0], read_off[npairs]), fb = pk_bytes(ref_off[0], ref_off[npairs]);
        uint8_t* sr = staging; uint8_t* sf = sr + ((rb + 47) & ~(size_t)15);
        int64_t* se = reinterpret_cast<int64_t*>(sf + ((fb + 47) & ~(size_t)15));
        if (rb) { cudaMemcpyAsync(sr, reads2 + pk_first(read_off[0]), rb, cudaMemcpyHostToDevice, st); launch_unpack2(sr, read_off[0] & 3, reads_total(), dst, st); }
        if (fb) { cudaMemcpyAsync(sf, refs2 + pk_first(ref_off[0]), fb, cudaMemcpyHostToDevice, st); launch_unpack2(sf, ref_off[0] & 3, refs_total(), dst + reads_total(), st); }
        int64_t a, b;
        exc_range(read_exc, n_read_exc, read_off[0], read_off[npairs], a, b);
        if (b > a) { cudaMemcpyAsync(se, read_exc + a, (size_t)(b - a) * sizeof(int64_t), cudaMemcpyHostToDevice, st); launch_patch_exceptions(se, b - a, read_off[0], dst, st); se += b - a; }
        exc_range(ref_exc, n_ref_exc, ref_off[0], ref_off[npairs], a, b);
        if (b > a) { cudaMemcpyAsync(se, ref_exc + a, (size_t)(b - a) * sizeof(int64_t), cudaMemcpyHostToDevice, st); launch_patch_exceptions(se, b - a, ref_off[0], dst + reads_total(), st); }
    }
    int64_t read_bases() const { return reads_total(); }
    int8_t* shared_arena() const { return nullptr; }
    cudaEvent_t shared_ready() const { return nullptr; }
    Csr2Pairs slice(int64_t first, int64_t count) const { return Csr2Pairs{reads2, read_off + first, refs2, ref_off + first, count, read_exc, n_read_exc, ref_exc, n_ref_exc}; }
};
struct SpanPairs {
    const int8_t* seq; int64_t seq_bytes; const int64_t* rd_start; const int32_t* rd_len; const int64_t* rf_start; const int32_t* rf_len; int64_t npairs;
    // ranges of one caller batch share ONE device copy of the arena (uploaded by the caller of run_ranges, `ready` recorded behind the copy)
    int8_t* dev_arena = nullptr; cudaEvent_t ready = nullptr;
    int8_t* shared_arena() const { return dev_arena; }
    cudaEvent_t shared_ready() const { return ready; }
    int64_t rl(int64_t i) const { return rd_len[i]; }
    int64_t fl(int64_t i) const { return rf_len[i]; }
    int64_t rd_base(int64_t i) const { return rd_start[i]; }
    int64_t rf_base(int64_t i) const { return rf_start[i]; }
    bool valid() const { return seq_bytes >= 0 && (npairs == 0 || (seq && rd_start && rd_len && rf_start && rf_len)); }
    bool span_ok(int64_t i) const { return rd_start[i] >= 0 && rf_start[i] >= 0 && rd_start[i] + rd_len[i] <= seq_bytes && rf_start[i] + rf_len[i] <= seq_bytes; }
    size_t arena_bytes() const { return (size_t)seq_bytes; }
    size_t staging_bytes() const { return 0; }
    size_t h2d_bytes() const { return dev_arena ? 0 : arena_bytes(); }
    void copy_arena(int8_t* dst, uint8_t*, cudaStream_t st) const { if (seq_bytes) cudaMemcpyAsync(dst, seq, (size_t)seq_bytes, cudaMemcpyHostToDevice, st); }
    int64_t read_bases() const { int64_t t = 0; for (int64_t i = 0; i < npairs; ++i) t += rd_len[i]; return t; }
    SpanPairs slice(int64_t first, int64_t count) const { return SpanPairs{seq, seq_bytes, rd_start + first, rd_len + first, rf_start + first, rf_len + first, count, dev_arena, ready}; }
};

// A contiguous run of pairs of a caller's batch, and where its CIGAR words go in the caller's arena.
struct RangeJob {
    int64_t first = 0, count = 0;
    int64_t cig_base = -1;     // < 0: append at the engine's running cursor (one engine, ranges in order); else the range owns [cig_base, cig_base + cig_cap)
    int64_t cig_cap = 0;
};

// One engine, many ranges: keeps up to four ranges in flight (own stream + own pinned staging each), so the host work of a range
// (scheduling, H2D enqueue, D2H + record conversion) hides behind the kernels of the others.  `next` hands out ranges until it
// returns false (it may be shared between engines: the pool's devices pull from one queue).  Records land at out[first ..).
template <class Pairs>
int run_ranges(mpn_engine* e, const mpn_params* p, const Pairs& all, const int32_t* masklen, const std::function<bool(RangeJob&)>& next,
               mpn_result* out, uint32_t* cigar, int64_t cigar_cap, int64_t* pairs_done, int64_t* cells_done);

// which score kernel a pair goes to and its relative cost per cell (strip16 = 1): the pool balances ranges on cells x cost
double pair_cost_per_cell(int n, int maxpos, int64_t rl, int64_t fl);
int engine_device(const mpn_engine* e);

}  // namespace mpn
