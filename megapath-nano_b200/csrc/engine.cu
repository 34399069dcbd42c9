// Host side of the batched Smith-Waterman engine: length-binned scheduler, device memory, kernel launches and the
// C ABI of include/mpn_ssw_batch.h.  No CPU alignment code lives here: every DP cell is computed by the kernels in
// sw_strip16.cuh / sw_wide32.cuh / sw_finish.cuh / sw_trace.cuh, and a missing or failing GPU aborts loudly.
#include "../../include/mpn_ssw_batch.h"
#include "sw_common.cuh"
#include "strip_table.h"
#include "sw_wide32.cuh"
#include "sw_long16.cuh"
#include "sw_finish.cuh"
#include "sw_revband.cuh"
#include "sw_trace.cuh"
#include "sw_trace_narrow.cuh"
#include "sw_trace_warp.cuh"
#include "sw_trace_rows.cuh"
#include "fastpass.cuh"
#include "host_shared.h"
#include "engine_internal.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

using namespace mpn;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
    fprintf(stderr, "[mpn_ssw] CUDA error %s (%s) at %s:%d -- no CPU fallback, aborting\n", cudaGetErrorName(e_), cudaGetErrorString(e_), __FILE__, __LINE__); \
    abort(); } } while (0)

namespace {

struct DevBuf {
    void* p = nullptr; size_t cap = 0;
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

// Device allocations are recycled through the engine: a batch takes the smallest cached block that fits and gives it
// back in mpn_batch_free, so steady-state batches of similar size never call cudaMalloc / cudaFree.
struct DevPool {
    std::vector<DevBuf> free_list;
    void take(DevBuf& d, size_t bytes) {
        if (bytes <= d.cap) return;
        give(d);
        int best = -1;
        for (size_t i = 0; i < free_list.size(); ++i)
            if (free_list[i].cap >= bytes && (best < 0 || free_list[i].cap < free_list[best].cap)) best = (int)i;
        if (best >= 0 && free_list[best].cap <= bytes * 4 + (1u << 20)) { d = free_list[best]; free_list.erase(free_list.begin() + best); return; }
        // big blocks are rounded up to a power of two so that batches of varying size (one realigner region after the other) keep
        // hitting the cache instead of alternating cudaMalloc / cudaFree
        size_t want = bytes + bytes / 16 + 256;
        if (want > (1u << 20)) { size_t p2 = 1u << 20; while (p2 < want) p2 <<= 1; want = p2; }
        cudaError_t err = cudaMalloc(&d.p, want);
        if (err != cudaSuccess) {            // out of memory: drop the cache and retry once
            cudaGetLastError();
            clear();
            CK(cudaMalloc(&d.p, want));
        }
        d.cap = want;
    }
    void give(DevBuf& d) { if (d.p) free_list.push_back(d); d.p = nullptr; d.cap = 0; }
    void clear() { for (DevBuf& f : free_list) cudaFree(f.p); free_list.clear(); }
};

// grow-only pinned host staging
struct PinBuf {
    void* p = nullptr; size_t cap = 0;
    void reserve(size_t bytes) {
        if (bytes <= cap) return;
        if (p) CK(cudaFreeHost(p));
        cap = bytes + bytes / 8 + 4096;
        CK(cudaHostAlloc(&p, cap, cudaHostAllocDefault));
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

constexpr int STRIP_BLOCK_THREADS = 128;   // = STRIP_BLOCK of sw_strip16.cuh (the kernels are compiled in strip_inst_*.cu)
struct StripCfg { int G, KR, cap; StripFn fn; StripFn fn_rev; size_t smem, smem_rev; };

// all instantiations of the packed kernel, sorted by the number of read rows one strip covers (cap = 2 * G * KR)
std::vector<StripCfg> g_strips;
std::vector<int16_t> g_bin_of_len;       // read length -> index into g_strips (smallest strip that fits)
int N_STRIPS = 0;
int LONG_BIN = 0;                        // pseudo-bin of the multi-strip clamped 16-bit kernel (sw_long16.cuh)
int WIDE_BIN = 0;                        // pseudo-bin of the 32-bit kernel
constexpr int LONG_KR = 16;

void build_strip_table()
{
    if (!g_strips.empty()) return;
    const StripEntry* parts[] = {g_strip_part_a, g_strip_part_b, g_strip_part_c, g_strip_part_d};
    const int counts[] = {g_strip_part_a_n, g_strip_part_b_n, g_strip_part_c_n, g_strip_part_d_n};
    for (int p = 0; p < 4; ++p)
        for (int k = 0; k < counts[p]; ++k) {
            const StripEntry& e = parts[p][k];
            g_strips.push_back(StripCfg{e.G, e.KR, 2 * e.G * e.KR, e.fn, e.fn_rev, e.smem, e.smem_rev});
        }
    std::stable_sort(g_strips.begin(), g_strips.end(), [](const StripCfg& a, const StripCfg& b) { return a.cap < b.cap; });
    N_STRIPS = (int)g_strips.size();
    LONG_BIN = N_STRIPS;
    WIDE_BIN = N_STRIPS + 1;
    g_bin_of_len.assign((size_t)g_strips.back().cap + 1, 0);
    int k = 0;
    for (int len = 0; len <= g_strips.back().cap; ++len) {
        while (g_strips[k].cap < len) ++k;
        g_bin_of_len[len] = (int16_t)k;
    }
}

}  // namespace

struct mpn_engine {
    int device = 0;
    int sm_count = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    int64_t launches = 0, pairs = 0, cells = 0, wide_pairs = 0;
    // per strip instantiation (forward / reverse, then the 4 N variants, then the multi-strip one): resident blocks per SM, 0 = not prepared yet on this device.
    // Preparation (shared-memory limit above 48 KB, all-shared carve-out, occupancy query) happens at the first launch of an instantiation,
    // not at engine creation: touching all ~80 kernels up front loads every one of them and costs seconds in a process that aligns a
    // handful of pairs (the reference's process-per-position model, realignment.sh:50-60).
    std::vector<int> strip_blocks;
    bool profile = false;
    static constexpr int NAUX = 8;                   // side streams: the bins of one score pass run concurrently, so the tail of one launch overlaps the next
    int naux = 3;                                    // ... of which the score passes use this many (MPN_NAUX, A/B switch); the band classes use the first three
    cudaStream_t aux[NAUX] = {};
    cudaEvent_t ev_fork = nullptr, ev_join[NAUX] = {};
    cudaEvent_t ev_arena = nullptr;                  // behind the upload of an arena shared by the ranges of one spans batch
    static constexpr int NEVSET = 16;                // phase events of the last NEVSET runs (mpn_engine_phase_ms_mean averages them)
    cudaEvent_t evs[NEVSET][5] = {};
    cudaEvent_t* ev = evs[0];                        // the set of the run being enqueued
    int ev_valid = 0;
    int64_t ev_runs = 0;                             // completed profiled runs since profiling was switched on
    DevPool pool;
    struct Slot { cudaStream_t st = nullptr; PinBuf pin_tasks, pin_fwd, pin_fin, pin_misc; };
    static constexpr int NSLOT = 9;                  // slot 0 serves the phased API on the engine stream; 1..8 are the pipeline of mpn_align_batch (run_ranges uses the first MPN_PIPE_DEPTH, default 4)
    Slot slot[NSLOT];
    std::vector<int32_t> h_bin;
    std::vector<int64_t> h_order, h_idx, h_cnt;
    // k-mer fast pass (mpn_fastpass)
    DevBuf fp_text, fp_haps, fp_regions, fp_rstart, fp_rlen, fp_places, fp_score, fp_flag;
    cudaEvent_t fp_ev[2] = {nullptr, nullptr};
    float fp_kernel_ms = 0.f;
    bool fp_attr_done = false;
    int revband_blocks[5] = {0, 0, 0, 0, 0};           // resident blocks per SM of the banded reverse kernels, one per class (0 = not asked yet)
};

struct BinLaunch { int cfg; int64_t first; int64_t count; };

struct mpn_batch {
    mpn_engine* e = nullptr;
    mpn_params p{};
    std::vector<int8_t> mat;
    int64_t npairs = 0;
    int64_t total_cells = 0;
    int max_rd = 0, max_rf = 0;
    std::vector<BinLaunch> bins;
    int64_t n_wide_pre = 0;               // pairs routed to the 32-bit kernel by the host classifier
    Score16 sc16{};
    FinishParams fin{};
    // device
    DevBuf seq, mask, tasks_fwd, tasks_rev, ends_fwd, ends_rev, colrec, fwdres, finalres, counters, dmat, scratch, cig, wide_boundary, long_boundary, warp_dir, bandrec, flaglist, bandq_items, bandq_meta, relist, pack_stage, revq_items, revq_meta;
    bool seq_shared = false;              // `seq` is a device arena shared by the ranges of one caller batch (not ours to free)
    bool revband = false;                 // reverse passes of the packed bins go through the banded lane-per-pair kernel (sw_revband.cuh)
    rb::Score rbsc{};
    int64_t strip_tasks = 0;              // tasks [0, strip_tasks) belong to the packed short-read bins
    size_t seq_reads_bytes = 0, seq_bytes = 0;
    int64_t colrec_words = 0;
    unsigned long long scratch_bytes = 0, cig_cap = 0;
    bool ran = false;
    mpn_engine::Slot* slot = nullptr;
    cudaStream_t st = nullptr;            // stream every copy and kernel of this batch is enqueued on
    bool pipelined = false;               // owned by mpn_align_batch's chunk pipeline: no host synchronisation inside upload
    size_t h2d_bytes = 0, d2h_bytes = 0;
    long long wide_stride = 0; int wide_blocks = 0, long_blocks = 0;
    unsigned long long warp_dir_stride = 0; int warp_trace_blocks = 1;
    int64_t sum_rd = 0, sum_rf = 0;       // over pairs (worst-case CIGAR words = sum_rd + sum_rf)
    int64_t n_long_rows = 0;              // pairs whose read has more than 512 rows
    int arena_retries = 0;
};

extern "C" mpn_engine* mpn_engine_create(int device)
{
    int ndev = 0;
    cudaError_t err = cudaGetDeviceCount(&ndev);
    if (err != cudaSuccess || ndev == 0) {
        fprintf(stderr, "[mpn_ssw] no CUDA device available (%s); this library has no CPU fallback\n", cudaGetErrorString(err));
        return nullptr;
    }
    mpn_engine* e = new mpn_engine();
    if (device < 0) CK(cudaGetDevice(&device));
    e->device = device;
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    e->sm_count = prop.multiProcessorCount;
    CK(cudaStreamCreateWithFlags(&e->own_stream, cudaStreamNonBlocking));
    e->stream = e->own_stream;
    for (int k = 1; k < mpn_engine::NSLOT; ++k) CK(cudaStreamCreateWithFlags(&e->slot[k].st, cudaStreamNonBlocking));
    for (int k = 0; k < mpn_engine::NAUX; ++k) {
        CK(cudaStreamCreateWithFlags(&e->aux[k], cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&e->ev_join[k], cudaEventDisableTiming));
    }
    CK(cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&e->ev_arena, cudaEventDisableTiming));
    { const char* v = getenv("MPN_NAUX"); const int n = v ? atoi(v) : 3; e->naux = n >= 1 && n <= mpn_engine::NAUX ? n : 3; }
    static std::once_flag once;
    std::call_once(once, build_strip_table);
    e->strip_blocks.assign((size_t)2 * (N_STRIPS + 5), 0);
    return e;
}

extern "C" void mpn_engine_destroy(mpn_engine* e)
{
    if (!e) return;
    cudaSetDevice(e->device);
    cudaStreamDestroy(e->own_stream);
    for (int k = 0; k < mpn_engine::NAUX; ++k) { cudaStreamDestroy(e->aux[k]); cudaEventDestroy(e->ev_join[k]); }
    cudaEventDestroy(e->ev_fork); cudaEventDestroy(e->ev_arena);
    for (DevBuf* d : {&e->fp_text, &e->fp_haps, &e->fp_regions, &e->fp_rstart, &e->fp_rlen, &e->fp_places, &e->fp_score, &e->fp_flag}) e->pool.give(*d);
    if (e->fp_attr_done) { cudaEventDestroy(e->fp_ev[0]); cudaEventDestroy(e->fp_ev[1]); }
    e->pool.clear();
    for (int k = 0; k < mpn_engine::NSLOT; ++k) {
        mpn_engine::Slot& sl = e->slot[k];
        sl.pin_tasks.release(); sl.pin_fwd.release(); sl.pin_fin.release(); sl.pin_misc.release();
        if (k > 0 && sl.st) cudaStreamDestroy(sl.st);
    }
    if (e->evs[0][0]) for (int k = 0; k < mpn_engine::NEVSET; ++k) for (int i = 0; i < 5; ++i) cudaEventDestroy(e->evs[k][i]);
    delete e;
}

extern "C" int mpn_engine_set_stream(mpn_engine* e, void* s)
{
    if (!e) return MPN_E_ARG;
    e->stream = s ? reinterpret_cast<cudaStream_t>(s) : e->own_stream;
    return 0;
}

extern "C" int mpn_engine_set_profile(mpn_engine* e, int on)
{
    if (!e) return MPN_E_ARG;
    CK(cudaSetDevice(e->device));
    if (on && !e->evs[0][0]) for (int k = 0; k < mpn_engine::NEVSET; ++k) for (int i = 0; i < 5; ++i) CK(cudaEventCreate(&e->evs[k][i]));
    e->profile = on != 0;
    e->ev_valid = 0; e->ev_runs = 0; e->ev = e->evs[0];
    return 0;
}

/* milliseconds of the last mpn_batch_run by phase: forward score kernels, score2/mode epilogue, reverse score kernels, traceback */
extern "C" int mpn_engine_phase_ms(mpn_engine* e, float* ms4)
{
    if (!e || !ms4 || !e->profile || e->ev_valid < 5) return MPN_E_ARG;
    CK(cudaEventSynchronize(e->ev[4]));
    for (int i = 0; i < 4; ++i) CK(cudaEventElapsedTime(&ms4[i], e->ev[i], e->ev[i + 1]));
    return 0;
}

/* the same, averaged over the profiled runs since mpn_engine_set_profile(e, 1) (at most the last 16); *nruns = how many */
extern "C" int mpn_engine_phase_ms_mean(mpn_engine* e, float* ms4, int* nruns)
{
    if (!e || !ms4 || !e->profile || e->ev_runs < 1) return MPN_E_ARG;
    const int n = (int)std::min<int64_t>(e->ev_runs, mpn_engine::NEVSET);
    double acc[4] = {0, 0, 0, 0};
    for (int k = 0; k < n; ++k) {
        cudaEvent_t* set = e->evs[(e->ev_runs - 1 - k) % mpn_engine::NEVSET];
        CK(cudaEventSynchronize(set[4]));
        for (int i = 0; i < 4; ++i) { float ms = 0; CK(cudaEventElapsedTime(&ms, set[i], set[i + 1])); acc[i] += ms; }
    }
    for (int i = 0; i < 4; ++i) ms4[i] = (float)(acc[i] / n);
    if (nruns) *nruns = n;
    return 0;
}

extern "C" int mpn_engine_stats(const mpn_engine* e, int64_t* launches, int64_t* pairs, int64_t* cells, int64_t* wide_pairs)
{
    if (!e) return MPN_E_ARG;
    if (launches) *launches = e->launches;
    if (pairs) *pairs = e->pairs;
    if (cells) *cells = e->cells;
    if (wide_pairs) *wide_pairs = e->wide_pairs;
    return 0;
}

extern "C" void mpn_batch_free(mpn_batch* b)
{
    if (!b) return;
    cudaSetDevice(b->e->device);
    // the blocks go back to a pool that batches on OTHER streams draw from: nothing of this batch may still be in flight (a batch freed
    // after run but before a successful fetch, or on an error path).  After a fetch the stream is already idle and this returns at once.
    if (b->st) cudaStreamSynchronize(b->st);
    DevBuf* bufs[] = {&b->seq, &b->mask, &b->tasks_fwd, &b->tasks_rev, &b->ends_fwd, &b->ends_rev,
                      &b->colrec, &b->fwdres, &b->finalres, &b->counters, &b->dmat, &b->scratch, &b->cig, &b->wide_boundary, &b->long_boundary, &b->warp_dir, &b->bandrec, &b->flaglist, &b->bandq_items, &b->bandq_meta, &b->relist, &b->pack_stage, &b->revq_items, &b->revq_meta};
    if (b->seq_shared) { b->seq.p = nullptr; b->seq.cap = 0; }          // the arena belongs to the caller of run_ranges
    for (DevBuf* d : bufs) b->e->pool.give(*d);
    delete b;
}

template <class Pairs>
static mpn_batch* upload_impl(mpn_engine* e, int slot_id, const mpn_params* p, const Pairs& src, const int32_t* masklen, int64_t npairs)
{
    if (!e || !p || !p->mat || p->n < 1 || p->n > 127 || npairs < 0 || npairs > 0x7fffffff) return nullptr;
    if (!src.valid() || (npairs > 0 && !masklen)) return nullptr;
    CK(cudaSetDevice(e->device));
    mpn_batch* b = new mpn_batch();
    b->e = e; b->p = *p; b->npairs = npairs;
    b->mat.assign(p->mat, p->mat + (size_t)p->n * p->n);
    b->p.mat = b->mat.data();
    b->slot = &e->slot[slot_id];
    b->st = slot_id == 0 ? e->stream : b->slot->st;
    b->pipelined = slot_id != 0;
    mpn_engine::Slot& sl = *b->slot;
    cudaStream_t st = b->st;
    const int n = p->n;

    // ---- scoring: matrix rows for the packed kernel, bias as ssw_init (ssw.c:741-745)
    int minv = 0, maxv = 0;
    for (int i = 0; i < n * n; ++i) { minv = std::min<int>(minv, p->mat[i]); maxv = std::max<int>(maxv, p->mat[i]); }
    memset(&b->sc16, 0, sizeof b->sc16);
    for (int t = 0; t < n && t < 8; ++t) {
        uint32_t w = 0;
        for (int q = 0; q < 4 && q < n; ++q) w |= (uint32_t)(uint8_t)p->mat[t * n + q] << (8 * q);
        b->sc16.matrow[t] = w;
    }
    if (n == 5) {       // N column constant over the target codes (both matrix builders of the reference: ssw_cpp.cpp:23-48, pyssw.py:61-79)
        bool same = true;
        for (int t = 1; t < 5; ++t) same = same && p->mat[t * 5 + 4] == p->mat[4];
        const uint32_t c = (uint32_t)(uint16_t)(int16_t)p->mat[4];
        b->sc16.ncol2 = c | (c << 16);
        b->sc16.ncol_ok = same ? 1u : 0u;
    }
    const uint32_t go = (uint32_t)(uint16_t)(int16_t)(-(p->gapO & 0xff)), ge = (uint32_t)(uint16_t)(int16_t)(-(p->gapE & 0xff));
    b->sc16.mgapO2 = go | (go << 16);
    b->sc16.mgapE2 = ge | (ge << 16);
    const bool have_byte = p->score_size == 0 || p->score_size == 2;
    b->fin.bias = have_byte ? (abs(minv) & 0xff) : 0;
    b->fin.have_byte = have_byte; b->fin.have_word = p->score_size == 1 || p->score_size == 2;
    b->fin.gapO = p->gapO & 0xff; b->fin.gapE = p->gapE & 0xff;
    b->fin.flag = p->flag & 0xff; b->fin.filters = p->filters & 0xffff;
    {   // banded reverse pass (sw_revband_core.h): 5-letter alphabet, one match score > 0 on the diagonal of ACGT, one mismatch score < 0
        // elsewhere among ACGT, N constant per column (so the N variants of the packed kernel can take what the band refuses), gapO >= gapE >= 1
        static const bool off = getenv("MPN_NO_REVBAND") != nullptr;         // A/B switch
        bool ok = !off && n == 5 && b->sc16.ncol_ok && b->fin.gapO >= b->fin.gapE && b->fin.gapE >= 1 && (b->fin.flag != 0);
        const int mt = ok ? p->mat[0] : 0, mm = ok ? p->mat[1] : 0;
        ok = ok && mt > 0 && mm < 0;
        bool n_mm = true;
        for (int r = 0; ok && r < 5; ++r)
            for (int c = 0; c < 5; ++c) {
                const int v = p->mat[r * 5 + c];
                if (r < 4 && c < 4) ok = ok && v == (r == c ? mt : mm);
                else n_mm = n_mm && v == mm;
            }
        b->revband = ok;
        if (ok) {
            const uint32_t m8 = (uint32_t)(uint8_t)(int8_t)mm;
            b->rbsc.tlo = (uint32_t)(uint8_t)(int8_t)mt | (m8 << 8) | (m8 << 16) | (m8 << 24);
            b->rbsc.thi = m8 * 0x01010101u;
            b->rbsc.mgo2 = b->sc16.mgapO2; b->rbsc.mge2 = b->sc16.mgapE2;
            b->rbsc.mt = mt; b->rbsc.gapO = b->fin.gapO; b->rbsc.gapE = b->fin.gapE; b->rbsc.n_is_mismatch = n_mm ? 1 : 0;
        }
    }

    // ---- per-pair lengths, binning
    std::vector<int32_t>& bin = e->h_bin;
    bin.resize(npairs);
    std::vector<int64_t> bin_count(N_STRIPS + 2, 0);
    int64_t cm_total = 0, cells = 0, n_long = 0;
    int max_rd = 0, max_rf = 0, min_rf = 0x7fffffff;
    const bool packed_ok = n <= 8;
    static const bool no_long16 = getenv("MPN_NO_LONG16") != nullptr;      // A/B switch: send long / saturating pairs to the 32-bit kernel
    const int64_t maxpos = std::max(maxv, 0);
    for (int64_t i = 0; i < npairs; ++i) {
        const int64_t rl = src.rl(i), fl = src.fl(i);
        if (rl < 0 || fl < 0 || rl > 0x3fffffff || fl > 0x3fffffff || !src.span_ok(i)) { delete b; return nullptr; }
        cells += rl * fl;
        n_long += rl > 512;
        max_rd = std::max<int>(max_rd, (int)rl); max_rf = std::max<int>(max_rf, (int)fl); min_rf = std::min<int>(min_rf, (int)fl);
        cm_total += fl;
        int c = packed_ok ? LONG_BIN : WIDE_BIN;
        // the short-read packed kernel is exact as long as no H can reach the int16 clamp of ssw.c:425 (and the add cannot wrap);
        // longer reads and pairs that can saturate go to the multi-strip kernel, which carries the clamp
        if (packed_ok && (std::min(rl, fl) + 1) * maxpos <= 32767 && rl < (int64_t)g_bin_of_len.size()) c = g_bin_of_len[rl];
        if (no_long16 && c == LONG_BIN) c = WIDE_BIN;
        bin[i] = c; bin_count[c]++;
    }
    b->total_cells = cells; b->max_rd = max_rd; b->max_rf = max_rf; b->colrec_words = cm_total;
    b->sum_rf = cm_total; b->sum_rd = src.read_bases(); b->n_long_rows = n_long;
    b->n_wide_pre = bin_count[WIDE_BIN] + bin_count[LONG_BIN];
    {   // a launch with too few tasks cannot fill the GPU and the launches of a batch run back to back: fold thin bins into the next
        // larger strip (a few more dead rows, far better occupancy).  The largest packed bin keeps whatever it has.
        const int64_t thin = (int64_t)e->sm_count * 64;
        std::vector<int> remap(N_STRIPS + 2);
        for (int c = 0; c <= N_STRIPS + 1; ++c) remap[c] = c;
        bool any = false;
        for (int c = 0; c + 1 < N_STRIPS; ++c) {
            if (bin_count[c] == 0 || bin_count[c] >= thin) continue;
            int d = c + 1;
            while (d + 1 < N_STRIPS && bin_count[d] == 0) ++d;
            if (bin_count[d] == 0) continue;                    // nothing above: keep the bin
            bin_count[d] += bin_count[c]; bin_count[c] = 0; remap[c] = d; any = true;
        }
        if (any) {
            for (int c = N_STRIPS - 1; c >= 0; --c) remap[c] = remap[c] == c ? c : remap[remap[c]];
            for (int64_t i = 0; i < npairs; ++i) bin[i] = remap[bin[i]];
        }
    }

    // ---- task lists in pinned staging: per bin, longest targets first (counting sort on target length) so that the groups of
    //      a warp run in step and the tail of a launch is made of short tasks
    std::vector<int64_t> bin_first(N_STRIPS + 3, 0);
    for (int c = 0; c <= N_STRIPS + 1; ++c) bin_first[c + 1] = bin_first[c] + bin_count[c];
    sl.pin_tasks.reserve(sizeof(SwTask) * (size_t)(npairs + 1));
    SwTask* h_tasks = sl.pin_tasks.as<SwTask>();
    {
        std::vector<int64_t>& order = e->h_order;
        order.resize(npairs);
        std::vector<int64_t> cursor(bin_first.begin(), bin_first.end() - 1);
        if (npairs == 0 || min_rf == max_rf || max_rf > (1 << 22)) {
            for (int64_t i = 0; i < npairs; ++i) order[cursor[bin[i]]++] = i;
        } else {
            std::vector<int64_t>& cnt = e->h_cnt; std::vector<int64_t>& idx = e->h_idx;
            cnt.assign((size_t)max_rf + 2, 0); idx.resize(npairs);
            for (int64_t i = 0; i < npairs; ++i) cnt[max_rf - src.fl(i) + 1]++;
            for (int v = 0; v <= max_rf; ++v) cnt[v + 1] += cnt[v];
            for (int64_t i = 0; i < npairs; ++i) idx[cnt[max_rf - src.fl(i)]++] = i;     // descending target length, stable
            for (int64_t k = 0; k < npairs; ++k) { const int64_t i = idx[k]; order[cursor[bin[i]]++] = i; }
        }
        int64_t cm = 0;
        for (int64_t k = 0; k < npairs; ++k) {
            const int64_t i = order[k];
            SwTask& t = h_tasks[k];
            t.rd_base = src.rd_base(i);
            t.rf_base = src.rf_base(i);
            t.rd_len = (int32_t)src.rl(i); t.rf_len = (int32_t)src.fl(i);
            t.cm_off = cm; cm += t.rf_len;            // column records are laid out in task order
            t.dir = 1; t.out = (int32_t)i; t.stop = 0; t.pad_ = 0;
        }
    }
    for (int c = 0; c <= N_STRIPS + 1; ++c) if (bin_count[c] > 0) b->bins.push_back(BinLaunch{c, bin_first[c], bin_count[c]});
    b->strip_tasks = bin_first[N_STRIPS];
    if (b->strip_tasks == 0) b->revband = false;

    // ---- device buffers + uploads (straight from the caller's buffers: pinned caller memory gives full PCIe rate)
    DevPool& pool = e->pool;
    const size_t reads_bytes = (size_t)src.read_bases();          // bases the traceback arenas are budgeted on
    b->seq_reads_bytes = reads_bytes; b->seq_bytes = src.arena_bytes();
    if (src.shared_arena()) { b->seq.p = src.shared_arena(); b->seq.cap = 0; b->seq_shared = true; }
    else pool.take(b->seq, b->seq_bytes + 16);
    pool.take(b->mask, sizeof(int32_t) * (size_t)(npairs + 1));
    pool.take(b->tasks_fwd, sizeof(SwTask) * (size_t)(npairs + 1)); pool.take(b->tasks_rev, sizeof(SwTask) * (size_t)(npairs + 1));
    pool.take(b->ends_fwd, sizeof(SwEnds) * (size_t)(npairs + 1)); pool.take(b->ends_rev, sizeof(SwEnds) * (size_t)(npairs + 1));
    pool.take(b->colrec, sizeof(uint32_t) * (size_t)(cm_total + 1));
    pool.take(b->fwdres, sizeof(FwdResult) * (size_t)(npairs + 1)); pool.take(b->finalres, sizeof(FinalResult) * (size_t)(npairs + 1));
    pool.take(b->counters, 512 * sizeof(unsigned long long));   // [0,128) misc (arena cursors at 64/65, task cursors at 100..104), [128,256) forward bins, [256,384) reverse bins
    pool.take(b->dmat, (size_t)n * n + 16);
    pool.take(b->relist, 2 * sizeof(int) * (size_t)(npairs + 2));       // pairs refused by the packed kernel (reads with N): [count, task indices...] per pass
    if (b->revband) { pool.take(b->revq_items, sizeof(int) * (size_t)REVBAND_CLASSES * (size_t)(b->strip_tasks + 1)); pool.take(b->revq_meta, 32 * sizeof(int)); }
    if (npairs) {
        if (src.staging_bytes()) pool.take(b->pack_stage, src.staging_bytes());
        if (b->seq_shared) { if (src.shared_ready()) CK(cudaStreamWaitEvent(st, src.shared_ready(), 0)); }
        else src.copy_arena(b->seq.as<int8_t>(), b->pack_stage.as<uint8_t>(), st);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(b->mask.p, masklen, sizeof(int32_t) * npairs, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(b->tasks_fwd.p, h_tasks, sizeof(SwTask) * npairs, cudaMemcpyHostToDevice, st));
    }
    sl.pin_misc.reserve(4096 + (size_t)n * n);
    memcpy(sl.pin_misc.as<char>() + 4096, b->mat.data(), (size_t)n * n);
    CK(cudaMemcpyAsync(b->dmat.p, sl.pin_misc.as<char>() + 4096, (size_t)n * n, cudaMemcpyHostToDevice, st));
    b->h2d_bytes = src.h2d_bytes() + (sizeof(int32_t) + sizeof(SwTask)) * (size_t)npairs + (size_t)n * n;

    // ---- boundary rows of the 32-bit kernel (one slot per resident warp)
    b->wide_blocks = e->sm_count * 3;          // (launch bounds of the 32-bit and the round-1 multi-strip kernels: 3 blocks per SM)
    b->wide_stride = ((long long)max_rf + 63) & ~63ll;
    pool.take(b->wide_boundary, sizeof(int) * (size_t)b->wide_blocks * (WIDE_BLOCK / 32) * 2 * (size_t)b->wide_stride + 256);
    // the multi-strip 16-bit kernel may run concurrently with the 32-bit one (bins of a pass are launched on side streams): own buffer
    b->long_blocks = e->sm_count * 4;          // one boundary slot (2 x wide_stride words) per resident warp of the multi-strip kernel: up to 4 blocks of 4 warps per SM
    if (bin_count[LONG_BIN] > 0) pool.take(b->long_boundary, sizeof(uint32_t) * (size_t)b->long_blocks * (LONG_BLOCK / 32) * 2 * (size_t)b->wide_stride + 256);

    // ---- traceback arenas.  Direction bytes: (2*band+1) per read row; the first attempt has band |dlen|+1 and most pairs
    // stop there.  Budget 16 band cells per read base (+ slack); pairs that do not fit are reported (status 5) and re-run by fetch.
    const bool want_cigar = (p->flag & 7) != 0;
    if ((p->flag & 0xff) != 0) {
        pool.take(b->bandrec, sizeof(BandRec) * (size_t)(npairs + 1)); pool.take(b->flaglist, sizeof(int) * (size_t)(npairs + 1));
        pool.take(b->bandq_items, sizeof(BandItem) * (size_t)ROWS_CLASSES * (size_t)(npairs + 1));     // one work queue per band-width class
        pool.take(b->bandq_meta, 64 * sizeof(int));
    }
    if (want_cigar) {
        int64_t read_bases = (int64_t)reads_bytes;
        b->scratch_bytes = (unsigned long long)read_bases * 24ull + (unsigned long long)npairs * 512ull + (64ull << 20);
        b->cig_cap = (unsigned long long)npairs * 24ull + (unsigned long long)read_bases / 4ull + 4096ull;
        pool.take(b->scratch, b->scratch_bytes);
        // wide-band traceback: a direction region per resident warp (band cells x longest read), reused for every attempt and pair;
        // with ONT-scale reads the grid shrinks so that the regions stay within a few GB
        b->warp_dir_stride = warptr_region_bytes(std::max(max_rd, 1));
        const unsigned long long budget = 6ull << 30;
        long long wb = (long long)(budget / (b->warp_dir_stride * WARPTR_WARPS));
        // warps that can be busy at all: the long reads (always traced by this kernel in small batches) plus a share of wide bands
        wb = std::min<long long>(wb, (n_long + npairs / 8 + 16 + WARPTR_WARPS - 1) / WARPTR_WARPS);
        b->warp_trace_blocks = (int)std::max<long long>(std::min<long long>(wb, (long long)e->sm_count * 4), 1);
        pool.take(b->warp_dir, (size_t)b->warp_dir_stride * WARPTR_WARPS * (size_t)b->warp_trace_blocks + 256);
        pool.take(b->cig, b->cig_cap * sizeof(uint32_t));
    }
    // the task list sits in slot-owned pinned staging that the next upload on this slot overwrites: the phased API waits for
    // its copy here; the chunk pipeline reuses a slot only after fetching (= synchronising) the batch that used it
    if (!b->pipelined) CK(cudaStreamSynchronize(st));
    return b;
}

extern "C" mpn_batch* mpn_batch_upload(mpn_engine* e, const mpn_params* p, const int8_t* reads, const int64_t* read_off, const int8_t* refs,
                                        const int64_t* ref_off, const int32_t* masklen, int64_t npairs)
{
    return upload_impl(e, 0, p, CsrPairs{reads, read_off, refs, ref_off, npairs}, masklen, npairs);
}

extern "C" mpn_batch* mpn_batch_upload_spans(mpn_engine* e, const mpn_params* p, const int8_t* seq, int64_t seq_bytes, const int64_t* rd_start, const int32_t* rd_len,
                                              const int64_t* rf_start, const int32_t* rf_len, const int32_t* masklen, int64_t npairs)
{
    return upload_impl(e, 0, p, SpanPairs{seq, seq_bytes, rd_start, rd_len, rf_start, rf_len, npairs}, masklen, npairs);
}

static int64_t range_chunk_pairs();
static std::vector<int64_t> range_bounds(int64_t npairs);

extern "C" int mpn_align_batch_spans(mpn_engine* e, const mpn_params* p, const int8_t* seq, int64_t seq_bytes, const int64_t* rd_start, const int32_t* rd_len,
                                     const int64_t* rf_start, const int32_t* rf_len, const int32_t* masklen, int64_t npairs,
                                     mpn_result* out, uint32_t* cigar, int64_t cigar_cap)
{
    if (!e || npairs < 0) return MPN_E_ARG;
    static const bool timing = getenv("MPN_TIMING") != nullptr;
    auto now = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    static const bool no_pipe = getenv("MPN_NO_SPANS_PIPE") != nullptr;       // A/B switch
    const int64_t CHUNK = range_chunk_pairs();
    if (npairs > CHUNK + CHUNK / 2 && !no_pipe && seq_bytes > 0) {
        // large batch (the realigner's regions x haplotypes x reads): the arena goes to the device once, the pairs through the chunk pipeline,
        // so scheduling, kernels, D2H and record conversion of different ranges overlap
        SpanPairs all{seq, seq_bytes, rd_start, rd_len, rf_start, rf_len, npairs};
        if (!p || !all.valid()) return MPN_E_ARG;
        CK(cudaSetDevice(e->device));
        DevBuf arena;
        e->pool.take(arena, (size_t)seq_bytes + 16);
        CK(cudaMemcpyAsync(arena.p, seq, (size_t)seq_bytes, cudaMemcpyHostToDevice, e->stream));
        CK(cudaEventRecord(e->ev_arena, e->stream));
        all.dev_arena = arena.as<int8_t>(); all.ready = e->ev_arena;
        const std::vector<int64_t> bounds = range_bounds(npairs);
        size_t c = 0;
        const int rc = mpn::run_ranges(e, p, all, masklen, [&](mpn::RangeJob& r) { if (c + 1 >= bounds.size()) return false; r.first = bounds[c]; r.count = bounds[c + 1] - bounds[c]; r.cig_base = -1; ++c; return r.count > 0; },
                                       out, cigar, cigar_cap, nullptr, nullptr);
        CK(cudaStreamSynchronize(e->stream));
        e->pool.give(arena);
        return rc;
    }
    const double t0 = now();
    mpn_batch* b = mpn_batch_upload_spans(e, p, seq, seq_bytes, rd_start, rd_len, rf_start, rf_len, masklen, npairs);
    if (!b) return MPN_E_ARG;
    const double t1 = now();
    int rc = mpn_batch_run(b);
    const double t2 = now();
    if (timing) CK(cudaStreamSynchronize(b->st));
    const double t3 = now();
    if (rc == 0) rc = mpn_batch_fetch(b, out, cigar, cigar_cap);
    const double t4 = now();
    mpn_batch_free(b);
    if (timing) fprintf(stderr, "[mpn_ssw] spans batch %lld pairs: upload %.3f ms, enqueue %.3f ms, kernels %.3f ms, fetch %.3f ms\n", (long long)npairs, t1 - t0, t2 - t1, t3 - t2, t4 - t3);
    return rc;
}

// list of the pairs the packed kernel refuses (reads with N) for the N variants to walk; nullptr when those variants do not apply
static int* relist_of(mpn_batch* b, bool forward)
{
    if (!b->sc16.ncol_ok) return nullptr;
    return b->relist.as<int>() + (forward ? 0 : b->npairs + 2);
}

// first use of a strip instantiation on this engine's device: raise its dynamic shared-memory limit (the checkpoint slots of
// sw_strip16.cuh need more than the 48 KB default), prefer the all-shared carve-out, and ask how many blocks fit on an SM
static int prepare_strip(mpn_engine* e, size_t slot, StripFn fn, size_t smem)
{
    int& nb = e->strip_blocks[slot];
    if (nb == 0) {
        if (smem > 48 * 1024) CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CK(cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        int q = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&q, fn, STRIP_BLOCK_THREADS, smem));
        nb = q > 0 ? q : 1;
        if (getenv("MPN_VERBOSE")) fprintf(stderr, "[mpn_ssw] device %d: strip kernel slot %zu smem=%zu blocks/SM=%d\n", e->device, slot, smem, nb);
    }
    return nb;
}

static void launch_strips(mpn_batch* b, const SwTask* tasks, bool forward, SwEnds* ends, int counter_base, bool skip_packed = false)
{
    mpn_engine* e = b->e;
    cudaStream_t main_st = b->st;
    int slot = counter_base;
    // fork: with more than one bin, launches alternate over the side streams (largest strips first) and join back on the batch stream
    static const bool pipe_fork = getenv("MPN_NO_PIPE_FORK") == nullptr;    // side streams for the ranges of the chunk pipeline too (A/B switch; with the merged band launch: 57.5 -> 56.0 ms per e2e step)
    const bool fork = b->bins.size() > 1 && (!b->pipelined || pipe_fork);
    if (fork) {
        CK(cudaEventRecord(e->ev_fork, main_st));
        for (int k = 0; k < e->naux; ++k) CK(cudaStreamWaitEvent(e->aux[k], e->ev_fork, 0));
    }
    int turn = 0;
    for (size_t bi = b->bins.size(); bi-- > 0;) {
        const BinLaunch& bl = b->bins[bi];
        if (skip_packed && bl.cfg < N_STRIPS) continue;          // done by the banded reverse kernels
        cudaStream_t st = fork ? e->aux[turn++ % e->naux] : main_st;
        int* counter = reinterpret_cast<int*>(b->counters.as<unsigned long long>() + slot++);
        if (bl.cfg == LONG_BIN) {
            // few long pairs: several warps per pair (strips pipelined across the warps of a block), else one warp per pair
            // measured on 10 kb x 12 kb pairs (profiles/r02_long_ab.md): 1024 pairs -- 4 warps per pair 3571 GCUPS, 1 warp 3288; 2048 pairs -- 2 warps 3953,
            // 1 warp 3581; 4096 pairs -- 1 warp (new kernel, 16 resident warps per SM) 4249, 2 warps 4064
            const int64_t fill = (int64_t)e->sm_count * 16;                         // warps that fill the GPU in the one-warp-per-pair kernel
            int nwp = (b->max_rf >= (1 << 20) || bl.count * 4 > fill * 5) ? 1 : (bl.count * 2 <= fill ? 4 : 2);
            static const int force_nwp = []() { const char* v = getenv("MPN_LONG_NWP"); return v ? atoi(v) : 0; }();      // A/B switch
            if (force_nwp == 1 || force_nwp == 2 || force_nwp == 4) nwp = force_nwp;
            static const bool old_long = getenv("MPN_OLD_LONG16") != nullptr;       // A/B switch: round-1 kernel for one-warp-per-pair batches too
            if (nwp == 1 && !old_long) {
                // throughput case: the multi-strip instantiation of the packed kernel (checkpoint / replay, 4 blocks per SM)
                const StripEntry& c = g_strip_long;
                const int per_sm = prepare_strip(e, (size_t)2 * (N_STRIPS + 4) + (forward ? 0 : 1), forward ? c.fn : c.fn_rev, forward ? c.smem : c.smem_rev);
                const int gpb = STRIP_BLOCK_THREADS / 32;
                const int64_t blocks = std::min<int64_t>({(bl.count + gpb - 1) / gpb, (int64_t)e->sm_count * per_sm, (int64_t)b->long_blocks});
                (forward ? c.fn : c.fn_rev)<<<(unsigned)blocks, STRIP_BLOCK_THREADS, forward ? c.smem : c.smem_rev, st>>>(tasks + bl.first, (int)bl.count, counter, b->seq.as<int8_t>(), b->sc16,
                    forward ? b->colrec.as<uint32_t>() : nullptr, ends, nullptr, (int)bl.first, b->long_boundary.as<uint32_t>(), b->wide_stride);
            } else {
                const int ppb = (LONG_BLOCK / 32) / nwp;
                const int blocks = (int)std::min<int64_t>((bl.count + ppb - 1) / ppb, b->wide_blocks);
                auto fn = nwp == 1 ? sw_long16_kernel<LONG_KR, 1> : (nwp == 2 ? sw_long16_kernel<LONG_KR, 2> : sw_long16_kernel<LONG_KR, 4>);
                fn<<<blocks, LONG_BLOCK, long16_smem_bytes<LONG_KR>(), st>>>(tasks + bl.first, (int)bl.count, counter, b->seq.as<int8_t>(), b->sc16,
                    forward ? b->colrec.as<uint32_t>() : nullptr, ends, b->long_boundary.as<uint32_t>(), b->wide_stride);
            }
            e->wide_pairs += forward ? bl.count : 0;
        } else if (bl.cfg == WIDE_BIN) {
            launch_wide32_impl(tasks + bl.first, (int)bl.count, counter, b->seq.as<int8_t>(), b->dmat.as<int8_t>(), b->p.n, b->fin.gapO, b->fin.gapE,
                               forward ? b->colrec.as<uint32_t>() : nullptr, ends, b->wide_boundary.as<int>(), b->wide_stride, b->wide_blocks, 0, b->max_rd <= 512, st);
            e->wide_pairs += forward ? bl.count : 0;
        } else {
            const StripCfg& c = g_strips[bl.cfg];
            const int groups_per_block = STRIP_BLOCK_THREADS / c.G;
            int64_t blocks = (bl.count + groups_per_block - 1) / groups_per_block;
            const int per_sm = prepare_strip(e, (size_t)2 * bl.cfg + (forward ? 0 : 1), forward ? c.fn : c.fn_rev, forward ? c.smem : c.smem_rev);
            blocks = std::min<int64_t>(blocks, (int64_t)e->sm_count * per_sm);
            (forward ? c.fn : c.fn_rev)<<<(unsigned)blocks, STRIP_BLOCK_THREADS, forward ? c.smem : c.smem_rev, st>>>(tasks + bl.first, (int)bl.count, counter, b->seq.as<int8_t>(), b->sc16,
                                                                  forward ? b->colrec.as<uint32_t>() : nullptr, ends, relist_of(b, forward), (int)bl.first, nullptr, 0ll);
        }
        CK(cudaGetLastError());
        e->launches++;
    }
    if (fork) {
        for (int k = 0; k < e->naux; ++k) {
            CK(cudaEventRecord(e->ev_join[k], e->aux[k]));
            CK(cudaStreamWaitEvent(main_st, e->ev_join[k], 0));
        }
    }
}

// reverse passes of the packed bins through the banded kernels: setup (one queue per band class; what does not qualify goes on the
// reverse list of the N variants, flagged), then one lane-per-pair kernel per class
static void launch_revband(mpn_batch* b)
{
    mpn_engine* e = b->e;
    cudaStream_t st = b->st;
    const int nt = (int)b->strip_tasks;
    CK(cudaMemsetAsync(b->revq_meta.p, 0, 32 * sizeof(int), st));
    RevBandQueues q{b->revq_items.as<int>(), b->revq_meta.as<int>(), b->revq_meta.as<int>() + 16, nt + 1};
    int* relist = b->relist.as<int>() + b->npairs + 2;
    const SwTask* tasks = b->tasks_rev.as<SwTask>();
    SwEnds* ends = b->ends_rev.as<SwEnds>();
    sw_revband_setup_kernel<<<(unsigned)((nt + 127) / 128), 128, 0, st>>>(tasks, nt, b->rbsc, q, ends, relist);
    CK(cudaGetLastError());
    // persistent grids: a warp takes 32 queue items at a time.  Widest class first (its pairs take longest).
    const int64_t warps = ((int64_t)nt + 31) / 32, wpb = REVBAND_BLOCK / 32;
    auto grid = [&](int per_sm) { return (unsigned)std::max<int64_t>(1, std::min<int64_t>((warps + wpb - 1) / wpb, (int64_t)e->sm_count * per_sm)); };
    static const bool rb_merge = getenv("MPN_NO_RB_MERGE") == nullptr;    // A/B switch
    if (b->pipelined && rb_merge) {
        // a range of the chunk pipeline: one launch for all classes
        static int occ_all = 0;
        if (occ_all == 0) { CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_all, sw_revband_all_kernel, REVBAND_BLOCK, 0)); occ_all = std::max(1, occ_all); }
        const unsigned gx = (unsigned)std::max<int64_t>(1, std::min<int64_t>((warps + wpb - 1) / wpb, ((int64_t)e->sm_count * occ_all + REVBAND_CLASSES - 1) / REVBAND_CLASSES * 2));
        sw_revband_all_kernel<<<dim3(gx, REVBAND_CLASSES), REVBAND_BLOCK, 0, st>>>(tasks, b->seq.as<int8_t>(), b->rbsc, q, ends, relist);
        CK(cudaGetLastError());
        e->launches += 2;
        return;
    }
    // the classes side by side on the side streams (a class's tail of half-empty SMs is filled by the next one); a batch of the
    // chunk pipeline stays on its own stream, the other ranges in flight fill its tails
    static const bool no_fork = getenv("MPN_RB_NOFORK") != nullptr;      // A/B switch
    static const bool pipe_fork = getenv("MPN_NO_PIPE_FORK") == nullptr;
    const bool fork = (!b->pipelined || pipe_fork) && !no_fork;
    cudaStream_t s16 = st, s12 = st, s8 = st;
    if (fork) {
        CK(cudaEventRecord(e->ev_fork, st));
        for (int k = 0; k < 3; ++k) CK(cudaStreamWaitEvent(e->aux[k], e->ev_fork, 0));
        s16 = e->aux[0]; s12 = e->aux[1]; s8 = e->aux[2];
    }
    int* occ = e->revband_blocks;
    if (occ[0] == 0) {
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[0], sw_revband_kernel<4>, REVBAND_BLOCK, 0));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[1], sw_revband_kernel<8>, REVBAND_BLOCK, 0));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[2], sw_revband_kernel<12>, REVBAND_BLOCK, 0));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[3], sw_revband_kernel<16>, REVBAND_BLOCK, 0));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[4], sw_revband_kernel<20>, REVBAND_BLOCK, 0));
        for (int k = 0; k < 5; ++k) occ[k] = std::max(1, std::min(occ[k], 8));
    }
    sw_revband_kernel<20><<<grid(occ[4]), REVBAND_BLOCK, 0, s16>>>(tasks, b->seq.as<int8_t>(), b->rbsc, q, ends, relist);
    sw_revband_kernel<16><<<grid(occ[3]), REVBAND_BLOCK, 0, s16>>>(tasks, b->seq.as<int8_t>(), b->rbsc, q, ends, relist);
    sw_revband_kernel<12><<<grid(occ[2]), REVBAND_BLOCK, 0, s12>>>(tasks, b->seq.as<int8_t>(), b->rbsc, q, ends, relist);
    sw_revband_kernel<8><<<grid(occ[1]), REVBAND_BLOCK, 0, s8>>>(tasks, b->seq.as<int8_t>(), b->rbsc, q, ends, relist);
    sw_revband_kernel<4><<<grid(occ[0]), REVBAND_BLOCK, 0, st>>>(tasks, b->seq.as<int8_t>(), b->rbsc, q, ends, relist);
    CK(cudaGetLastError());
    if (fork) {
        for (int k = 0; k < 3; ++k) {
            CK(cudaEventRecord(e->ev_join[k], e->aux[k]));
            CK(cudaStreamWaitEvent(st, e->ev_join[k], 0));
        }
    }
    e->launches += 6;
}

// re-run of the listed pairs in the N variants of the packed kernel (one per group width, each takes its length class)
static void launch_n_variants(mpn_batch* b, const SwTask* tasks, bool forward, SwEnds* ends, int counter_base)
{
    int* relist = relist_of(b, forward);
    if (!relist) return;
    mpn_engine* e = b->e;
    const StripEntry* nv[4] = {&g_strip_n_a, &g_strip_n_b, &g_strip_n_c, &g_strip_n_d};
    int min_len = 0;
    for (int k = 0; k < 4; ++k) {
        const StripEntry& c = *nv[k];
        const int cap = 2 * c.G * c.KR;
        if (b->max_rd > min_len) {
            int* counter = reinterpret_cast<int*>(b->counters.as<unsigned long long>() + counter_base + k);
            const int per_sm = prepare_strip(e, (size_t)2 * (N_STRIPS + k) + (forward ? 0 : 1), forward ? c.fn : c.fn_rev, forward ? c.smem : c.smem_rev);
            // (reverse passes in band mode: these kernels take every pair the band refuses, which may be most of a noisy batch -> full grid)
            const unsigned nblocks = (unsigned)(e->sm_count * ((!forward && b->revband) ? per_sm : std::min(per_sm, 2)));
            (forward ? c.fn : c.fn_rev)<<<nblocks, STRIP_BLOCK_THREADS, forward ? c.smem : c.smem_rev, b->st>>>(tasks, 0, counter, b->seq.as<int8_t>(), b->sc16,
                                                                  forward ? b->colrec.as<uint32_t>() : nullptr, ends, relist, min_len, nullptr, 0ll);
            CK(cudaGetLastError());
            e->launches++;
        }
        min_len = cap;
    }
}

extern "C" int mpn_batch_run(mpn_batch* b)
{
    if (!b) return MPN_E_ARG;
    mpn_engine* e = b->e;
    CK(cudaSetDevice(e->device));
    cudaStream_t st = b->st;
    const int64_t n = b->npairs;
    if (n == 0) { b->ran = true; return 0; }
    CK(cudaMemsetAsync(b->counters.p, 0, 512 * sizeof(unsigned long long), st));
    if (b->sc16.ncol_ok) {
        CK(cudaMemsetAsync(relist_of(b, true), 0, sizeof(int), st));
        CK(cudaMemsetAsync(relist_of(b, false), 0, sizeof(int), st));
    }
    PairArrays pa{b->mask.as<int32_t>()};

    if (e->profile) { e->ev = e->evs[e->ev_runs % mpn_engine::NEVSET]; CK(cudaEventRecord(e->ev[0], st)); e->ev_valid = 1; }
    // forward score pass -> ends + column records
    launch_strips(b, b->tasks_fwd.as<SwTask>(), true, b->ends_fwd.as<SwEnds>(), 128);
    // pairs the packed kernel refused (read code >= 4): its N variants take reads with N when the matrix's N column is constant,
    // whatever is still flagged after that is re-run in the 32-bit kernel before anything reads the ends
    launch_n_variants(b, b->tasks_fwd.as<SwTask>(), true, b->ends_fwd.as<SwEnds>(), 110);
    launch_wide32_impl(b->tasks_fwd.as<SwTask>(), (int)n, reinterpret_cast<int*>(b->counters.as<unsigned long long>() + 100),
                       b->seq.as<int8_t>(), b->dmat.as<int8_t>(), b->p.n, b->fin.gapO, b->fin.gapE, b->colrec.as<uint32_t>(), b->ends_fwd.as<SwEnds>(),
                       b->wide_boundary.as<int>(), b->wide_stride, b->wide_blocks, 1, b->max_rd <= 512, st);
    CK(cudaGetLastError());
    e->launches++;
    if (e->profile) { CK(cudaEventRecord(e->ev[1], st)); e->ev_valid = 2; }
    // second best + mode + reverse tasks
    {
        const int warps_per_block = 4;
        const unsigned blocks = (unsigned)((n + warps_per_block - 1) / warps_per_block);
        sw_finish_kernel<<<blocks, 128, 0, st>>>(b->tasks_fwd.as<SwTask>(), (int)n, b->ends_fwd.as<SwEnds>(), b->colrec.as<uint32_t>(), pa, b->fin,
                                                  b->fwdres.as<FwdResult>(), b->tasks_rev.as<SwTask>());
        CK(cudaGetLastError());
        e->launches++;
    }
    if (e->profile) { CK(cudaEventRecord(e->ev[2], st)); e->ev_valid = 3; }
    const bool any_rev = !(b->fin.flag == 0);
    if (any_rev) {
        if (b->revband) launch_revband(b);
        launch_strips(b, b->tasks_rev.as<SwTask>(), false, b->ends_rev.as<SwEnds>(), 256, b->revband);
        launch_n_variants(b, b->tasks_rev.as<SwTask>(), false, b->ends_rev.as<SwEnds>(), 114);
        launch_wide32_impl(b->tasks_rev.as<SwTask>(), (int)n, reinterpret_cast<int*>(b->counters.as<unsigned long long>() + 101),
                           b->seq.as<int8_t>(), b->dmat.as<int8_t>(), b->p.n, b->fin.gapO, b->fin.gapE, nullptr, b->ends_rev.as<SwEnds>(),
                           b->wide_boundary.as<int>(), b->wide_stride, b->wide_blocks, 1, b->max_rd <= 512, st);
        CK(cudaGetLastError());
        e->launches++;
        if (e->profile) { CK(cudaEventRecord(e->ev[3], st)); e->ev_valid = 4; }
        // lane-per-pair traceback is the throughput path; a few long reads in a batch (the haplotype-vs-reference pairs of a realigner
        // region) would serialise a thousand rows on single lanes, so they take the warp-per-pair kernel unless there are many of them
        const int lane_max_rows = b->n_long_rows <= 4096 ? 512 : NARROW_MAX_ROWS;
        TraceParams tp{b->fin.flag, b->fin.filters, b->p.filterd, b->fin.gapO, b->fin.gapE, b->p.n, b->dmat.as<int8_t>(), lane_max_rows};
        Arena ar{b->scratch.as<uint8_t>(), b->scratch_bytes, b->counters.as<unsigned long long>() + 64};
        {
            // narrow bands: setup (begin positions, filters, one queue per band-width class), then one lane-persistent row kernel per
            // class in increasing order (a failed attempt re-queues the pair for the doubled band = a later class), then one-thread-per-pair traceback
            CK(cudaMemsetAsync(b->bandq_meta.p, 0, 64 * sizeof(int), st));
            BandQueues bq{b->bandq_items.as<BandItem>(), b->bandq_meta.as<int>(), b->bandq_meta.as<int>() + 16, (int)n};
            sw_band_setup_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(b->tasks_fwd.as<SwTask>(), (int)n, b->fwdres.as<FwdResult>(), b->ends_rev.as<SwEnds>(), tp,
                b->cig.as<uint32_t>(), b->cig_cap, b->counters.as<unsigned long long>() + 65, b->finalres.as<FinalResult>(), b->bandrec.as<BandRec>(), b->flaglist.as<int>(),
                reinterpret_cast<int*>(b->counters.as<unsigned long long>() + 103), bq);
            CK(cudaGetLastError());
            {
                const unsigned rows_blocks = (unsigned)std::max<int64_t>(1, std::min<int64_t>((n + ROWS_BLOCK - 1) / ROWS_BLOCK, (int64_t)e->sm_count * 8));
                int* const nflag = reinterpret_cast<int*>(b->counters.as<unsigned long long>() + 103);
#define MPN_ROWS_LAUNCH(NY, ...) sw_band_rows_kernel<__VA_ARGS__><<<dim3(rows_blocks, NY), ROWS_BLOCK, 0, st>>>(b->seq.as<int8_t>(), tp, ar, \
                    b->finalres.as<FinalResult>(), b->bandrec.as<BandRec>(), b->flaglist.as<int>(), nflag, bq)
                // a failed attempt re-queues its pair for the doubled band: widths that feed each other go to successive launches
                MPN_ROWS_LAUNCH(4, 1, 3, 5, 7); MPN_ROWS_LAUNCH(2, 2, 6, 0, 0); MPN_ROWS_LAUNCH(1, 4, 0, 0, 0);
#undef MPN_ROWS_LAUNCH
                CK(cudaGetLastError());
                e->launches += 4;
            }
            sw_band_trace_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(b->tasks_fwd.as<SwTask>(), (int)n, b->fwdres.as<FwdResult>(), b->bandrec.as<BandRec>(), ar,
                b->cig.as<uint32_t>(), b->cig_cap, b->counters.as<unsigned long long>() + 65, b->finalres.as<FinalResult>());
            CK(cudaGetLastError());
            // wide bands: one warp per flagged pair
            sw_trace_warp_kernel<<<b->warp_trace_blocks, 32 * WARPTR_WARPS, warptr_smem_bytes(b->p.n), st>>>(b->tasks_fwd.as<SwTask>(), b->flaglist.as<int>(),
                reinterpret_cast<int*>(b->counters.as<unsigned long long>() + 103), reinterpret_cast<int*>(b->counters.as<unsigned long long>() + 104), b->seq.as<int8_t>(),
                b->fwdres.as<FwdResult>(), tp, b->warp_dir.as<uint8_t>(), b->warp_dir_stride, b->cig.as<uint32_t>(), b->cig_cap, b->counters.as<unsigned long long>() + 65, b->finalres.as<FinalResult>(), b->bandrec.as<BandRec>());
            CK(cudaGetLastError());
            e->launches += 3;
        }
        // bands beyond the warp kernel's 512 cells (status 8) are redone by the generic kernel
        const unsigned blocks = (unsigned)((n + TRACE_BLOCK - 1) / TRACE_BLOCK);
        sw_trace_wide_kernel<<<blocks, TRACE_BLOCK, trace_smem_bytes(b->p.n), st>>>(b->tasks_fwd.as<SwTask>(), (int)n, b->seq.as<int8_t>(), b->fwdres.as<FwdResult>(), b->ends_rev.as<SwEnds>(), tp, ar,
                                               b->cig.as<uint32_t>(), b->cig_cap, b->counters.as<unsigned long long>() + 65, b->finalres.as<FinalResult>(), 1);
        CK(cudaGetLastError());
        e->launches++;
    }
    if (e->profile) {
        if (!any_rev) CK(cudaEventRecord(e->ev[3], st));
        CK(cudaEventRecord(e->ev[4], st)); e->ev_valid = 5; e->ev_runs++;
    }
    b->ran = true;
    e->pairs += n; e->cells += b->total_cells;
    return 0;
}

// fetch: device -> host.  `cigar_base` is added to every cigar_off (chunks of one user batch share the caller's arena);
// *cigar_words receives the number of words this batch appended at cigar[0 ..).
static int fetch_impl(mpn_batch* b, mpn_result* out, uint32_t* cigar, int64_t cigar_cap, int64_t cigar_base, int64_t* cigar_words)
{
    mpn_engine* e = b->e;
    CK(cudaSetDevice(e->device));
    cudaStream_t st = b->st;
    mpn_engine::Slot& sl = *b->slot;
    const int64_t n = b->npairs;
    if (cigar_words) *cigar_words = 0;
    if (n == 0) return 0;
    const bool any_rev = !(b->fin.flag == 0);
    sl.pin_fwd.reserve(sizeof(FwdResult) * (size_t)n);
    CK(cudaMemcpyAsync(sl.pin_fwd.p, b->fwdres.p, sizeof(FwdResult) * n, cudaMemcpyDeviceToHost, st));
    b->d2h_bytes = sizeof(FwdResult) * (size_t)n;
    sl.pin_misc.reserve(4096);
    unsigned long long* used = sl.pin_misc.as<unsigned long long>() + 256;
    used[0] = used[1] = 0;
    if (any_rev) {
        sl.pin_fin.reserve(sizeof(FinalResult) * (size_t)n);
        CK(cudaMemcpyAsync(sl.pin_fin.p, b->finalres.p, sizeof(FinalResult) * n, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(used, b->counters.as<unsigned long long>() + 64, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        b->d2h_bytes += sizeof(FinalResult) * (size_t)n + 16;
    }
    static const bool timing = getenv("MPN_TIMING_FETCH") != nullptr;
    auto now = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double tf0 = now();
    CK(cudaStreamSynchronize(st));
    const double tf1 = now();
    int rc = 0;
    if (any_rev && (b->p.flag & 7)) {
        const unsigned long long words = std::min<unsigned long long>(used[1], b->cig_cap);
        if ((int64_t)words > cigar_cap || (!cigar && words)) return MPN_E_CIGAR_SPACE;
        if (words) CK(cudaMemcpyAsync(cigar, b->cig.p, words * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        b->d2h_bytes += words * sizeof(uint32_t);
        if (cigar_words) *cigar_words = (int64_t)words;
    }
    const FwdResult* h_fwd = sl.pin_fwd.as<FwdResult>();
    const FinalResult* h_fin = sl.pin_fin.as<FinalResult>();
    // record conversion: independent per pair; on a few host threads for large chunks (it is the serial tail of mpn_align_batch's pipeline)
    std::atomic<long long> first_bad(-1);
    const bool report = b->arena_retries > 0;
    auto convert = [&](int64_t i) {
        const FwdResult& f = h_fwd[i];
        mpn_result& r = out[i];
        r.score1 = (uint16_t)f.score1; r.score2 = (uint16_t)f.score2;
        r.ref_end1 = f.ref_end1; r.read_end1 = f.read_end1; r.ref_end2 = f.ref_end2;
        r.ref_begin1 = -1; r.read_begin1 = -1; r.cigar_len = 0; r.cigar_off = 0;
        r.status = f.status ? MPN_ST_NULL : MPN_ST_OK;      // forward pass: 8-bit overflow without a 16-bit profile (ssw.c:793-796)
        if (any_rev && f.status == 0) {
            const FinalResult& g = h_fin[i];
            r.ref_begin1 = g.ref_begin1; r.read_begin1 = g.read_begin1;
            r.cigar_len = g.cigar_len; r.cigar_off = g.cigar_len > 0 ? g.cigar_off + cigar_base : 0;
            if (g.status == 3) r.status = MPN_ST_NULL_TRACE;   // traceback met an invalid direction (ssw.c:840-843)
            else if (g.status != 0) {
                long long none = -1;
                if (report || (g.status != 5 && g.status != 6)) first_bad.compare_exchange_strong(none, (long long)i);
                else { long long none2 = -1; first_bad.compare_exchange_strong(none2, -2 - (long long)i); }       // arena exhausted: silent on the first attempt
            }
        }
    };
    if (n >= 65536) mpn::parallel_for(n, 16384, convert, 4);
    else for (int64_t i = 0; i < n; ++i) convert(i);
    if (first_bad.load() != -1) {
        rc = MPN_E_UNSUPPORTED;
        const long long fb = first_bad.load();
        if (fb >= 0) fprintf(stderr, "[mpn_ssw] traceback could not complete (at pair %lld, status %d: 5/6 = arena exhausted, 7/8 = band beyond kernel limits)\n", fb, h_fin[fb].status);
    }
    const double tf2 = now();
    CK(cudaStreamSynchronize(st));      // CIGAR arena copy
    if (timing) fprintf(stderr, "[mpn_ssw] fetch %lld pairs: wait %.3f ms, convert %.3f ms, cigar wait %.3f ms\n", (long long)n, tf1 - tf0, tf2 - tf1, now() - tf2);
    if (rc == MPN_E_UNSUPPORTED && b->arena_retries == 0) {
        // An arena of the traceback ran out (statuses 5 / 6: direction words or CIGAR words beyond the typical-case budget, e.g.
        // reads made of alternating indels).  Size both for the worst case and run the batch once more; results are deterministic.
        bool arena = false;
        for (int64_t i = 0; i < n && !arena; ++i) arena = h_fwd[i].status == 0 && (h_fin[i].status == 5 || h_fin[i].status == 6);
        if (arena) {
            b->arena_retries = 1;
            b->scratch_bytes = b->scratch_bytes * 8ull + (unsigned long long)b->sum_rd * 64ull;
            b->cig_cap = (unsigned long long)(b->sum_rd + b->sum_rf) + 2ull * (unsigned long long)n + 4096ull;
            e->pool.take(b->scratch, b->scratch_bytes);
            e->pool.take(b->cig, b->cig_cap * sizeof(uint32_t));
            mpn_batch_run(b);
            return fetch_impl(b, out, cigar, cigar_cap, cigar_base, cigar_words);
        }
    }
    return rc;
}

extern "C" int mpn_batch_fetch(mpn_batch* b, mpn_result* out, uint32_t* cigar, int64_t cigar_cap)
{
    if (!b || (!out && b->npairs)) return MPN_E_ARG;
    if (!b->ran) return MPN_E_ARG;
    return fetch_impl(b, out, cigar, cigar_cap, 0, nullptr);
}

// ---- nibble-packed input (mpn_align_batch_packed4): expand `nbases` nibbles starting at nibble `first` of src into int8 codes
__global__ void __launch_bounds__(256) unpack4_kernel(const uint8_t* __restrict__ src, int64_t first, int64_t nbases, int8_t* __restrict__ dst, int mis)
{
    // 16 output bases per thread, laid out so that the interior threads store one ALIGNED 16-byte word (dst itself may sit at any byte of the
    // arena: `mis` = its address modulo 16); the nibbles of a thread lie in 8 or 9 staging bytes (the staging buffer is padded by 16 bytes)
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t k0 = t * 16 - mis;               // first output base of this thread (negative part of thread 0 is skipped)
    if (k0 >= nbases) return;
    const int64_t ka = k0 < 0 ? 0 : k0;
    const int64_t n0 = first + ka;                 // nibble index of the first base handled
    const uint8_t* p = src + (n0 >> 1);
    unsigned long long lo = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) lo |= (unsigned long long)p[q] << (8 * q);
    if (n0 & 1) lo = (lo >> 4) | ((unsigned long long)p[8] << 60);
    uint32_t w[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint32_t x = (uint32_t)(lo >> (16 * q)) & 0xffffu;             // 4 nibbles -> 4 bytes
        w[q] = (x & 0xfu) | ((x & 0xf0u) << 4) | ((x & 0xf00u) << 8) | ((x & 0xf000u) << 12);
    }
    if (k0 >= 0 && k0 + 16 <= nbases) *reinterpret_cast<uint4*>(dst + k0) = make_uint4(w[0], w[1], w[2], w[3]);
    else {
        const int cnt = (int)(((k0 + 16 < nbases) ? k0 + 16 : nbases) - ka);
        for (int q = 0; q < cnt; ++q) dst[ka + q] = (int8_t)((w[q >> 2] >> (8 * (q & 3))) & 0xffu);
    }
}

void mpn::launch_unpack4(const uint8_t* src, int64_t first_base, int64_t nbases, int8_t* dst, cudaStream_t st)
{
    if (nbases <= 0) return;
    const int mis = (int)(reinterpret_cast<uintptr_t>(dst) & 15u);
    const int64_t threads = (nbases + mis + 15) / 16;
    unpack4_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(src, first_base, nbases, dst, mis);
}

extern "C" void mpn_pack4(const int8_t* codes, int64_t n, uint8_t* out)
{
    mpn::parallel_for((n + 1) / 2, 1 << 20, [&](int64_t k) {
        const int64_t i = 2 * k;
        out[k] = (uint8_t)((codes[i] & 15) | ((i + 1 < n ? codes[i + 1] & 15 : 0) << 4));
    }, 8);
}

// ---- 2-bit packed input (mpn_align_batch_packed2): expand `nbases` 2-bit codes starting at code `first` of src into int8 codes
__global__ void __launch_bounds__(256) unpack2_kernel(const uint8_t* __restrict__ src, int64_t first, int64_t nbases, int8_t* __restrict__ dst, int mis)
{
    // same layout as unpack4_kernel: 16 output bases per thread, interior threads store one aligned 16-byte word; the 16 codes of a thread
    // lie in 4 or 5 staging bytes (the staging buffer is padded)
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t k0 = t * 16 - mis;
    if (k0 >= nbases) return;
    const int64_t ka = k0 < 0 ? 0 : k0;
    const int64_t n0 = first + ka;                 // index of the first code handled
    const uint8_t* p = src + (n0 >> 2);
    unsigned long long v = 0;
#pragma unroll
    for (int q = 0; q < 5; ++q) v |= (unsigned long long)p[q] << (8 * q);
    v >>= 2 * (unsigned)(n0 & 3);
    uint32_t w[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint32_t x = (uint32_t)(v >> (8 * q)) & 0xffu;                 // 4 codes -> 4 bytes
        w[q] = (x & 0x3u) | ((x & 0xcu) << 6) | ((x & 0x30u) << 12) | ((x & 0xc0u) << 18);
    }
    if (k0 >= 0 && k0 + 16 <= nbases) *reinterpret_cast<uint4*>(dst + k0) = make_uint4(w[0], w[1], w[2], w[3]);
    else {
        const int cnt = (int)(((k0 + 16 < nbases) ? k0 + 16 : nbases) - ka);
        for (int q = 0; q < cnt; ++q) dst[ka + q] = (int8_t)((w[q >> 2] >> (8 * (q & 3))) & 0xffu);
    }
}

void mpn::launch_unpack2(const uint8_t* src, int64_t first_base, int64_t nbases, int8_t* dst, cudaStream_t st)
{
    if (nbases <= 0) return;
    const int mis = (int)(reinterpret_cast<uintptr_t>(dst) & 15u);
    const int64_t threads = (nbases + mis + 15) / 16;
    unpack2_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(src, first_base, nbases, dst, mis);
}

// codes that do not fit two bits: entry = position << 4 | code, position counted in the caller's stream; base_pos = stream position of dst[0]
__global__ void __launch_bounds__(256) patch_exceptions_kernel(const int64_t* __restrict__ exc, int64_t n, int64_t base_pos, int8_t* __restrict__ dst)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) dst[(exc[t] >> 4) - base_pos] = (int8_t)(exc[t] & 15);
}

void mpn::launch_patch_exceptions(const int64_t* exc_dev, int64_t n, int64_t base_pos, int8_t* dst, cudaStream_t st)
{
    if (n <= 0) return;
    patch_exceptions_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(exc_dev, n, base_pos, dst);
}

extern "C" int64_t mpn_pack2(const int8_t* codes, int64_t n, uint8_t* out, int64_t* exc, int64_t exc_cap)
{
    mpn::parallel_for((n + 3) / 4, 1 << 20, [&](int64_t k) {
        uint8_t b = 0;
        for (int q = 0; q < 4; ++q) { const int64_t i = 4 * k + q; if (i < n && (codes[i] & 15) < 4) b |= (uint8_t)((codes[i] & 3) << (2 * q)); }
        out[k] = b;
    }, 8);
    int64_t ne = 0;
    for (int64_t i = 0; i < n; ++i)
        if ((codes[i] & 15) >= 4) { if (ne < exc_cap && exc) exc[ne] = (i << 4) | (codes[i] & 15); ++ne; }
    return ne;          // > exc_cap: call again with a larger array
}

// Range boundaries of a large batch for the chunk pipeline: equal ranges of about CHUNK pairs, except that the first ones are an eighth, a
// quarter and a half of that, so the GPU starts after a short upload instead of waiting for a full range to be scheduled and copied, and
// the last ones shrink the same way, so little is left to copy back and convert after the last kernel ends.
static int64_t range_chunk_pairs()
{
    static const int64_t CHUNK = []() { const char* v = getenv("MPN_CHUNK_PAIRS"); const long long c = v ? atoll(v) : 0; return c >= 1024 ? (int64_t)c : (int64_t)196608; }();
    return CHUNK;
}
static std::vector<int64_t> range_bounds(int64_t npairs)
{
    const int64_t CHUNK = range_chunk_pairs();
    static const int LADDER = []() { const char* v = getenv("MPN_CHUNK_LADDER"); return v ? atoi(v) : 3; }();     // steps of the ramp at either end (A/B on config 2: 2 -> 75.6 ms, 3 -> 74.2 ms, 4 -> 79.9 ms per step)
    std::vector<int64_t> bounds(1, 0);
    if (npairs <= CHUNK + CHUNK / 2) { bounds.push_back(npairs); return bounds; }
    int64_t at = 0, tail = 0;
    std::vector<int64_t> tail_sizes;
    for (int k = LADDER; k >= 1; --k) {
        const int64_t want = CHUNK >> k;
        if (want >= 4096 && npairs - at - tail > 3 * CHUNK) { at += want; bounds.push_back(at); tail_sizes.push_back(want); tail += want; }
    }
    const int64_t rest = npairs - at - tail, nrest = (rest + CHUNK - 1) / CHUNK, per = (rest + nrest - 1) / nrest;
    while (at < npairs - tail) { at = std::min(npairs - tail, at + per); bounds.push_back(at); }
    for (size_t k = tail_sizes.size(); k-- > 0;) { at += tail_sizes[k]; bounds.push_back(at); }
    return bounds;
}

extern "C" int mpn_align_batch_packed4(mpn_engine* e, const mpn_params* p, const uint8_t* reads4, const int64_t* read_off, const uint8_t* refs4,
                                       const int64_t* ref_off, const int32_t* masklen, int64_t npairs, mpn_result* out, uint32_t* cigar, int64_t cigar_cap)
{
    if (!e || !p || npairs < 0 || p->n > 16) return MPN_E_ARG;
    const Csr4Pairs all{reads4, read_off, refs4, ref_off, npairs};
    const std::vector<int64_t> bounds = range_bounds(npairs);
    size_t c = 0;
    return mpn::run_ranges(e, p, all, masklen, [&](mpn::RangeJob& r) { if (c + 1 >= bounds.size()) return false; r.first = bounds[c]; r.count = bounds[c + 1] - bounds[c]; r.cig_base = -1; ++c; return r.count > 0; },
                           out, cigar, cigar_cap, nullptr, nullptr);
}

extern "C" int mpn_align_batch_packed2(mpn_engine* e, const mpn_params* p, const uint8_t* reads2, const int64_t* read_off, const int64_t* read_exc, int64_t n_read_exc,
                                       const uint8_t* refs2, const int64_t* ref_off, const int64_t* ref_exc, int64_t n_ref_exc,
                                       const int32_t* masklen, int64_t npairs, mpn_result* out, uint32_t* cigar, int64_t cigar_cap)
{
    if (!e || !p || npairs < 0 || n_read_exc < 0 || n_ref_exc < 0) return MPN_E_ARG;
    const Csr2Pairs all{reads2, read_off, refs2, ref_off, npairs, read_exc, n_read_exc, ref_exc, n_ref_exc};
    if (!all.valid()) return MPN_E_ARG;
    const std::vector<int64_t> bounds = range_bounds(npairs);
    size_t c = 0;
    return mpn::run_ranges(e, p, all, masklen, [&](mpn::RangeJob& r) { if (c + 1 >= bounds.size()) return false; r.first = bounds[c]; r.count = bounds[c + 1] - bounds[c]; r.cig_base = -1; ++c; return r.count > 0; },
                           out, cigar, cigar_cap, nullptr, nullptr);
}

extern "C" int mpn_batch_io_bytes(const mpn_batch* b, int64_t* h2d, int64_t* d2h)
{
    if (!b) return MPN_E_ARG;
    if (h2d) *h2d = (int64_t)b->h2d_bytes;
    if (d2h) *d2h = (int64_t)b->d2h_bytes;
    return 0;
}

extern "C" int mpn_align_batch(mpn_engine* e, const mpn_params* p, const int8_t* reads, const int64_t* read_off, const int8_t* refs,
                               const int64_t* ref_off, const int32_t* masklen, int64_t npairs, mpn_result* out, uint32_t* cigar, int64_t cigar_cap)
{
    if (!e || npairs < 0) return MPN_E_ARG;
    // Small batches: one chunk on the engine stream.  Large batches: ranges of ~192 k pairs go through the pipeline slots
    // (own stream + own pinned staging each), so the H2D copies and host-side scheduling of range k+1 overlap the kernels of range k
    // and the D2H of range k-1.
    const int64_t CHUNK = range_chunk_pairs();
    if (npairs <= CHUNK + CHUNK / 2) {
        mpn_batch* b = mpn_batch_upload(e, p, reads, read_off, refs, ref_off, masklen, npairs);
        if (!b) return MPN_E_ARG;
        int rc = mpn_batch_run(b);
        if (rc == 0) rc = mpn_batch_fetch(b, out, cigar, cigar_cap);
        mpn_batch_free(b);
        return rc;
    }
    const std::vector<int64_t> bounds = range_bounds(npairs);
    const int64_t nchunks = (int64_t)bounds.size() - 1;
    int64_t c = 0;
    return mpn::run_ranges(e, p, CsrPairs{reads, read_off, refs, ref_off, npairs}, masklen,
                           [&](mpn::RangeJob& r) { if (c >= nchunks) return false; r.first = bounds[c]; r.count = bounds[c + 1] - bounds[c]; r.cig_base = -1; ++c; return r.count > 0; },
                           out, cigar, cigar_cap, nullptr, nullptr);
}

template <class Pairs>
int mpn::run_ranges(mpn_engine* e, const mpn_params* p, const Pairs& all, const int32_t* masklen, const std::function<bool(RangeJob&)>& next,
                    mpn_result* out, uint32_t* cigar, int64_t cigar_cap, int64_t* pairs_done, int64_t* cells_done)
{
    // ring of in-flight ranges, one pipeline slot each (own stream + own pinned staging).  Several ranges are queued on the GPU at any
    // time, so the persistent grids of range k+1 fill the SMs that the tail of range k leaves idle, and the host work of a range
    // (scheduling, H2D enqueue, D2H + record conversion) hides behind the kernels of the others.  Fetch order = issue order.
    constexpr int MAXDEPTH = mpn_engine::NSLOT - 1;
    static const int DEPTH = []() { const char* v = getenv("MPN_PIPE_DEPTH"); const int d = v ? atoi(v) : 4; return d >= 1 && d <= MAXDEPTH ? d : 4; }();
    static const bool timing = getenv("MPN_TIMING") != nullptr;
    auto now = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_begin = now();
    double t_upload = 0, t_run = 0, t_drain = 0, t_first = 0;
    mpn_batch* inflight[MAXDEPTH] = {};
    RangeJob job_of[MAXDEPTH];
    int64_t cig_cursor = 0, npairs = 0, ncells = 0;
    int rc = 0;
    auto drain = [&](int s) {
        if (!inflight[s]) return;
        const double td = now();
        int64_t words = 0;
        const RangeJob& j = job_of[s];
        const int64_t base = j.cig_base < 0 ? cig_cursor : j.cig_base, cap = j.cig_base < 0 ? cigar_cap - cig_cursor : j.cig_cap;
        const int r2 = fetch_impl(inflight[s], out + j.first, cigar ? cigar + base : nullptr, cap, base, &words);
        if (r2 != 0 && rc == 0) rc = r2;
        if (j.cig_base < 0) cig_cursor += words;
        npairs += inflight[s]->npairs; ncells += inflight[s]->total_cells;
        mpn_batch_free(inflight[s]);
        inflight[s] = nullptr;
        t_drain += now() - td;
    };
    int64_t issued = 0;
    for (;; ++issued) {
        const int s = (int)(issued % DEPTH);
        drain(s);                                   // the oldest range (issued - DEPTH) used this slot
        RangeJob j;
        if (rc != 0 || !next(j)) break;
        const double tu = now();
        mpn_batch* b = upload_impl(e, 1 + s, p, all.slice(j.first, j.count), masklen + j.first, j.count);
        if (!b) { rc = MPN_E_ARG; break; }
        const double tr = now();
        mpn_batch_run(b);
        inflight[s] = b; job_of[s] = j;
        t_upload += tr - tu; t_run += now() - tr;
        if (issued == 0) t_first = now() - t_begin;
    }
    const double t_issued = now();
    // the remaining ranges, oldest first
    for (int64_t c = std::max<int64_t>(0, issued - DEPTH); c < issued + DEPTH; ++c) drain((int)(c % DEPTH));
    if (timing) fprintf(stderr, "[mpn_ssw] device %d: %lld pairs in %lld ranges: total %.2f ms (first range enqueued at %.2f, all issued at %.2f); host: upload %.2f, enqueue %.2f, drain (wait + copy + convert) %.2f\n",
                        e->device, (long long)npairs, (long long)issued, now() - t_begin, t_first, t_issued - t_begin, t_upload, t_run, t_drain);
    if (pairs_done) *pairs_done = npairs;
    if (cells_done) *cells_done = ncells;
    return rc;
}
template int mpn::run_ranges<mpn::CsrPairs>(mpn_engine*, const mpn_params*, const mpn::CsrPairs&, const int32_t*, const std::function<bool(mpn::RangeJob&)>&, mpn_result*, uint32_t*, int64_t, int64_t*, int64_t*);
template int mpn::run_ranges<mpn::Csr2Pairs>(mpn_engine*, const mpn_params*, const mpn::Csr2Pairs&, const int32_t*, const std::function<bool(mpn::RangeJob&)>&, mpn_result*, uint32_t*, int64_t, int64_t*, int64_t*);
template int mpn::run_ranges<mpn::Csr4Pairs>(mpn_engine*, const mpn_params*, const mpn::Csr4Pairs&, const int32_t*, const std::function<bool(mpn::RangeJob&)>&, mpn_result*, uint32_t*, int64_t, int64_t*, int64_t*);
template int mpn::run_ranges<mpn::SpanPairs>(mpn_engine*, const mpn_params*, const mpn::SpanPairs&, const int32_t*, const std::function<bool(mpn::RangeJob&)>&, mpn_result*, uint32_t*, int64_t, int64_t*, int64_t*);

int mpn::engine_device(const mpn_engine* e) { return e ? e->device : -1; }

// mirror of the host classifier in upload_impl: which score kernel a pair takes, as a relative cost per cell
double mpn::pair_cost_per_cell(int n, int maxpos, int64_t rl, int64_t fl)
{
    if (n > 8) return 4.0;                                                     // 32-bit kernel
    if ((std::min(rl, fl) + 1) * (int64_t)maxpos <= 32767 && rl <= 1280) return 1.0;      // packed short-read kernel
    return 1.55;                                                               // multi-strip kernel with the int16 clamp
}


// ------------------------------------------------------------------------------------------------ k-mer fast pass of the realigner
extern "C" int mpn_fastpass(mpn_engine* e, const char* text, int64_t text_bytes,
                            const int64_t* hap_start, const int32_t* hap_len, const uint8_t* hap_is_ref, int32_t nhaps,
                            const int64_t* read_start, const int32_t* read_len, int32_t nreads,
                            const mpn_fp_region* regions, int32_t nregions,
                            mpn_placement* places, int32_t* hap_score, uint8_t* region_flag)
{
    if (!e || nhaps < 0 || nreads < 0 || nregions < 0 || text_bytes < 0) return MPN_E_ARG;
    if (nregions == 0 || nhaps == 0) return 0;
    if (!text || !hap_start || !hap_len || !hap_is_ref || !regions || !places || !hap_score || !region_flag || (nreads > 0 && (!read_start || !read_len))) return MPN_E_ARG;
    static_assert(sizeof(FpRegion) == sizeof(mpn_fp_region) && sizeof(mpn_placement) == sizeof(int2), "ABI structs mirror the kernel's");
    int max_hap = 0;
    for (int h = 0; h < nhaps; ++h) {
        if (hap_len[h] < 0 || hap_start[h] < 0 || hap_start[h] + hap_len[h] > text_bytes) return MPN_E_ARG;
        if (hap_len[h] > MPN_FP_MAX_HAP) return MPN_E_UNSUPPORTED;
        max_hap = std::max(max_hap, hap_len[h]);
    }
    for (int r = 0; r < nreads; ++r) {
        if (read_len[r] < 0 || read_start[r] < 0 || read_start[r] + read_len[r] > text_bytes) return MPN_E_ARG;
        if (read_len[r] > MPN_FP_MAX_READ) return MPN_E_UNSUPPORTED;
    }
    std::vector<FpHap> haps((size_t)nhaps, FpHap{0, 0, -1, 0, 0});
    int64_t nplaces = 0;
    for (int g = 0; g < nregions; ++g) {
        const mpn_fp_region& rg = regions[g];
        if (rg.nhap < 0 || rg.nread < 0 || rg.hap_first < 0 || rg.read_first < 0 || rg.hap_first + rg.nhap > nhaps || rg.read_first + rg.nread > nreads || rg.place_first < 0) return MPN_E_ARG;
        for (int h = 0; h < rg.nhap; ++h) haps[(size_t)(rg.hap_first + h)] = FpHap{hap_start[rg.hap_first + h], hap_len[rg.hap_first + h], g, h, hap_is_ref[rg.hap_first + h] ? 1 : 0};
        nplaces = std::max<int64_t>(nplaces, rg.place_first + (int64_t)rg.nhap * rg.nread);
    }
    for (const FpHap& h : haps) if (h.region < 0) return MPN_E_ARG;          // every haplotype belongs to a region
    // one block per haplotype, heaviest (reads x length) first so that the last wave of blocks is the light one
    std::stable_sort(haps.begin(), haps.end(), [&](const FpHap& a, const FpHap& b) {
        return (int64_t)regions[a.region].nread * (a.len + 256) > (int64_t)regions[b.region].nread * (b.len + 256); });
    static const bool timing = getenv("MPN_TIMING") != nullptr;
    auto now = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t0 = now();
    CK(cudaSetDevice(e->device));
    cudaStream_t st = e->stream;
    if (!e->fp_attr_done) {
        CK(cudaFuncSetAttribute(fastpass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fp_smem_bytes(MPN_FP_MAX_HAP)));
        CK(cudaEventCreate(&e->fp_ev[0])); CK(cudaEventCreate(&e->fp_ev[1]));
        e->fp_attr_done = true;
    }
    e->pool.take(e->fp_text, (size_t)text_bytes + 16);
    e->pool.take(e->fp_haps, sizeof(FpHap) * (size_t)nhaps);
    e->pool.take(e->fp_regions, sizeof(FpRegion) * (size_t)nregions);
    e->pool.take(e->fp_rstart, sizeof(int64_t) * (size_t)std::max(nreads, 1));
    e->pool.take(e->fp_rlen, sizeof(int32_t) * (size_t)std::max(nreads, 1));
    e->pool.take(e->fp_places, sizeof(int2) * (size_t)std::max<int64_t>(nplaces, 1));
    e->pool.take(e->fp_score, sizeof(int) * (size_t)nhaps);
    e->pool.take(e->fp_flag, sizeof(int) * (size_t)nregions);
    CK(cudaMemcpyAsync(e->fp_text.p, text, (size_t)text_bytes, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(e->fp_haps.p, haps.data(), sizeof(FpHap) * (size_t)nhaps, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(e->fp_regions.p, regions, sizeof(FpRegion) * (size_t)nregions, cudaMemcpyHostToDevice, st));
    if (nreads > 0) {
        CK(cudaMemcpyAsync(e->fp_rstart.p, read_start, sizeof(int64_t) * (size_t)nreads, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(e->fp_rlen.p, read_len, sizeof(int32_t) * (size_t)nreads, cudaMemcpyHostToDevice, st));
    }
    CK(cudaMemsetAsync(e->fp_flag.p, 0, sizeof(int) * (size_t)nregions, st));
    CK(cudaEventRecord(e->fp_ev[0], st));
    fastpass_kernel<<<(unsigned)nhaps, FP_BLOCK, fp_smem_bytes(max_hap), st>>>(e->fp_text.as<char>(), e->fp_haps.as<FpHap>(), e->fp_regions.as<FpRegion>(),
        e->fp_rstart.as<long long>(), e->fp_rlen.as<int>(), e->fp_places.as<int2>(), e->fp_score.as<int>(), e->fp_flag.as<int>());
    CK(cudaGetLastError());
    CK(cudaEventRecord(e->fp_ev[1], st));
    e->launches++;
    std::vector<int> flags((size_t)nregions);
    if (nplaces > 0) CK(cudaMemcpyAsync(places, e->fp_places.p, sizeof(int2) * (size_t)nplaces, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(hap_score, e->fp_score.p, sizeof(int) * (size_t)nhaps, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(flags.data(), e->fp_flag.p, sizeof(int) * (size_t)nregions, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaEventElapsedTime(&e->fp_kernel_ms, e->fp_ev[0], e->fp_ev[1]));
    for (int g = 0; g < nregions; ++g) region_flag[g] = flags[(size_t)g] ? 1 : 0;
    if (timing) fprintf(stderr, "[mpn_ssw] fast pass: %d haplotypes x reads of %d regions, %lld placements, text %.1f MB: call %.3f ms, kernel %.3f ms\n",
                        nhaps, nregions, (long long)nplaces, text_bytes / 1e6, now() - t0, e->fp_kernel_ms);
    return 0;
}

extern "C" float mpn_fastpass_last_kernel_ms(const mpn_engine* e) { return e ? e->fp_kernel_ms : 0.f; }
