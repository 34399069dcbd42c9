// Multi-device pool (include/mpn_ssw_batch.h, mpn_pool_*): ONE caller batch sharded over the GPUs of a box.
// Replaces the process fan-out of the reference (bin/realignment/realignment.sh:34-39, 50-60: GNU parallel, one process per
// chromosome / candidate position) by one engine + one host thread per device inside one process:
//   * the batch is cut, in the caller's order, into ranges of roughly equal cost (forward cells x relative cost of the score kernel the
//     pair takes); the ranges are sorted by cost, heaviest first, and the device threads pull them from one queue, so every device
//     stays busy until the queue is empty whatever its speed (the last ranges are the light ones);
//   * inside a range the engine bins by read length and sorts by target length as for any batch (engine.cu);
//   * there is no collective and no peer traffic: every device copies its ranges straight from the caller's buffers and writes its
//     records straight into the caller's arrays at the range's position (the "host gather" is by construction); CIGAR words go to a
//     per-range region of the caller's arena, mpn_result::cigar_off is the absolute index as everywhere.
// No CPU alignment code here either: a pool without a CUDA device is NULL.
#include "engine_internal.h"
#include "host_shared.h"
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <thread>
#include <vector>

using namespace mpn;

namespace {
struct PoolJob {
    const mpn_params* p = nullptr;
    const CsrPairs* csr = nullptr;
    const SpanPairs* spans = nullptr;
    const int32_t* masklen = nullptr;
    std::vector<RangeJob> ranges;            // heaviest first
    std::atomic<size_t> next{0};
    mpn_result* out = nullptr;
    uint32_t* cigar = nullptr;
    int64_t cigar_cap = 0;
    std::atomic<int> rc{0};
};
}  // namespace

struct mpn_pool {
    std::vector<mpn_engine*> eng;
    std::vector<std::thread> workers;
    std::mutex mu;
    std::condition_variable cv_job, cv_done;
    PoolJob* job = nullptr;
    uint64_t generation = 0;
    int pending = 0;
    bool stop = false;
    std::vector<double> last_ms;             // host wall time of each device's share of the last batch
    std::vector<int64_t> last_pairs, last_cells;
    std::mutex call_mu;                      // one batch at a time per pool
};

static void pool_worker(mpn_pool* pl, int k)
{
    uint64_t seen = 0;
    for (;;) {
        PoolJob* j = nullptr;
        {
            std::unique_lock<std::mutex> lk(pl->mu);
            pl->cv_job.wait(lk, [&] { return pl->stop || pl->generation != seen; });
            if (pl->stop) return;
            seen = pl->generation;
            j = pl->job;
        }
        const auto t0 = std::chrono::steady_clock::now();
        auto next = [j](RangeJob& r) {
            if (j->rc.load(std::memory_order_relaxed) != 0) return false;
            const size_t i = j->next.fetch_add(1);
            if (i >= j->ranges.size()) return false;
            r = j->ranges[i];
            return true;
        };
        int64_t pairs = 0, cells = 0;
        int rc = j->csr ? run_ranges(pl->eng[k], j->p, *j->csr, j->masklen, next, j->out, j->cigar, j->cigar_cap, &pairs, &cells)
                        : run_ranges(pl->eng[k], j->p, *j->spans, j->masklen, next, j->out, j->cigar, j->cigar_cap, &pairs, &cells);
        if (rc != 0) { int zero = 0; j->rc.compare_exchange_strong(zero, rc); }
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        {
            std::lock_guard<std::mutex> lk(pl->mu);
            pl->last_ms[k] = ms; pl->last_pairs[k] = pairs; pl->last_cells[k] = cells;
            if (--pl->pending == 0) pl->cv_done.notify_all();
        }
    }
}

extern "C" mpn_pool* mpn_pool_create(const int* devices, int ndev)
{
    int have = 0;
    if (cudaGetDeviceCount(&have) != cudaSuccess || have == 0) {
        fprintf(stderr, "[mpn_ssw] no CUDA device available; this library has no CPU fallback\n");
        return nullptr;
    }
    if (ndev <= 0) ndev = have;
    if (!devices && ndev > have) return nullptr;
    mpn_pool* pl = new mpn_pool();
    for (int k = 0; k < ndev; ++k) {
        const int d = devices ? devices[k] : k;
        mpn_engine* e = (d >= 0 && d < have) ? mpn_engine_create(d) : nullptr;
        if (!e) { for (mpn_engine* x : pl->eng) mpn_engine_destroy(x); delete pl; return nullptr; }
        pl->eng.push_back(e);
    }
    pl->last_ms.assign(ndev, 0.0); pl->last_pairs.assign(ndev, 0); pl->last_cells.assign(ndev, 0);
    for (int k = 0; k < ndev; ++k) pl->workers.emplace_back(pool_worker, pl, k);
    return pl;
}

extern "C" void mpn_pool_destroy(mpn_pool* pl)
{
    if (!pl) return;
    { std::lock_guard<std::mutex> lk(pl->mu); pl->stop = true; }
    pl->cv_job.notify_all();
    for (std::thread& t : pl->workers) t.join();
    for (mpn_engine* e : pl->eng) mpn_engine_destroy(e);
    delete pl;
}

extern "C" int mpn_pool_ndev(const mpn_pool* pl) { return pl ? (int)pl->eng.size() : 0; }
extern "C" mpn_engine* mpn_pool_engine(mpn_pool* pl, int k) { return (pl && k >= 0 && k < (int)pl->eng.size()) ? pl->eng[k] : nullptr; }

extern "C" int mpn_pool_last_shares(const mpn_pool* pl, double* ms, int64_t* pairs, int64_t* cells)
{
    if (!pl) return MPN_E_ARG;
    for (size_t k = 0; k < pl->eng.size(); ++k) {
        if (ms) ms[k] = pl->last_ms[k];
        if (pairs) pairs[k] = pl->last_pairs[k];
        if (cells) cells[k] = pl->last_cells[k];
    }
    return 0;
}

// cut [0, npairs) into ranges of about equal cost, in the caller's order; CIGAR regions proportional to the pairs and read bases of a range.
// Costs are summed per block of PLAN_BLOCK pairs on host threads (a 1 M-pair batch is planned in about a millisecond); range boundaries fall
// on block boundaries (4096 pairs, fewer in small batches).
template <class Pairs>
static int plan_ranges(const mpn_pool* pl, const mpn_params* p, const Pairs& all, int64_t npairs, int64_t cigar_cap, bool one_per_device, std::vector<RangeJob>& ranges)
{
    const int ndev = (int)pl->eng.size();
    int maxpos = 0;
    for (int i = 0; i < p->n * p->n; ++i) maxpos = std::max<int>(maxpos, p->mat[i]);
    const int64_t PLAN_BLOCK = std::max<int64_t>(16, std::min<int64_t>(4096, npairs / (64 * (int64_t)ndev)));      // small batches: finer boundaries
    const int64_t nblocks = (npairs + PLAN_BLOCK - 1) / PLAN_BLOCK;
    std::vector<double> bcost((size_t)nblocks, 0.0);
    std::vector<int64_t> bbases((size_t)nblocks, 0);
    std::atomic<int> bad{0};
    const int n = p->n;
    parallel_for(nblocks, 8, [&](int64_t blk) {
        const int64_t lo = blk * PLAN_BLOCK, hi = std::min(npairs, lo + PLAN_BLOCK);
        double c = 0; int64_t bases = 0;
        for (int64_t i = lo; i < hi; ++i) {
            const int64_t rl = all.rl(i), fl = all.fl(i);
            if (rl < 0 || fl < 0) { bad.store(1); continue; }
            // a pair is never free: scheduling, copies and the record cost about as much as a few thousand cells
            c += (double)rl * (double)fl * pair_cost_per_cell(n, maxpos, rl, fl) + 4096.0;
            bases += rl;
        }
        bcost[(size_t)blk] = c; bbases[(size_t)blk] = bases;
    }, 8);
    if (bad.load()) return MPN_E_ARG;
    double total = 0;
    for (double c : bcost) total += c;
    // ranges: enough of them that the queue balances the devices (about 12 per device), small enough for the engine's pipeline
    // (at most 192 k pairs)
    const int64_t max_pairs = 196608;
    int64_t want = one_per_device ? ndev : std::max<int64_t>((int64_t)ndev * 12, (npairs + max_pairs - 1) / max_pairs);
    // ... and a range must be worth a launch sequence: at least ~1.5e9 cost units (a fraction of a millisecond of one GPU)
    want = std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)(total / 1.5e9)));
    if (ndev == 1 && !one_per_device) want = std::max<int64_t>(1, (npairs + max_pairs - 1) / max_pairs);
    want = std::min<int64_t>(want, nblocks);
    const double per = total / (double)want;
    std::vector<double> rcost, rneed;
    int64_t first_blk = 0; double acc = 0; int64_t bases = 0;
    for (int64_t blk = 0; blk < nblocks; ++blk) {
        acc += bcost[(size_t)blk]; bases += bbases[(size_t)blk];
        const int64_t end_pair = std::min(npairs, (blk + 1) * PLAN_BLOCK), first_pair = first_blk * PLAN_BLOCK;
        const bool last = blk + 1 == nblocks;
        if (last || (acc >= per && (int64_t)ranges.size() + 1 < want) || (!one_per_device && end_pair + PLAN_BLOCK - first_pair > max_pairs)) {
            RangeJob r; r.first = first_pair; r.count = end_pair - first_pair;
            ranges.push_back(r); rcost.push_back(acc);
            rneed.push_back(24.0 * (double)r.count + (double)bases / 4.0 + 1024.0);
            first_blk = blk + 1; acc = 0; bases = 0;
        }
    }
    // CIGAR regions
    if ((p->flag & 7) != 0 && cigar_cap > 0) {
        double need_total = 0;
        for (double v : rneed) need_total += v;
        int64_t at = 0;
        for (size_t k = 0; k < ranges.size(); ++k) {
            const int64_t cap = (int64_t)((double)cigar_cap * (rneed[k] / need_total));
            ranges[k].cig_base = at; ranges[k].cig_cap = cap; at += cap;
        }
    } else {
        for (RangeJob& r : ranges) { r.cig_base = 0; r.cig_cap = 0; }
    }
    // heaviest first
    std::vector<size_t> order(ranges.size());
    for (size_t k = 0; k < order.size(); ++k) order[k] = k;
    std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return rcost[a] > rcost[b]; });
    std::vector<RangeJob> sorted;
    sorted.reserve(ranges.size());
    for (size_t k : order) sorted.push_back(ranges[k]);
    ranges.swap(sorted);
    return 0;
}

static int pool_run(mpn_pool* pl, PoolJob& job)
{
    std::lock_guard<std::mutex> call(pl->call_mu);
    {
        std::lock_guard<std::mutex> lk(pl->mu);
        pl->job = &job; pl->pending = (int)pl->eng.size(); ++pl->generation;
        std::fill(pl->last_ms.begin(), pl->last_ms.end(), 0.0);
    }
    pl->cv_job.notify_all();
    {
        std::unique_lock<std::mutex> lk(pl->mu);
        pl->cv_done.wait(lk, [&] { return pl->pending == 0; });
        pl->job = nullptr;
    }
    return job.rc.load();
}

extern "C" int mpn_pool_align_batch(mpn_pool* pl, const mpn_params* p, const int8_t* reads, const int64_t* read_off, const int8_t* refs,
                                    const int64_t* ref_off, const int32_t* masklen, int64_t npairs, mpn_result* out, uint32_t* cigar, int64_t cigar_cap)
{
    if (!pl || !p || !p->mat || p->n < 1 || p->n > 127 || npairs < 0) return MPN_E_ARG;
    if (npairs == 0) return 0;
    const CsrPairs all{reads, read_off, refs, ref_off, npairs};
    if (!all.valid() || !masklen || !out) return MPN_E_ARG;
    PoolJob job;
    job.p = p; job.csr = &all; job.masklen = masklen; job.out = out; job.cigar = cigar; job.cigar_cap = cigar_cap;
    const int rc = plan_ranges(pl, p, all, npairs, cigar ? cigar_cap : 0, false, job.ranges);
    if (rc != 0) return rc;
    return pool_run(pl, job);
}

extern "C" int mpn_pool_align_batch_spans(mpn_pool* pl, const mpn_params* p, const int8_t* seq, int64_t seq_bytes, const int64_t* rd_start, const int32_t* rd_len,
                                          const int64_t* rf_start, const int32_t* rf_len, const int32_t* masklen, int64_t npairs,
                                          mpn_result* out, uint32_t* cigar, int64_t cigar_cap)
{
    if (!pl || !p || !p->mat || p->n < 1 || p->n > 127 || npairs < 0) return MPN_E_ARG;
    if (npairs == 0) return 0;
    const SpanPairs all{seq, seq_bytes, rd_start, rd_len, rf_start, rf_len, npairs};
    if (!all.valid() || !masklen || !out) return MPN_E_ARG;
    PoolJob job;
    job.p = p; job.spans = &all; job.masklen = masklen; job.out = out; job.cigar = cigar; job.cigar_cap = cigar_cap;
    // every range uploads the whole arena: one range per device
    const int rc = plan_ranges(pl, p, all, npairs, cigar ? cigar_cap : 0, true, job.ranges);
    if (rc != 0) return rc;
    return pool_run(pl, job);
}
