// Process-wide engine shared by the per-pair ABI (ssw_abi.cu), the C++ front end (ssw_cpp_layer.cpp) and the region
// realigner (realign_region.cpp).  One CUDA context, one set of pooled buffers per process; calls serialise on a mutex
// (the reference's library is re-entrant, SURVEY.md section 8b "Threading", so concurrent callers must be tolerated).
#pragma once
#include "../../include/mpn_ssw_batch.h"
#include <algorithm>
#include <atomic>
#include <cstdint>
#include <mutex>
#include <thread>
#include <vector>

namespace mpn {

std::mutex& shared_engine_mutex();
mpn_engine* shared_engine_locked();      // creates the engine on first use; aborts (loudly) if there is no GPU

class SharedEngineLock {
public:
    SharedEngineLock() : lk_(shared_engine_mutex()), e_(shared_engine_locked()) {}
    mpn_engine* engine() const { return e_; }
private:
    std::lock_guard<std::mutex> lk_;
    mpn_engine* e_;
};


// fn(i) for i in [0, n) on up to `max_threads` host threads (dynamic chunks); runs inline when the range is small
template <class F>
inline void parallel_for(int64_t n, int64_t grain, F&& fn, unsigned max_threads = 0)
{
    unsigned hw = std::thread::hardware_concurrency();
    if (hw == 0) hw = 4;
    if (max_threads) hw = std::min(hw, max_threads);
    const int64_t chunks = (n + grain - 1) / std::max<int64_t>(grain, 1);
    const unsigned nt = (unsigned)std::min<int64_t>(hw, chunks);
    if (nt <= 1) { for (int64_t i = 0; i < n; ++i) fn(i); return; }
    std::atomic<int64_t> next(0);
    auto work = [&]() {
        for (;;) {
            const int64_t c = next.fetch_add(1);
            if (c >= chunks) break;
            const int64_t lo = c * grain, hi = std::min(n, lo + grain);
            for (int64_t i = lo; i < hi; ++i) fn(i);
        }
    };
    std::vector<std::thread> th;
    th.reserve(nt - 1);
    for (unsigned t = 1; t < nt; ++t) th.emplace_back(work);
    work();
    for (std::thread& t : th) t.join();
}

}  // namespace mpn
