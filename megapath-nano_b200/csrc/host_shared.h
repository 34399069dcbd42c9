// Process-wide engine shared by the per-pair ABI (ssw_abi.cu), the C++ front end (ssw_cpp_layer.cpp) and the region
// realigner (realign_region.cpp).  One CUDA context, one set of pooled buffers per process; calls serialise on a mutex
// (the reference's library is re-entrant, SURVEY.md section 8b "Threading", so concurrent callers must be tolerated).
#pragma once
#include "../../include/mpn_ssw_batch.h"
#include <mutex>

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

}  // namespace mpn
