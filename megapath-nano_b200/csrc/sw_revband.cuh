// Banded reverse pass (sm_100a): begin positions of ssw_align (ssw.c:820-832) for the pairs whose score leaves little room for gaps.
// The per-lane routine, the band bound and why it is exact are in sw_revband_core.h.  Here: the kernel that sorts the reverse tasks
// of the packed short-read bins into one queue per band class, and the kernel that walks a queue with one pair per lane.
// Pairs that do not qualify (band wider than 80 diagonals, an N the score table cannot express) are put on the list the N variants of
// sw_strip16_kernel walk, flagged SW_FLAG_NEEDS_WIDE: they get the full-matrix reverse pass.
#pragma once
#include "sw_common.cuh"
#include "sw_revband_core.h"

namespace mpn {

constexpr int REVBAND_CLASSES = 5;                 // NW = 4, 8, 12, 16, 20 registers per anti-diagonal: bands of up to 16 / 32 / 48 / 64 / 80 diagonals
constexpr int REVBAND_BLOCK = 128;

struct RevBandQueues {
    int* items;          // REVBAND_CLASSES x capacity task indices
    int* count;          // [REVBAND_CLASSES] filled by the setup kernel
    int* cursor;         // [REVBAND_CLASSES] next batch of 32 to take
    int capacity;
};

// one thread per reverse task of the packed bins [0, ntasks)
__global__ void __launch_bounds__(128)
sw_revband_setup_kernel(const SwTask* __restrict__ rev_tasks, int ntasks, rb::Score sc, RevBandQueues q, SwEnds* __restrict__ ends, int* __restrict__ relist)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = k < ntasks;                  // whole warps stay: the appends are warp-aggregated
    int dest = -1;                                 // -1 nothing to do, 0 .. REVBAND_CLASSES - 1 band class, REVBAND_CLASSES full-matrix pass
    int out = 0;
    if (live) {
        const SwTask tk = rev_tasks[k];
        out = tk.out;
        if (tk.rd_len > 0 && tk.rf_len > 0) {
            int h0;
            const int nw = rb::classify(tk.rd_len, tk.rf_len, tk.stop, sc, h0);
            dest = nw == 0 ? REVBAND_CLASSES : nw / 4 - 1;
        }
    }
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 0; d <= REVBAND_CLASSES; ++d) {
        const unsigned m = __ballot_sync(0xffffffffu, dest == d);
        if (m == 0u) continue;
        int base = 0;
        if (lane == __ffs((int)m) - 1) base = atomicAdd(d < REVBAND_CLASSES ? q.count + d : relist, __popc(m));
        base = __shfl_sync(0xffffffffu, base, __ffs((int)m) - 1);
        if (dest == d) {
            const int at = base + __popc(m & ((1u << lane) - 1u));
            if (d < REVBAND_CLASSES) q.items[(size_t)d * q.capacity + at] = k;
            else {
                relist[1 + at] = k;
                SwEnds e; e.score = 0; e.col = -1; e.row = 0; e.flags = SW_FLAG_NEEDS_WIDE;
                ends[out] = e;
            }
        }
    }
}

// one pair per lane; a warp takes 32 consecutive queue items (queue order is about task order, i.e. read-length bin: similar trip counts)
template <int NW>
__device__ __forceinline__ void revband_body(const SwTask* __restrict__ rev_tasks, const int8_t* __restrict__ seq, const rb::Score& sc, const RevBandQueues& q,
                                             SwEnds* __restrict__ ends, int* __restrict__ relist)
{
    constexpr int cls = NW / 4 - 1;
    const int lane = threadIdx.x & 31;
    const int n = q.count[cls];
    const int* items = q.items + (size_t)cls * q.capacity;
    for (;;) {
        int base = 0;
        if (lane == 0) base = atomicAdd(q.cursor + cls, 32);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= n) break;
        const int at = base + lane;
        if (at < n) {
            const int k = items[at];
            const SwTask tk = rev_tasks[k];
            int h0;
            rb::classify(tk.rd_len, tk.rf_len, tk.stop, sc, h0);
            int col = 0, row = 0;
            const int rc = rb::lane<NW>(seq, tk.rd_base, tk.rf_base, tk.rd_len, tk.rf_len, tk.stop, h0, sc, col, row);
            SwEnds e;
            if (rc == 0) { e.score = tk.stop; e.col = col; e.row = row; e.flags = 0; }
            else {
                e.score = 0; e.col = -1; e.row = 0; e.flags = SW_FLAG_NEEDS_WIDE;
                relist[1 + atomicAdd(relist, 1)] = k;
            }
            ends[tk.out] = e;
        }
        __syncwarp();
    }
}

template <int NW>
__global__ void __launch_bounds__(REVBAND_BLOCK)
sw_revband_kernel(const SwTask* __restrict__ rev_tasks, const int8_t* __restrict__ seq, const rb::Score sc, RevBandQueues q, SwEnds* __restrict__ ends, int* __restrict__ relist)
{
    revband_body<NW>(rev_tasks, seq, sc, q, ends, relist);
}

// all classes in one launch (blockIdx.y = 0 is the widest): for the small ranges of the chunk pipeline, where a launch per class would cost
// five times the latency of one pair's DP and none of them could fill the GPU; every class then runs at the widest one's register budget
__global__ void __launch_bounds__(REVBAND_BLOCK)
sw_revband_all_kernel(const SwTask* __restrict__ rev_tasks, const int8_t* __restrict__ seq, const rb::Score sc, RevBandQueues q, SwEnds* __restrict__ ends, int* __restrict__ relist)
{
    switch (blockIdx.y) {
        case 0: revband_body<20>(rev_tasks, seq, sc, q, ends, relist); break;
        case 1: revband_body<16>(rev_tasks, seq, sc, q, ends, relist); break;
        case 2: revband_body<12>(rev_tasks, seq, sc, q, ends, relist); break;
        case 3: revband_body<8>(rev_tasks, seq, sc, q, ends, relist); break;
        default: revband_body<4>(rev_tasks, seq, sc, q, ends, relist); break;
    }
}

}  // namespace mpn
