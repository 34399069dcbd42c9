// Table of the packed score-kernel instantiations (sw_strip16.cuh).  The template is instantiated in four translation
// units (strip_inst_*.cu) only so that they compile in parallel; engine.cu merges the parts into one length-binned table.
#pragma once
#include "sw_common.cuh"
#include <cstddef>

namespace mpn {

typedef void (*StripFn)(const SwTask*, int, int*, const int8_t*, const Score16, uint32_t*, SwEnds*, int*, int, uint32_t*, long long);
struct StripEntry { int G, KR; StripFn fn; StripFn fn_rev; size_t smem, smem_rev; };      // forward / reverse instantiation of the same strip and their dynamic shared memory

extern const StripEntry g_strip_part_a[]; extern const int g_strip_part_a_n;
extern const StripEntry g_strip_part_b[]; extern const int g_strip_part_b_n;
extern const StripEntry g_strip_part_c[]; extern const int g_strip_part_c_n;
extern const StripEntry g_strip_part_d[]; extern const int g_strip_part_d_n;

#define MPN_STRIP_ENTRY(KR, G) { G, KR, sw_strip16_kernel<KR, G, false>, sw_strip16_kernel<KR, G, true>, strip16_smem_bytes<KR, G, false, false>(), strip16_smem_bytes<KR, G, false, true>() }
// the N variants (sw_strip16.cuh, NM = true): one strip per group width, its largest; each translation unit holds one
extern const StripEntry g_strip_n_a, g_strip_n_b, g_strip_n_c, g_strip_n_d;
// the multi-strip instantiation with the int16 clamp (LONG = true): reads of any length, one warp per pair
extern const StripEntry g_strip_long;
#define MPN_STRIP_LONG_ENTRY(KR) { 32, KR, sw_strip16_kernel<KR, 32, false, false, true>, sw_strip16_kernel<KR, 32, true, false, true>, strip16_smem_bytes<KR, 32, false, false, true>(), strip16_smem_bytes<KR, 32, false, true, true>() }
#define MPN_STRIP_N_ENTRY(KR, G) { G, KR, sw_strip16_kernel<KR, G, false, true>, sw_strip16_kernel<KR, G, true, true>, strip16_smem_bytes<KR, G, true, false>(), strip16_smem_bytes<KR, G, true, true>() }

}  // namespace mpn
