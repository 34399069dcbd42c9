// packed score kernel, groups of 8 threads: 224 .. 320 rows in steps of 16
#include "sw_strip16.cuh"
#include "strip_table.h"
namespace mpn {
const StripEntry g_strip_part_b[] = {
    MPN_STRIP_ENTRY(14, 8), MPN_STRIP_ENTRY(15, 8), MPN_STRIP_ENTRY(16, 8), MPN_STRIP_ENTRY(17, 8), MPN_STRIP_ENTRY(18, 8),
    MPN_STRIP_ENTRY(19, 8), MPN_STRIP_ENTRY(20, 8),
};
const StripEntry g_strip_n_b = MPN_STRIP_N_ENTRY(20, 8);
const int g_strip_part_b_n = sizeof(g_strip_part_b) / sizeof(g_strip_part_b[0]);
}
