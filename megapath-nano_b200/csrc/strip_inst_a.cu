// packed score kernel, reads up to 208 rows: groups of 4 threads (32 / 64 rows) and of 8 threads (96 .. 208 rows in steps of 16)
#include "sw_strip16.cuh"
#include "strip_table.h"
namespace mpn {
const StripEntry g_strip_part_a[] = {
    MPN_STRIP_ENTRY(4, 4), MPN_STRIP_ENTRY(8, 4),
    MPN_STRIP_ENTRY(6, 8), MPN_STRIP_ENTRY(7, 8), MPN_STRIP_ENTRY(8, 8), MPN_STRIP_ENTRY(9, 8), MPN_STRIP_ENTRY(10, 8),
    MPN_STRIP_ENTRY(11, 8), MPN_STRIP_ENTRY(12, 8), MPN_STRIP_ENTRY(13, 8),
};
const StripEntry g_strip_n_a = MPN_STRIP_N_ENTRY(8, 4);
const StripEntry g_strip_long = MPN_STRIP_LONG_ENTRY(16);
const int g_strip_part_a_n = sizeof(g_strip_part_a) / sizeof(g_strip_part_a[0]);
}
