// packed score kernel, groups of 16 threads: 352 .. 640 rows in steps of 32
#include "sw_strip16.cuh"
#include "strip_table.h"
namespace mpn {
const StripEntry g_strip_part_c[] = {
    MPN_STRIP_ENTRY(11, 16), MPN_STRIP_ENTRY(12, 16), MPN_STRIP_ENTRY(13, 16), MPN_STRIP_ENTRY(14, 16), MPN_STRIP_ENTRY(15, 16),
    MPN_STRIP_ENTRY(16, 16), MPN_STRIP_ENTRY(17, 16), MPN_STRIP_ENTRY(18, 16), MPN_STRIP_ENTRY(19, 16), MPN_STRIP_ENTRY(20, 16),
};
const StripEntry g_strip_n_c = MPN_STRIP_N_ENTRY(20, 16);
const int g_strip_part_c_n = sizeof(g_strip_part_c) / sizeof(g_strip_part_c[0]);
}
