// packed score kernel, one warp per pair: 704 .. 1280 rows in steps of 64
#include "sw_strip16.cuh"
#include "strip_table.h"
namespace mpn {
const StripEntry g_strip_part_d[] = {
    MPN_STRIP_ENTRY(11, 32), MPN_STRIP_ENTRY(12, 32), MPN_STRIP_ENTRY(13, 32), MPN_STRIP_ENTRY(14, 32), MPN_STRIP_ENTRY(15, 32),
    MPN_STRIP_ENTRY(16, 32), MPN_STRIP_ENTRY(17, 32), MPN_STRIP_ENTRY(18, 32), MPN_STRIP_ENTRY(19, 32), MPN_STRIP_ENTRY(20, 32),
};
const StripEntry g_strip_n_d = MPN_STRIP_N_ENTRY(20, 32);
const int g_strip_part_d_n = sizeof(g_strip_part_d) / sizeof(g_strip_part_d[0]);
}
