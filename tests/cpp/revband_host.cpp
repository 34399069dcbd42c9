// Host build of the per-lane routine of the banded reverse pass (megapath-nano_b200/csrc/sw_revband_core.h), for
// tests/test_revband_core.py: the same source the kernel runs, packed arithmetic emulated, checked against the compiled reference.
// Test infrastructure only.
#include "../../megapath-nano_b200/csrc/sw_revband_core.h"

using namespace mpn::rb;

extern "C" int revband_host(const int8_t* seq, int64_t rd_base, int64_t rf_base, int L, int C, int S, int mt, int mm, int gapO, int gapE,
                            int n_is_mismatch, int force_nw, int* col, int* row, int* nw_out, int* h0_out)
{
    Score sc;
    const uint32_t m8 = (uint32_t)(uint8_t)(int8_t)mm;
    sc.tlo = (uint32_t)(uint8_t)(int8_t)mt | (m8 << 8) | (m8 << 16) | (m8 << 24);
    sc.thi = m8 * 0x01010101u;
    sc.mgo2 = pack16(-gapO, -gapO);
    sc.mge2 = pack16(-gapE, -gapE);
    sc.mt = mt; sc.gapO = gapO; sc.gapE = gapE; sc.n_is_mismatch = n_is_mismatch;
    int h0 = 0;
    int nw = classify(L, C, S, sc, h0);
    if (nw != 0 && force_nw > nw) nw = force_nw;       // a wider class than needed must give the same answer
    *nw_out = nw; *h0_out = h0;
    switch (nw) {
        case 4: return lane<4>(seq, rd_base, rf_base, L, C, S, h0, sc, *col, *row);
        case 8: return lane<8>(seq, rd_base, rf_base, L, C, S, h0, sc, *col, *row);
        case 12: return lane<12>(seq, rd_base, rf_base, L, C, S, h0, sc, *col, *row);
        case 16: return lane<16>(seq, rd_base, rf_base, L, C, S, h0, sc, *col, *row);
        case 20: return lane<20>(seq, rd_base, rf_base, L, C, S, h0, sc, *col, *row);
        default: return 2;
    }
}
