// Test helper: the part of StripedSmithWaterman::Aligner that the reference and the product share (ssw_cpp.h:60-120 of the reference) --
// SetReferenceSequence + Align per query, and the two-sequence overload -- printing every Alignment field.  Compiled twice by the tests:
// with the reference's ssw_cpp.cpp + ssw.c (all CPU) and with include/ssw_cpp.h + the product library; the outputs must be identical,
// which pins ConvertAlignment / CalculateNumberMismatch (ssw_cpp.cpp:50-86, 123-207) including the `mismatches` field.
#include "ssw_cpp.h"
#include <cstdio>
#include <fstream>
#include <string>
#include <vector>

int main(int argc, char** argv)
{
    if (argc < 2) return 2;
    std::ifstream in(argv[1]);
    std::string ref;
    std::getline(in, ref);
    std::vector<std::string> queries;
    for (std::string q; std::getline(in, q);) if (!q.empty()) queries.push_back(q);
    StripedSmithWaterman::Aligner aligner(4, 6, 8, 2);
    StripedSmithWaterman::Filter filter;
    aligner.SetReferenceSequence(ref.c_str(), (int)ref.size());
    for (size_t i = 0; i < queries.size(); ++i) {
        StripedSmithWaterman::Alignment a;
        if (!aligner.Align(queries[i].c_str(), filter, &a)) return 4;
        printf("%d %d %d %d %d %d %d %d %s", (int)a.sw_score, (int)a.sw_score_next_best, a.ref_begin, a.ref_end, a.query_begin, a.query_end, a.ref_end_next_best,
               a.mismatches, a.cigar_string.c_str());
        for (size_t k = 0; k < a.cigar.size(); ++k) printf(" %u", a.cigar[k]);
        printf("\n");
    }
    for (size_t i = 0; i < queries.size() && i < 4; ++i) {      // the two-sequence overload (ssw_cpp.cpp:362-403)
        StripedSmithWaterman::Alignment c;
        if (!aligner.Align(queries[i].c_str(), ref.c_str(), (int)ref.size(), filter, &c)) return 5;
        printf("overload %d %d %d %d %s\n", (int)c.sw_score, c.ref_begin, c.query_begin, c.mismatches, c.cigar_string.c_str());
    }
    return 0;
}
