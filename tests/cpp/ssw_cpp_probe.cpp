// Test helper (GPU tier): drives the C++ front end of include/ssw_cpp.h the way realigner.cpp drives the reference's class --
// SetReferenceSequence + Align per query -- and through the batched AlignBatch, and prints one line per query for the Python test.
#include "ssw_cpp.h"
#include <cstdio>
#include <fstream>
#include <iostream>
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
    std::vector<StripedSmithWaterman::Alignment> batch;
    if (!aligner.AlignBatch(queries, filter, &batch)) return 3;
    for (size_t i = 0; i < queries.size(); ++i) {
        StripedSmithWaterman::Alignment a;
        if (!aligner.Align(queries[i].c_str(), filter, &a)) return 4;
        const StripedSmithWaterman::Alignment& b = batch[i];
        const bool same = a.sw_score == b.sw_score && a.sw_score_next_best == b.sw_score_next_best && a.ref_begin == b.ref_begin && a.ref_end == b.ref_end &&
                          a.query_begin == b.query_begin && a.query_end == b.query_end && a.ref_end_next_best == b.ref_end_next_best &&
                          a.mismatches == b.mismatches && a.cigar_string == b.cigar_string && a.cigar == b.cigar;
        printf("%d %d %d %d %d %d %d %d %d %s\n", (int)same, a.sw_score, a.sw_score_next_best, a.ref_begin, a.ref_end, a.query_begin, a.query_end, a.ref_end_next_best,
               a.mismatches, a.cigar_string.c_str());
    }
    // the two-sequence overload (ssw_cpp.cpp:362-403)
    StripedSmithWaterman::Alignment c;
    if (!aligner.Align(queries[0].c_str(), ref.c_str(), (int)ref.size(), filter, &c)) return 5;
    printf("overload %d %s\n", c.sw_score, c.cigar_string.c_str());
    return 0;
}
