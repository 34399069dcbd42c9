/*
 * ssw_cpp.h -- C++ front end of the B200 Smith-Waterman engine, source-compatible with the reference's
 * bin/realignment/realign/ssw_cpp.h (namespace StripedSmithWaterman: Alignment :10-35, Filter :37-68, Aligner :70-194).
 *
 * Same public names, argument lists and results as the reference class, so realigner.cpp-style callers compile unchanged;
 * what differs is underneath: Align() is one GPU batch of one pair through include/mpn_ssw_batch.h, and the new
 * AlignBatch() / AlignPairs() submit any number of pairs in one go (the reason this engine exists).
 * No CPU alignment code is behind this header.
 */
#ifndef MPN_SSW_CPP_H_
#define MPN_SSW_CPP_H_

#include <stdint.h>
#include <string>
#include <vector>

namespace StripedSmithWaterman {

/* result of one alignment (reference ssw_cpp.h:10-35) */
struct Alignment {
  uint16_t sw_score = 0;
  uint16_t sw_score_next_best = 0;
  int32_t ref_begin = 0;
  int32_t ref_end = 0;
  int32_t query_begin = 0;
  int32_t query_end = 0;
  int32_t ref_end_next_best = 0;
  int32_t mismatches = 0;            /* mismatching bases + inserted + deleted bases (ssw_cpp.cpp:123-207) */
  std::string cigar_string;          /* soft clips and '=' / 'X' runs included */
  std::vector<uint32_t> cigar;       /* BAM words, length << 4 | index into "MIDNSHP=X" */

  void Clear() {
    sw_score = sw_score_next_best = 0;
    ref_begin = ref_end = query_begin = query_end = ref_end_next_best = mismatches = 0;
    cigar_string.clear();
    cigar.clear();
  }
};

/* what to report (reference ssw_cpp.h:37-68); the defaults ask for everything */
struct Filter {
  bool report_begin_position = true;
  bool report_cigar = true;          /* implies report_begin_position */
  uint16_t score_filter = 0;         /* CIGAR only if score >= score_filter */
  uint16_t distance_filter = 32767;  /* CIGAR only if both spans are below it */

  Filter() {}
  Filter(const bool& pos, const bool& cigar, const uint16_t& score, const uint16_t& dis)
      : report_begin_position(pos), report_cigar(cigar), score_filter(score), distance_filter(dis) {}
};

/* one (query, target) pair of a batch; the pointers are only read during the call */
struct PairView {
  const char* query; int query_len;
  const char* ref;   int ref_len;
};

/* AlignIndexed: a pool of distinct sequences plus pairs that name their query and target by index into the pool */
struct SeqView { const char* text; int len; };
struct PairIndex { int32_t query; int32_t target; };
/* AlignIndexedCompact: what the region realigner consumes of an Alignment (realigner.cpp:336-348, :369-379), without one heap string and one
 * vector per pair: sw_score, ref_begin and the '=' / 'X' / 'I' / 'D' / 'S' CIGAR text as (pointer, length) into buffers owned by the result */
struct CompactAlignments {
  std::vector<int32_t> sw_score, ref_begin, mismatches;
  std::vector<const char*> cigar;       /* not NUL terminated */
  std::vector<int32_t> cigar_len;
  std::vector<std::string> text;        /* the buffers cigar[] points into (one per block of pairs) */
  size_t size() const { return sw_score.size(); }
};

class Aligner {
 public:
  Aligner(void);                                                       /* {A,C,G,T,N}, +4/-6, gaps 8/2 (ssw_cpp.cpp:218-230 with realigner.cpp's values) */
  Aligner(const uint8_t& match_score, const uint8_t& mismatch_penalty,
          const uint8_t& gap_opening_penalty, const uint8_t& gap_extending_penalty);
  Aligner(const int8_t* score_matrix, const int& score_matrix_size,
          const int8_t* translation_matrix, const int& translation_matrix_size);
  ~Aligner(void);

  int SetReferenceSequence(const char* seq, const int& length);        /* ssw_cpp.cpp:286-309 */
  void CleanReferenceSequence(void);
  void SetGapPenalty(const uint8_t& opening, const uint8_t& extending) { gap_open_ = opening; gap_extend_ = extending; }

  /* per-pair calls of the reference (ssw_cpp.cpp:326-359, 362-403) */
  bool Align(const char* query, const Filter& filter, Alignment* alignment) const;
  bool Align(const char* query, const char* ref, const int& ref_len, const Filter& filter, Alignment* alignment) const;

  /* NEW: every query against the stored reference, one GPU submit; out[i] is what Align(queries[i], ...) returns */
  bool AlignBatch(const std::vector<std::string>& queries, const Filter& filter, std::vector<Alignment>* out) const;
  /* NEW: arbitrary (query, target) pairs, one GPU submit.  Empty queries give a cleared Alignment (Align would return false). */
  bool AlignPairs(const std::vector<PairView>& pairs, const Filter& filter, std::vector<Alignment>* out) const;
  /* NEW: the same for callers that already know which pairs share sequences (a haplotype met by hundreds of reads): every
   * sequence of `pool` is translated and uploaded once; pairs with an empty query or target give a cleared Alignment. */
  bool AlignIndexed(const std::vector<SeqView>& pool, const std::vector<PairIndex>& pairs, const Filter& filter, std::vector<Alignment>* out) const;
  /* NEW: same call, compact result (see CompactAlignments); a pair with an empty query has sw_score 0, ref_begin 0 and an empty CIGAR, like a cleared Alignment */
  bool AlignIndexedCompact(const std::vector<SeqView>& pool, const std::vector<PairIndex>& pairs, const Filter& filter, CompactAlignments* out) const;

  void Clear(void);
  bool ReBuild(void);
  bool ReBuild(const uint8_t& match_score, const uint8_t& mismatch_penalty,
               const uint8_t& gap_opening_penalty, const uint8_t& gap_extending_penalty);
  bool ReBuild(const int8_t* score_matrix, const int& score_matrix_size,
               const int8_t* translation_matrix, const int& translation_matrix_size);

 private:
  struct IndexedRun;
  bool RunIndexed(const std::vector<SeqView>& pool, const std::vector<PairIndex>& pairs, const Filter& filter, IndexedRun* run) const;
  std::vector<int8_t> matrix_;       /* n x n */
  int n_ = 5;
  std::vector<int8_t> translate_;    /* ASCII -> code; empty = aligner disabled */
  uint8_t match_ = 4, mismatch_ = 6, gap_open_ = 8, gap_extend_ = 2;
  std::vector<int8_t> reference_;    /* translated */

  void default_tables();
  Aligner& operator=(const Aligner&);
  Aligner(const Aligner&);
};

}  // namespace StripedSmithWaterman

#endif  // MPN_SSW_CPP_H_
