// C++ front end (include/ssw_cpp.h) over the batched GPU engine.  Mirrors StripedSmithWaterman::Aligner of the reference
// (ssw_cpp.cpp:214-477): translation, the flag/filter mapping of SetFlag (:209-212), and the post-processing that turns the
// raw M/I/D words of ssw_align into the soft-clipped '=' / 'X' CIGAR plus a mismatch count (ConvertAlignment :50-86,
// CalculateNumberMismatch :123-207).  Host code only: every DP cell is computed by the CUDA kernels behind mpn_align_batch.
#include "../../include/ssw_cpp.h"
#include "../../include/mpn_ssw_batch.h"
#include "host_shared.h"

#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <cstring>
#include <functional>
#include <memory>
#include <unordered_map>

namespace StripedSmithWaterman {

namespace {

// ASCII -> {A0 C1 G2 T3 N4}; U/u read as A exactly like the reference table (ssw_cpp.cpp:8-21)
void fill_dna_translation(std::vector<int8_t>& t)
{
    t.assign(128, 4);
    const char* letters = "AaCcGgTtUu";
    const int8_t codes[] = {0, 0, 1, 1, 2, 2, 3, 3, 0, 0};
    for (int k = 0; letters[k]; ++k) t[(unsigned char)letters[k]] = codes[k];
}

void fill_dna_matrix(std::vector<int8_t>& m, uint8_t match, uint8_t mismatch)
{
    // 5 x 5: +match on the ACGT diagonal, -mismatch everywhere else including the whole N row and column (ssw_cpp.cpp:23-48)
    m.assign(25, (int8_t)(-(int)mismatch));
    for (int i = 0; i < 4; ++i) m[i * 5 + i] = (int8_t)match;
}

inline void append_run(std::string& s, std::vector<uint32_t>& words, uint32_t len, char op, uint32_t code)
{
    char buf[16];                                   // decimal digits back to front, then the operation letter (no snprintf on this path:
    int at = 15;                                    // it runs once per CIGAR run of every pair)
    buf[at] = op;
    uint32_t v = len;
    do { buf[--at] = (char)('0' + v % 10u); v /= 10u; } while (v);
    s.append(buf + at, (size_t)(16 - at));
    words.push_back((len << 4) | code);
}

}  // namespace

// raw engine record -> Alignment, identical to ConvertAlignment followed by CalculateNumberMismatch
void finish_alignment(const mpn_result& r, const uint32_t* cigar_arena, const int8_t* ref, const int8_t* query, int query_len, Alignment* al)
{
    al->Clear();
    al->sw_score = r.score1; al->sw_score_next_best = r.score2;
    al->ref_begin = r.ref_begin1; al->ref_end = r.ref_end1;
    al->query_begin = r.read_begin1; al->query_end = r.read_end1;
    al->ref_end_next_best = r.ref_end2;
    enum { OP_EQ = 7, OP_X = 8, OP_S = 4 };
    std::string& s = al->cigar_string;
    std::vector<uint32_t>& w = al->cigar;
    s.reserve(48);
    w.reserve((size_t)std::max(r.cigar_len, 0) * 3 + 4);
    if (al->query_begin > 0) append_run(s, w, (uint32_t)al->query_begin, 'S', OP_S);
    int mism = 0;
    // With a negative begin (not computed, or score 0) the reference walks from ref[-1]; there is no M run longer than the
    // "1M" placeholder in that case and the comparison it makes is between out-of-bounds bytes: treat that one base as a match.
    const bool walkable = r.ref_begin1 >= 0 && r.read_begin1 >= 0;
    const int8_t* tp = ref + (walkable ? r.ref_begin1 : 0);
    const int8_t* qp = query + (walkable ? r.read_begin1 : 0);
    uint32_t run = 0; bool run_is_x = false;
    auto flush = [&]() { if (run) append_run(s, w, run, run_is_x ? 'X' : '=', run_is_x ? OP_X : OP_EQ); run = 0; };
    for (int k = 0; k < r.cigar_len; ++k) {
        const uint32_t word = cigar_arena[r.cigar_off + k];
        const uint32_t len = word >> 4, op = word & 15u;
        if (op == 0) {
            for (uint32_t j = 0; j < len; ++j) {
                const bool x = walkable ? (*tp != *qp) : false;
                if (run && x != run_is_x) flush();
                run_is_x = x; ++run; mism += x;
                ++tp; ++qp;
            }
        } else if (op == 1) {
            flush(); qp += len; mism += (int)len; append_run(s, w, len, 'I', 1);
        } else if (op == 2) {
            flush(); tp += len; mism += (int)len; append_run(s, w, len, 'D', 2);
        }
    }
    flush();
    const int tail = query_len - al->query_end - 1;
    if (tail > 0) append_run(s, w, (uint32_t)tail, 'S', OP_S);
    al->mismatches = mism;
}

Aligner::Aligner(void) { default_tables(); }

Aligner::Aligner(const uint8_t& match_score, const uint8_t& mismatch_penalty, const uint8_t& gap_opening_penalty, const uint8_t& gap_extending_penalty)
    : match_(match_score), mismatch_(mismatch_penalty), gap_open_(gap_opening_penalty), gap_extend_(gap_extending_penalty)
{
    default_tables();
}

Aligner::Aligner(const int8_t* score_matrix, const int& score_matrix_size, const int8_t* translation_matrix, const int& translation_matrix_size)
    : n_(score_matrix_size)
{
    matrix_.assign(score_matrix, score_matrix + (size_t)n_ * n_);
    translate_.assign(translation_matrix, translation_matrix + translation_matrix_size);
}

Aligner::~Aligner(void) {}

void Aligner::default_tables()
{
    n_ = 5;
    fill_dna_matrix(matrix_, match_, mismatch_);
    fill_dna_translation(translate_);
}

static inline void translate_into(const std::vector<int8_t>& table, const char* s, int len, int8_t* out)
{
    const int tn = (int)table.size();
    for (int i = 0; i < len; ++i) {
        const int c = (unsigned char)s[i];
        out[i] = c < tn ? table[c] : table[tn - 1];
    }
}

int Aligner::SetReferenceSequence(const char* seq, const int& length)
{
    reference_.clear();
    if (translate_.empty() || length <= 0) return 0;
    reference_.resize((size_t)length);
    translate_into(translate_, seq, length, reference_.data());
    return length;
}

void Aligner::CleanReferenceSequence(void) { reference_.clear(); }

bool Aligner::AlignPairs(const std::vector<PairView>& pairs, const Filter& filter, std::vector<Alignment>* out) const
{
    if (translate_.empty() || !out) return false;
    // every DISTINCT sequence (by address and length) becomes one pool entry -- in the realigner one haplotype meets hundreds of
    // reads and one read meets every haplotype (realigner.cpp:351-384)
    struct Key { const void* p; int len; bool operator==(const Key& o) const { return p == o.p && len == o.len; } };
    struct KeyHash { size_t operator()(const Key& k) const { return std::hash<const void*>()(k.p) * 31u + (size_t)k.len; } };
    std::unordered_map<Key, int32_t, KeyHash> where;
    where.reserve(pairs.size());
    std::vector<SeqView> pool;
    auto intern = [&](const char* p, int len) -> int32_t {
        const Key k{p ? (const void*)p : (const void*)reference_.data(), len};
        auto it = where.find(k);
        if (it != where.end()) return it->second;
        const int32_t id = (int32_t)pool.size();
        where.emplace(k, id);
        pool.push_back(SeqView{p, len});             // text == nullptr: the stored (already translated) reference
        return id;
    };
    std::vector<PairIndex> idx(pairs.size());
    for (size_t i = 0; i < pairs.size(); ++i) idx[i] = PairIndex{intern(pairs[i].query, pairs[i].query_len), intern(pairs[i].ref, pairs[i].ref_len)};
    return AlignIndexed(pool, idx, filter, out);
}

// engine part shared by AlignIndexed and AlignIndexedCompact: translate the pool, build the spans of the live pairs, run the batch
struct Aligner::IndexedRun {
    std::unique_ptr<int8_t[]> arena_store;
    int8_t* arena = nullptr;
    std::vector<int64_t> rd_start, rf_start;
    std::vector<int32_t> rd_len, rf_len, mask;
    std::vector<size_t> slot;                    // position of live pair k in the caller's pair list
    std::unique_ptr<mpn_result[]> res;
    std::unique_ptr<uint32_t[]> cig;
    int64_t n = 0;
    double t_start = 0, t_packed = 0, t_gpu = 0;
};

static double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

bool Aligner::RunIndexed(const std::vector<SeqView>& pool, const std::vector<PairIndex>& pairs, const Filter& filter, IndexedRun* run) const
{
    IndexedRun& R = *run;
    const size_t np = pairs.size();
    // ---- arena: the pool back to back, translated on host threads
    std::vector<int64_t> start(pool.size() + 1, 0);
    for (size_t d = 0; d < pool.size(); ++d) start[d + 1] = start[d] + std::max(pool[d].len, 0);
    const int64_t arena_bytes = start[pool.size()];
    R.arena_store.reset(new int8_t[(size_t)arena_bytes + 16]);      // no zero fill: every byte is written below
    int8_t* const arena = R.arena = R.arena_store.get();
    mpn::parallel_for((int64_t)pool.size(), 64, [&](int64_t d) {
        if (pool[d].len <= 0) return;
        if (pool[d].text == nullptr) memcpy(arena + start[d], reference_.data(), (size_t)pool[d].len);
        else translate_into(translate_, pool[d].text, pool[d].len, arena + start[d]);
    });
    // ---- spans of the live pairs (an empty query makes Align return false: the slot stays cleared)
    int64_t qbytes = 0, tbytes = 0;
    {
        // positions of the live pairs first (serial, one compare per pair), then the five span arrays filled on host threads
        R.slot.reserve(np);
        for (size_t i = 0; i < np; ++i) if (pool[(size_t)pairs[i].query].len > 0) R.slot.push_back(i);
        const size_t nl = R.slot.size();
        R.rd_start.resize(nl); R.rf_start.resize(nl); R.rd_len.resize(nl); R.rf_len.resize(nl); R.mask.resize(nl);
        mpn::parallel_for((int64_t)nl, 32768, [&](int64_t k) {
            const size_t i = R.slot[(size_t)k];
            const SeqView& q = pool[(size_t)pairs[i].query]; const SeqView& t = pool[(size_t)pairs[i].target];
            R.rd_start[(size_t)k] = start[(size_t)pairs[i].query]; R.rd_len[(size_t)k] = q.len;
            R.rf_start[(size_t)k] = start[(size_t)pairs[i].target]; R.rf_len[(size_t)k] = std::max(t.len, 0);
            R.mask[(size_t)k] = q.len;                      // maskLen = query_len (ssw_cpp.cpp:346)
        });
        for (size_t k = 0; k < nl; ++k) { qbytes += R.rd_len[k]; tbytes += R.rf_len[k]; }
    }
    const int64_t n = R.n = (int64_t)R.slot.size();
    R.t_packed = R.t_gpu = now_ms();
    if (n == 0) return true;
    uint8_t flag = 0;                                      // SetFlag, ssw_cpp.cpp:209-212
    if (filter.report_begin_position) flag |= 0x08;
    if (filter.report_cigar) flag |= 0x0f;
    mpn_params pr;
    pr.mat = matrix_.data(); pr.n = n_; pr.gapO = gap_open_; pr.gapE = gap_extend_; pr.score_size = 2;
    pr.flag = flag; pr.filters = filter.score_filter; pr.filterd = filter.distance_filter;
    R.res.reset(new mpn_result[(size_t)n]);
    size_t cig_cap = (size_t)(n * 24 + qbytes / 4 + 4096);
    R.cig.reset(new uint32_t[cig_cap]);
    int rc;
    for (int attempt = 0;; ++attempt) {
        mpn::SharedEngineLock lk;
        rc = mpn_align_batch_spans(lk.engine(), &pr, arena, arena_bytes, R.rd_start.data(), R.rd_len.data(), R.rf_start.data(), R.rf_len.data(), R.mask.data(), n,
                                   R.res.get(), R.cig.get(), (int64_t)cig_cap);
        if (rc != MPN_E_CIGAR_SPACE || attempt == 1) break;
        cig_cap = (size_t)(2 * (qbytes + tbytes) + 16 * n);             // always enough: a CIGAR has at most read + target runs
        R.cig.reset(new uint32_t[cig_cap]);
    }
    if (rc != 0) {
        fprintf(stderr, "[ssw_cpp] GPU alignment failed (code %d); this library has no CPU fallback\n", rc);
        abort();
    }
    for (int64_t k = 0; k < n; ++k)
        if (R.res[k].status != MPN_ST_OK) {                // the reference dereferences a NULL s_align here; fail loudly instead
            fprintf(stderr, "[ssw_cpp] ssw_align returned no result for pair %lld\n", (long long)R.slot[k]);
            abort();
        }
    R.t_gpu = now_ms();
    return true;
}

bool Aligner::AlignIndexed(const std::vector<SeqView>& pool, const std::vector<PairIndex>& pairs, const Filter& filter, std::vector<Alignment>* out) const
{
    if (translate_.empty() || !out) return false;
    static const bool timing = getenv("MPN_TIMING") != nullptr;
    IndexedRun R;
    R.t_start = now_ms();
    out->clear();
    out->resize(pairs.size());
    if (!RunIndexed(pool, pairs, filter, &R)) return false;
    // ---- '=' / 'X' CIGAR + mismatch count per pair (independent: host threads)
    mpn::parallel_for(R.n, 256, [&](int64_t k) {
        finish_alignment(R.res[k], R.cig.get(), R.arena + R.rf_start[k], R.arena + R.rd_start[k], R.rd_len[k], &(*out)[R.slot[k]]);
    });
    if (timing) fprintf(stderr, "[ssw_cpp] %lld pairs: pack %.3f ms, engine %.3f ms, post-process %.3f ms\n", (long long)R.n, R.t_packed - R.t_start, R.t_gpu - R.t_packed, now_ms() - R.t_gpu);
    return true;
}

bool Aligner::AlignIndexedCompact(const std::vector<SeqView>& pool, const std::vector<PairIndex>& pairs, const Filter& filter, CompactAlignments* out) const
{
    if (translate_.empty() || !out) return false;
    static const bool timing = getenv("MPN_TIMING") != nullptr;
    IndexedRun R;
    R.t_start = now_ms();
    const size_t np = pairs.size();
    out->sw_score.assign(np, 0); out->ref_begin.assign(np, 0); out->mismatches.assign(np, 0);
    out->cigar.assign(np, nullptr); out->cigar_len.assign(np, 0);
    out->text.clear();
    if (!RunIndexed(pool, pairs, filter, &R)) return false;
    // ---- the same '=' / 'X' walk as finish_alignment, CIGAR text appended to one buffer per block of pairs (no per-pair allocation)
    constexpr int64_t BLOCK = 1024;
    const int64_t nblocks = (R.n + BLOCK - 1) / BLOCK;
    out->text.resize((size_t)nblocks);
    mpn::parallel_for(nblocks, 1, [&](int64_t blk) {
        std::string& buf = out->text[(size_t)blk];
        const int64_t k0 = blk * BLOCK, k1 = std::min(R.n, k0 + BLOCK);
        buf.reserve((size_t)(k1 - k0) * 24);
        std::vector<int32_t> at((size_t)(k1 - k0) + 1, 0);
        Alignment tmp;
        for (int64_t k = k0; k < k1; ++k) {
            finish_alignment(R.res[k], R.cig.get(), R.arena + R.rf_start[k], R.arena + R.rd_start[k], R.rd_len[k], &tmp);
            const size_t i = R.slot[(size_t)k];
            out->sw_score[i] = tmp.sw_score; out->ref_begin[i] = tmp.ref_begin; out->mismatches[i] = tmp.mismatches;
            at[(size_t)(k - k0)] = (int32_t)buf.size();
            buf.append(tmp.cigar_string);
            out->cigar_len[i] = (int32_t)tmp.cigar_string.size();
        }
        for (int64_t k = k0; k < k1; ++k) out->cigar[R.slot[(size_t)k]] = buf.data() + at[(size_t)(k - k0)];      // after the last append: the buffer no longer moves
    });
    if (timing) fprintf(stderr, "[ssw_cpp] %lld pairs (compact): pack %.3f ms, engine %.3f ms, post-process %.3f ms\n", (long long)R.n, R.t_packed - R.t_start, R.t_gpu - R.t_packed, now_ms() - R.t_gpu);
    return true;
}

bool Aligner::AlignBatch(const std::vector<std::string>& queries, const Filter& filter, std::vector<Alignment>* out) const
{
    if (translate_.empty() || reference_.empty()) return false;
    std::vector<PairView> pv(queries.size());
    for (size_t i = 0; i < queries.size(); ++i) pv[i] = PairView{queries[i].c_str(), (int)strlen(queries[i].c_str()), nullptr, (int)reference_.size()};
    return AlignPairs(pv, filter, out);
}

bool Aligner::Align(const char* query, const Filter& filter, Alignment* alignment) const
{
    if (translate_.empty() || reference_.empty()) return false;
    const int qlen = (int)strlen(query);
    if (qlen == 0) return false;
    std::vector<PairView> pv(1, PairView{query, qlen, nullptr, (int)reference_.size()});
    std::vector<Alignment> res;
    if (!AlignPairs(pv, filter, &res)) return false;
    *alignment = res[0];
    return true;
}

bool Aligner::Align(const char* query, const char* ref, const int& ref_len, const Filter& filter, Alignment* alignment) const
{
    if (translate_.empty()) return false;
    const int qlen = (int)strlen(query);
    if (qlen == 0) return false;
    std::vector<PairView> pv(1, PairView{query, qlen, ref, ref_len});
    std::vector<Alignment> res;
    if (!AlignPairs(pv, filter, &res)) return false;
    *alignment = res[0];
    return true;
}

void Aligner::Clear(void)
{
    matrix_.clear(); translate_.clear(); reference_.clear();
}

bool Aligner::ReBuild(void)
{
    if (!translate_.empty()) return false;
    match_ = 4; mismatch_ = 6; gap_open_ = 8; gap_extend_ = 2;
    default_tables();
    return true;
}

bool Aligner::ReBuild(const uint8_t& match_score, const uint8_t& mismatch_penalty, const uint8_t& gap_opening_penalty, const uint8_t& gap_extending_penalty)
{
    if (!translate_.empty()) return false;
    match_ = match_score; mismatch_ = mismatch_penalty; gap_open_ = gap_opening_penalty; gap_extend_ = gap_extending_penalty;
    default_tables();
    return true;
}

bool Aligner::ReBuild(const int8_t* score_matrix, const int& score_matrix_size, const int8_t* translation_matrix, const int& translation_matrix_size)
{
    n_ = score_matrix_size;
    matrix_.assign(score_matrix, score_matrix + (size_t)n_ * n_);
    translate_.assign(translation_matrix, translation_matrix + translation_matrix_size);
    return true;
}

}  // namespace StripedSmithWaterman
