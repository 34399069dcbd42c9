// Region realigner on the batched GPU engine: the host logic around the Smith-Waterman calls of the reference's
// realigner.cpp, reorganised so that every alignment of a region (or of many regions) is ONE submit to the CUDA kernels.
//
// Per region, in the reference's order (ReAligner::AlignReads, realigner.cpp:88-117):
//   1. k-mer index over the reads (k = 32, BuildIndex :429-451) and the <= 2-mismatch ungapped "fast pass" of every read
//      against every haplotype (FastAlignReadsToHaplotype :170-230, FastAlignStrings :232-253)
//      [GPU: mpn_fastpass, all regions in one launch, no index; host k-mer index only for regions outside the kernel's limits]
//   2. haplotype -> reference alignments (AlignHaplotypesToReference :325-349) and, for reads the fast pass could not
//      place, read -> haplotype alignments (SswAlignReadsToHaplotypes :351-384)                      [GPU, one batch]
//   3. haplotype position maps (SetPositionsMap :453-509), best haplotype per read (GetBestReadAlignment :517-540) and
//      composition of read->haplotype with haplotype->reference CIGARs (CalculateReadToRefAlignment :640-777)     [host]
// Steps 1 and 3 are plain host code written against flat vectors (no std::list / std::regex); step 2 never runs on the CPU.
#include "../../include/realigner.h"
#include "../../include/ssw_cpp.h"
#include "host_shared.h"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

namespace {

using StripedSmithWaterman::Alignment;
using StripedSmithWaterman::PairIndex;
using StripedSmithWaterman::SeqView;

// ---- options fixed by ReAligner::set_options (realigner.cpp:62-72)
constexpr int kKmer = 32;
constexpr int kNominalReadSize = 250;
constexpr int kMaxMismatches = 2;
constexpr double kSimilarity = 0.16934;
constexpr int kMatch = 4, kMismatch = 6, kGapOpen = 8, kGapExtend = 2;
constexpr int kNotPlaced = -1;

enum OpKind : int { OP_NONE = 0, OP_MATCH = 1, OP_INS = 2, OP_DEL = 3, OP_SKIP = 4, OP_SOFT = 5, OP_HARD = 6 };
struct Op { OpKind kind; int len; };

struct Placement {          // a read on a haplotype (ReadAlignment, realigner.h:103-127)
    int pos = kNotPlaced;
    int score = 0;
    std::string cigar;
    void reset() { pos = kNotPlaced; score = 0; cigar.clear(); }
};

struct HapRecord {          // HaplotypeReadsAlignment, realigner.h:163-211
    int index = 0;
    int score = 0;
    std::vector<Placement> reads;
    std::string cigar;      // haplotype -> reference
    std::vector<Op> ops;
    int ref_pos = 0;
    std::vector<int> shift; // hap position -> cumulative shift against the reference
    bool is_reference = false;
    bool operator<(const HapRecord& o) const { return score < o.score; }
};

int ssw_score_threshold()
{
    // realigner.cpp:74-84: evaluated in double, truncated, floored at 1
    int t = kMatch * kNominalReadSize * kSimilarity - kMismatch * kNominalReadSize * (1 - kSimilarity);
    return t < 0 ? 1 : t;
}

// "<digits><op>" tokens with op in [XIDS=] (either case); anything else is skipped, like the reference's regex_search loop
// (realigner.cpp:272-291).  Lower-case letters tokenise but map to OP_NONE (CigarOperationFromChar :255-270 is case sensitive).
template <class F> void scan_cigar(const std::string& s, F&& emit)
{
    const size_t n = s.size();
    size_t i = 0;
    while (i < n) {
        if (s[i] < '0' || s[i] > '9') { ++i; continue; }
        size_t j = i;
        while (j < n && s[j] >= '0' && s[j] <= '9') ++j;
        if (j < n && strchr("XIDS=xids", s[j])) {
            emit(atoi(s.substr(i, j - i).c_str()), s[j]);
            i = j + 1;
        } else {
            i = j;
        }
    }
}

OpKind kind_of(char c)
{
    switch (c) {
        case '=': case 'X': return OP_MATCH;
        case 'S': return OP_SOFT;
        case 'D': return OP_DEL;
        case 'I': return OP_INS;
        default: return OP_NONE;
    }
}

std::vector<Op> parse_ops(const std::string& s)
{
    std::vector<Op> v;
    scan_cigar(s, [&](int len, char c) { v.push_back(Op{kind_of(c), len}); });
    return v;
}

std::string ops_to_string(const std::vector<Op>& ops)      // CigarVectorToString, realigner.cpp:294-317 (matches print as X)
{
    std::string out;
    for (const Op& o : ops) {
        out += std::to_string(o.len);
        switch (o.kind) {
            case OP_MATCH: out += 'X'; break;
            case OP_INS: out += 'I'; break;
            case OP_DEL: out += 'D'; break;
            case OP_SOFT: out += 'S'; break;
            default: break;
        }
    }
    return out;
}

// ------------------------------------------------------------------------------------------------ k-mer index
// The reference keys an unordered_map by the 32-character substring.  Here: k-mers made only of A/C/G/T pack into one
// 64-bit word (flat hash table, occurrences grouped per k-mer in insertion order = (read, offset) ascending); k-mers with
// any other character (N, lower case) go to a small string-keyed map so that equality stays exact string equality.
struct Occurrence { uint64_t key; int read; int offset; };

inline int base2(char c)
{
    switch (c) { case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3; default: return -1; }
}

struct KmerIndex {
    // packed k-mers: open-addressing table key -> (first, count) into `occ`, where the occurrences of one k-mer are contiguous and
    // keep their insertion order = (read, offset) ascending (the order the reference's vectors have, realigner.cpp:425-427)
    struct Slot { uint64_t key; int first; int count; };
    std::vector<Slot> table;
    std::vector<Occurrence> occ;
    uint64_t mask = 0;
    std::unordered_map<std::string, std::vector<Occurrence>> odd;

    static inline uint64_t mix(uint64_t k) { k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; return k; }

    void build(const std::vector<std::string>& reads)
    {
        size_t total = 0;
        for (const std::string& r : reads) if ((int)r.size() > kKmer) total += r.size() - kKmer + 1;
        std::vector<Occurrence> raw;
        raw.reserve(total);
        for (int id = 0; id < (int)reads.size(); ++id) {
            const std::string& r = reads[id];
            if (r.length() <= (size_t)kKmer) continue;                      // AddReadToIndex, realigner.cpp:436-438
            uint64_t key = 0; int valid = 0;
            for (int i = 0; i < (int)r.size(); ++i) {
                const int b = base2(r[i]);
                if (b < 0) { valid = 0; key = 0; } else { key = (key << 2) | (uint64_t)b; ++valid; }
                const int start = i - kKmer + 1;
                if (start < 0) continue;
                if (valid >= kKmer) raw.push_back(Occurrence{key, id, start});
                else odd[r.substr((size_t)start, kKmer)].push_back(Occurrence{0, id, start});
            }
        }
        size_t cap = 64;
        while (cap < raw.size() * 2 + 2) cap <<= 1;
        mask = cap - 1;
        table.assign(cap, Slot{0, -1, 0});
        // pass 1: count per key; pass 2: prefix offsets; pass 3: scatter in insertion order
        std::vector<int> slot_of(raw.size());
        for (size_t q = 0; q < raw.size(); ++q) {
            uint64_t h = mix(raw[q].key) & mask;
            while (table[h].first != -1 && table[h].key != raw[q].key) h = (h + 1) & mask;
            if (table[h].first == -1) { table[h].key = raw[q].key; table[h].first = 0; }
            table[h].count++;
            slot_of[q] = (int)h;
        }
        int at = 0;
        for (Slot& sl : table) if (sl.first != -1) { sl.first = at; at += sl.count; sl.count = 0; }
        occ.resize(raw.size());
        for (size_t q = 0; q < raw.size(); ++q) { Slot& sl = table[(size_t)slot_of[q]]; occ[(size_t)(sl.first + sl.count++)] = raw[q]; }
    }

    // occurrences of the k-mer starting at hap[i]; `key`/`valid` are the caller's rolling state for that window
    std::pair<const Occurrence*, const Occurrence*> find(const std::string& hap, int i, uint64_t key, bool valid) const
    {
        if (valid) {
            if (table.empty()) return {nullptr, nullptr};
            uint64_t h = mix(key) & mask;
            while (table[h].first != -1 && table[h].key != key) h = (h + 1) & mask;
            if (table[h].first == -1) return {nullptr, nullptr};
            return {occ.data() + table[h].first, occ.data() + table[h].first + table[h].count};
        }
        if (odd.empty()) return {nullptr, nullptr};
        auto it = odd.find(hap.substr((size_t)i, kKmer));
        if (it == odd.end()) return {nullptr, nullptr};
        return {it->second.data(), it->second.data() + it->second.size()};
    }
};

// ungapped comparison, FastAlignStrings (realigner.cpp:232-253): N on either side counts as a match; gives up at the
// (max+1)-th mismatch with score 0
int fast_compare(const char* hap, const char* read, int len, int give_up_at, int* mism)
{
    int same = 0;
    *mism = 0;
    for (int i = 0; i < len; ++i) {
        const char a = hap[i], b = read[i];
        if (a != b && a != 'N' && b != 'N') {
            if (++*mism == give_up_at) return 0;
        } else {
            ++same;
        }
    }
    return same * kMatch - *mism * kMismatch;
}

struct Region {
    std::string reference;
    std::vector<std::string> haplotypes;
    std::vector<std::string> reads;
    std::vector<int> in_pos;
    std::vector<std::string> in_cigar;
    int ref_start = 0, prefix = 0, suffix = 0;
    // state
    std::vector<HapRecord> haps;
    // the Smith-Waterman work of this region inside the global batch
    size_t first_pair = 0;
    std::vector<std::pair<int, int>> read_hap_pairs;     // (read, index into haps) in the reference's loop order
};

// fast pass of every read against one haplotype (realigner.cpp:170-230)
void fast_pass_one(const Region& rg, const KmerIndex& index, const std::string& hap, int* hap_score, std::vector<Placement>* placed)
{
    const bool is_ref = hap == rg.reference;
    std::vector<int> coverage(hap.size(), 0);
    if (hap.length() < (size_t)kKmer) return;            // the reference's unsigned loop bound wraps here and throws; nothing sensible to reproduce
    const int last = (int)hap.length() - kKmer;
    // Consecutive k-mers of a read point at the same start on the haplotype, and repeating the comparison there can change nothing
    // (a failed one has no effect, a passed one only re-increments coverage that is tested against zero and cannot raise the
    // read's score again): remember the last start compared per read and skip exact repeats.
    std::vector<int> last_start(rg.reads.size(), -2);
    uint64_t key = 0; int valid = 0;
    for (int i = 0; i < kKmer - 1; ++i) {
        const int b = base2(hap[i]);
        if (b < 0) { valid = 0; key = 0; } else { key = (key << 2) | (uint64_t)b; ++valid; }
    }
    for (int i = 0; i <= last; ++i) {
        const int b = base2(hap[i + kKmer - 1]);
        if (b < 0) { valid = 0; key = 0; } else { key = (key << 2) | (uint64_t)b; ++valid; }
        auto range = index.find(hap, i, key, valid >= kKmer);
        if (range.first == range.second) continue;       // no read shares this k-mer: the coverage test below is skipped too
        for (const Occurrence* oc = range.first; oc != range.second; ++oc) {
            const std::string& read = rg.reads[oc->read];
            const int start = std::max(0, i - oc->offset);
            const int rlen = (int)read.size();
            if ((size_t)start + (size_t)rlen > hap.length()) continue;
            Placement& pl = (*placed)[oc->read];
            if (pl.pos != kNotPlaced && pl.pos == start) continue;
            if (last_start[oc->read] == start) continue;
            last_start[oc->read] = start;
            int mism = 0;
            const int sc = fast_compare(hap.data() + start, read.data(), rlen, kMaxMismatches + 1, &mism);
            if (mism <= kMaxMismatches) {
                const int old = pl.score;
                for (int p = start; p < start + rlen; ++p) coverage[p]++;
                if (old < sc) {
                    pl.score = sc;
                    *hap_score += sc - old;
                    pl.pos = start;
                    pl.cigar = std::to_string(rlen) + "=";
                }
            }
        }
        if (coverage[i] == 0 && i >= rg.prefix && (size_t)i < hap.size() - (size_t)rg.suffix && !is_ref) {
            *hap_score = 0;                               // a non-reference haplotype with an uncovered base inside the window is dropped
            return;
        }
    }
}

void fast_pass(Region& rg, bool threads)
{
    KmerIndex index;
    index.build(rg.reads);
    const int nh = (int)rg.haplotypes.size();
    rg.haps.assign((size_t)nh, HapRecord());
    auto one = [&](int64_t h) {                       // haplotypes are independent of each other (realigner.cpp:146-168)
        HapRecord& rec = rg.haps[(size_t)h];
        rec.index = (int)h; rec.score = 0;
        rec.reads.assign(rg.reads.size(), Placement());
        fast_pass_one(rg, index, rg.haplotypes[(size_t)h], &rec.score, &rec.reads);
        if (rec.score == 0) for (Placement& p : rec.reads) p.reset();
    };
    if (threads) mpn::parallel_for(nh, 1, one); else for (int h = 0; h < nh; ++h) one(h);
}

double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// The fast pass of many regions in one GPU call (mpn_fastpass, include/mpn_ssw_batch.h).  Regions the kernel does not take -- reads
// longer than 256 bases, haplotypes longer than 2816, bases other than A,C,G,T,N (the kernel's 3-bit alphabet; the reference compares
// raw characters) -- keep the string-exact host pass above, which is the placement the reference itself has for this step.
void fast_pass_regions(std::vector<Region>& rgs, double* kernel_ms)
{
    const int nr = (int)rgs.size();
    static const bool timing = getenv("MPN_TIMING") != nullptr;
    const double t_in = now_s();
    static const bool force_host = getenv("MPN_FASTPASS_HOST") != nullptr;      // A/B timing of the two placements of this step
    std::vector<char> on_gpu((size_t)nr, force_host ? 0 : 1);
    mpn::parallel_for(nr, 4, [&](int64_t g) {
        const Region& rg = rgs[(size_t)g];
        if (rg.haplotypes.empty() || rg.reads.empty()) { on_gpu[(size_t)g] = 0; return; }
        for (const std::string& h : rg.haplotypes) if (h.size() > (size_t)MPN_FP_MAX_HAP) { on_gpu[(size_t)g] = 0; return; }
        for (const std::string& q : rg.reads) if (q.size() > (size_t)MPN_FP_MAX_READ) { on_gpu[(size_t)g] = 0; return; }
    });
    // spans of every haplotype and read inside one text buffer
    std::vector<mpn_fp_region> fr;
    std::vector<int> region_of;
    std::vector<int64_t> hap_start, read_start;
    std::vector<int32_t> hap_len, read_len;
    std::vector<uint8_t> is_ref;
    std::vector<const std::string*> src;
    int64_t bytes = 0, nplaces = 0;
    for (int g = 0; g < nr; ++g) {
        if (!on_gpu[(size_t)g]) continue;
        const Region& rg = rgs[(size_t)g];
        mpn_fp_region f;
        f.place_first = nplaces; f.hap_first = (int32_t)hap_start.size(); f.nhap = (int32_t)rg.haplotypes.size();
        f.read_first = (int32_t)read_start.size(); f.nread = (int32_t)rg.reads.size(); f.prefix = rg.prefix; f.suffix = rg.suffix;
        for (const std::string& h : rg.haplotypes) { hap_start.push_back(bytes); hap_len.push_back((int32_t)h.size()); is_ref.push_back(h == rg.reference); src.push_back(&h); bytes += (int64_t)h.size(); }
        for (const std::string& q : rg.reads) { read_start.push_back(bytes); read_len.push_back((int32_t)q.size()); src.push_back(&q); bytes += (int64_t)q.size(); }
        nplaces += (int64_t)f.nhap * f.nread;
        fr.push_back(f); region_of.push_back(g);
    }
    // grow-only buffers kept across calls: fresh allocations of this size cost more in page faults than the copies themselves
    static thread_local std::vector<mpn_placement> places;
    static thread_local std::vector<char> text_buf;
    if (places.size() < (size_t)nplaces) places.resize((size_t)nplaces);
    mpn_placement* const places_p = places.data();       // thread_local: worker threads must not name `places` themselves
    std::vector<int32_t> hap_score(hap_start.size());
    std::vector<uint8_t> flag(fr.size(), 0);
    if (!fr.empty()) {
        if (text_buf.size() < (size_t)bytes + 1) text_buf.resize((size_t)bytes + 1);
        char* const text_p = text_buf.data();
        {
            std::vector<int64_t> at(src.size());
            int64_t b = 0;
            for (size_t k = 0; k < src.size(); ++k) { at[k] = b; b += (int64_t)src[k]->size(); }
            mpn::parallel_for((int64_t)src.size(), 256, [&](int64_t k) { memcpy(text_p + at[(size_t)k], src[(size_t)k]->data(), src[(size_t)k]->size()); });
        }
        mpn::SharedEngineLock lk;
        const int rc = mpn_fastpass(lk.engine(), text_p, bytes, hap_start.data(), hap_len.data(), is_ref.data(), (int32_t)hap_start.size(),
                                    read_start.data(), read_len.data(), (int32_t)read_start.size(), fr.data(), (int32_t)fr.size(),
                                    places_p, hap_score.data(), flag.data());
        if (rc != 0) { fprintf(stderr, "[realigner] mpn_fastpass failed (code %d)\n", rc); abort(); }
        if (kernel_ms) *kernel_ms = mpn_fastpass_last_kernel_ms(lk.engine());
    }
    const double t_gpu = now_s();
    std::vector<int> slot_of((size_t)nr, -1);
    for (size_t k = 0; k < region_of.size(); ++k) slot_of[(size_t)region_of[k]] = flag[k] ? -1 : (int)k;
    mpn::parallel_for(nr, 1, [&](int64_t g) {
        Region& rg = rgs[(size_t)g];
        const int k = slot_of[(size_t)g];
        if (k < 0) { fast_pass(rg, nr == 1 && rg.reads.size() * rg.haplotypes.size() >= 2048); return; }
        const mpn_fp_region& f = fr[(size_t)k];
        rg.haps.assign((size_t)f.nhap, HapRecord());
        for (int h = 0; h < f.nhap; ++h) {
            HapRecord& rec = rg.haps[(size_t)h];
            rec.index = h; rec.score = hap_score[(size_t)(f.hap_first + h)];
            rec.reads.assign((size_t)f.nread, Placement());
            if (rec.score == 0) continue;                    // dropped or nothing placed: every read stays unplaced (realigner.cpp:160-165)
            const mpn_placement* pl = places_p + f.place_first + (int64_t)h * f.nread;
            for (int r = 0; r < f.nread; ++r)
                if (pl[r].score > 0) { rec.reads[(size_t)r].score = pl[r].score; rec.reads[(size_t)r].pos = pl[r].pos; rec.reads[(size_t)r].cigar = std::to_string(rg.reads[(size_t)r].size()) + "="; }
        }
    });
    if (timing) fprintf(stderr, "[realigner] fast pass of %d regions: through the GPU call %.3f ms, records %.3f ms\n", nr, 1e3 * (t_gpu - t_in), 1e3 * (now_s() - t_gpu));
}

// which (read, haplotype) pairs need Smith-Waterman (realigner.cpp:351-366)
void plan_read_pairs(Region& rg)
{
    rg.read_hap_pairs.clear();
    for (int r = 0; r < (int)rg.reads.size(); ++r) {
        bool placed = false;
        for (const HapRecord& h : rg.haps) if (h.reads[r].score > 0) { placed = true; break; }
        if (placed) continue;
        for (int k = 0; k < (int)rg.haps.size(); ++k) if (rg.haps[k].score != 0) rg.read_hap_pairs.push_back({r, k});
    }
}

// hap position -> shift against the reference, SetPositionsMap (realigner.cpp:453-509)
void build_shift_map(HapRecord& h, int hap_len)
{
    h.shift.assign((size_t)hap_len, 0);
    int shift = 0, pos = 0;
    auto put = [&](int p, int v) { if (p >= 0 && p < hap_len) h.shift[p] = v; };
    scan_cigar(h.cigar, [&](int len, char c) {
        switch (c) {
            case '=': case 'X': for (int e = pos + len; pos != e; ++pos) put(pos, shift); break;
            case 'S': shift -= len; for (int e = pos + len; pos != e; ++pos) put(pos, shift); break;
            case 'D': shift += len; break;
            case 'I': for (int e = pos + len; pos != e; ++pos) { put(pos, shift); --shift; } break;
            default: break;
        }
    });
}

// ---- CIGAR composition.  `out` plus the number of read bases it already covers (deletions cover none).
struct Composed {
    std::vector<Op> ops;
    int covered = 0;
    // MergeCigarOp (realigner.cpp:553-577): clip to the read length, fuse with an equal trailing op
    void merge(OpKind kind, int len, int read_len)
    {
        const OpKind last = ops.empty() ? OP_NONE : ops.back().kind;
        const int before = covered;
        const int take = kind != OP_DEL ? std::min(len, read_len - before) : len;
        if (take <= 0 || before == read_len) return;
        if (kind == last) ops.back().len += take; else ops.push_back(Op{kind, take});
        if (kind != OP_DEL) covered += take;
    }
};

inline bool matchlike(const Op& o) { return o.kind == OP_MATCH || o.kind == OP_SOFT; }

// drop the part of the haplotype->reference CIGAR left of the read's start on the haplotype (realigner.cpp:581-612)
std::deque<Op> trim_left(const std::vector<Op>& hap_ops, int read_pos, bool* ok)
{
    std::deque<Op> q(hap_ops.begin(), hap_ops.end());
    int cur = 0;
    *ok = true;
    while (cur != read_pos) {
        if (q.empty()) { *ok = false; return q; }           // the reference reads front() of an empty list here (undefined)
        const Op o = q.front();
        q.pop_front();
        if (o.kind == OP_MATCH || o.kind == OP_HARD || o.kind == OP_SOFT || o.kind == OP_INS) {
            if (o.len + cur > read_pos) q.push_front(Op{o.kind, o.len - (read_pos - cur)});
            cur = std::min(o.len + cur, read_pos);
        }
    }
    if (q.empty()) { *ok = false; return q; }
    if (q.front().kind == OP_DEL) q.pop_front();
    return q;
}

// read->haplotype o haplotype->reference (CalculateReadToRefAlignment, realigner.cpp:640-777); empty result = "keep the read as it was"
std::vector<Op> compose(int read_len, const Placement& on_hap, const std::vector<Op>& hap_ops)
{
    Composed out;
    std::vector<Op> parsed = parse_ops(on_hap.cigar);
    std::deque<Op> rh(parsed.begin(), parsed.end());
    bool ok = true;
    std::deque<Op> hr = trim_left(hap_ops, on_hap.pos, &ok);
    if (!ok) return {};
    if (!rh.empty() && rh.front().kind == OP_SOFT) {
        out.merge(OP_SOFT, rh.front().len, read_len);
        rh.pop_front();
    }
    while ((!rh.empty() || !hr.empty()) && out.covered < read_len) {
        if (!rh.empty() && hr.empty()) {
            out.merge(rh.front().kind, rh.front().len, read_len);
            rh.pop_front();
            continue;
        }
        if (rh.empty()) break;
        Op a = rh.front(); rh.pop_front();       // read -> haplotype
        Op b = hr.front(); hr.pop_front();       // haplotype -> reference
        if (matchlike(a) && matchlike(b)) {
            const int n = std::min(a.len, b.len);
            out.merge((a.kind == OP_SOFT || b.kind == OP_SOFT) ? OP_SOFT : OP_MATCH, n, read_len);
            a.len -= n; if (a.len > 0) rh.push_front(a);
            b.len -= n; if (b.len > 0) hr.push_front(b);
        } else if (a.kind == OP_DEL && matchlike(b)) {
            out.merge(OP_DEL, a.len, read_len);
            b.len -= a.len; if (b.len > 0) hr.push_front(b);
        } else if (b.kind == OP_DEL && matchlike(a)) {
            out.merge(OP_DEL, b.len, read_len);
            if (a.len > 0) rh.push_front(a);
        } else if (a.kind == OP_DEL && b.kind == OP_DEL) {
            out.merge(OP_DEL, a.len + b.len, read_len);
        } else if (a.kind == OP_INS && matchlike(b)) {
            a.len = std::min(read_len - out.covered, a.len);
            out.merge(OP_INS, a.len, read_len);
            if (b.len > 0) hr.push_front(b);
        } else if (b.kind == OP_INS && matchlike(a)) {
            b.len = std::min(read_len - out.covered, b.len);
            out.merge(OP_INS, b.len, read_len);
            a.len = std::max(0, a.len - b.len);
            if (a.len > 0) rh.push_front(a);
        } else if (a.kind == OP_INS && b.kind == OP_INS) {
            out.merge(OP_INS, a.len + b.len, read_len);
        } else {
            return {};                               // combination the reference does not handle: alignment discarded
        }
    }
    return out.ops;
}

// GetBestReadAlignment (realigner.cpp:517-540): highest score; on ties the later non-reference haplotype wins
bool best_haplotype(const Region& rg, int read, int* best)
{
    int top = 0;
    bool found = false;
    for (int k = 0; k < (int)rg.haplotypes.size() && k < (int)rg.haps.size(); ++k) {
        const int s = rg.haps[k].reads[read].score;
        if (s > top || (top > 0 && s == top && !rg.haps[k].is_reference)) { top = s; *best = k; found = true; }
    }
    return found;
}

struct Stats { long long pairs = 0, cells = 0; double t_fast = 0, t_gpu = 0, t_compose = 0, fp_kernel_ms = 0; };
Stats g_last;


void load_region(Region& rg, const mpn_region& in)
{
    rg.reference = in.reference ? in.reference : "";
    rg.haplotypes.clear();
    if (in.haplotypes) {                                     // white-space separated (realigner.cpp:788-794)
        const char* p = in.haplotypes;
        while (*p) {
            while (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r' || *p == '\v' || *p == '\f') ++p;
            const char* q = p;
            while (*q && !(*q == ' ' || *q == '\t' || *q == '\n' || *q == '\r' || *q == '\v' || *q == '\f')) ++q;
            if (q > p) rg.haplotypes.emplace_back(p, q);
            p = q;
        }
    }
    rg.reads.resize((size_t)in.read_size); rg.in_pos.resize((size_t)in.read_size); rg.in_cigar.resize((size_t)in.read_size);
    for (int i = 0; i < in.read_size; ++i) { rg.reads[i] = in.seqs[i]; rg.in_pos[i] = in.positions[i]; rg.in_cigar[i] = in.cigars[i]; }
    rg.ref_start = in.ref_start; rg.prefix = in.ref_prefix; rg.suffix = in.ref_suffix;
}

int realign_many(const mpn_region* regions, int nregions, struct_str_arr** out)
{
    std::vector<Region> rgs((size_t)nregions);
    Stats st;
    double t0 = now_s();
    // ---- 1. k-mer fast pass (GPU), then the list of Smith-Waterman pairs of every region
    mpn::parallel_for(nregions, 1, [&](int64_t r) { load_region(rgs[(size_t)r], regions[r]); });
    const double t_loaded = now_s();
    // every (haplotype, read) of every region in one kernel launch; then the Smith-Waterman work list per region
    fast_pass_regions(rgs, &st.fp_kernel_ms);
    const double t_fp = now_s();
    mpn::parallel_for(nregions, 1, [&](int64_t r) { plan_read_pairs(rgs[(size_t)r]); });
    if (getenv("MPN_TIMING")) fprintf(stderr, "[realigner] load %.3f ms, fast pass %.3f ms, plan %.3f ms\n", 1e3 * (t_loaded - t0), 1e3 * (t_fp - t_loaded), 1e3 * (now_s() - t_fp));
    // sequence pool: per region the reference, its haplotypes and its reads, each once; pairs name them by index
    std::vector<SeqView> pool;
    std::vector<PairIndex> pairs;
    for (int r = 0; r < nregions; ++r) {
        Region& rg = rgs[r];
        const int32_t ref_id = (int32_t)pool.size();
        pool.push_back(SeqView{rg.reference.c_str(), (int)rg.reference.length()});
        const int32_t hap0 = (int32_t)pool.size();
        for (const std::string& h : rg.haplotypes) pool.push_back(SeqView{h.c_str(), (int)strlen(h.c_str())});
        const int32_t read0 = (int32_t)pool.size();
        for (const std::string& q : rg.reads) pool.push_back(SeqView{q.c_str(), (int)strlen(q.c_str())});
        rg.first_pair = pairs.size();
        for (const HapRecord& h : rg.haps) pairs.push_back(PairIndex{hap0 + h.index, ref_id});
        for (const auto& rh : rg.read_hap_pairs) pairs.push_back(PairIndex{read0 + rh.first, hap0 + rg.haps[rh.second].index});
    }
    for (const PairIndex& p : pairs) { st.pairs++; st.cells += (long long)pool[(size_t)p.query].len * pool[(size_t)p.target].len; }
    double t1 = now_s();
    // ---- 2. GPU: one batch (flag 0x0f, no filters, maskLen = query length: StripedSmithWaterman defaults, ssw_cpp.cpp:343-346)
    // results in compact form (score, begin, CIGAR text in shared buffers): one heap string + one vector per pair would cost more to build and
    // to destroy than the whole GPU batch takes
    StripedSmithWaterman::CompactAlignments aln;
    std::vector<int64_t> at(pairs.size(), -1);             // pair -> entry of aln; -1: the pair was not aligned (stays "score 0")
    {
        StripedSmithWaterman::Aligner aligner(kMatch, kMismatch, kGapOpen, kGapExtend);
        StripedSmithWaterman::Filter filter;
        // a zero-length reference makes Aligner::Align return false in the reference (ssw_cpp.cpp:330); those pairs stay cleared
        bool all_live = true;
        for (const PairIndex& p : pairs) if (pool[(size_t)p.target].len <= 0 || pool[(size_t)p.query].len <= 0) { all_live = false; break; }
        if (all_live) {
            aligner.AlignIndexedCompact(pool, pairs, filter, &aln);
            for (size_t i = 0; i < pairs.size(); ++i) at[i] = (int64_t)i;
        } else {
            std::vector<PairIndex> live;
            for (size_t i = 0; i < pairs.size(); ++i) if (pool[(size_t)pairs[i].target].len > 0 && pool[(size_t)pairs[i].query].len > 0) { at[i] = (int64_t)live.size(); live.push_back(pairs[i]); }
            aligner.AlignIndexedCompact(pool, live, filter, &aln);
        }
    }
    double t2 = now_s();
    // ---- 3. host: consume
    const int threshold = ssw_score_threshold();
    mpn::parallel_for(nregions, 1, [&](int64_t r) {
        Region& rg = rgs[r];
        size_t k = rg.first_pair;
        for (HapRecord& h : rg.haps) {                                   // realigner.cpp:336-348
            const int64_t a = at[k++];
            if (a >= 0 && aln.sw_score[(size_t)a] > 0) {
                h.cigar.assign(aln.cigar[(size_t)a], (size_t)aln.cigar_len[(size_t)a]);
                h.is_reference = h.cigar == std::to_string(rg.haplotypes[h.index].size()) + "=";
                h.ops = parse_ops(h.cigar);
                h.ref_pos = aln.ref_begin[(size_t)a];
            }
            build_shift_map(h, (int)rg.haplotypes[h.index].size());
        }
        for (const auto& rh : rg.read_hap_pairs) {                       // realigner.cpp:369-379
            const int64_t a = at[k++];
            Placement& pl = rg.haps[rh.second].reads[rh.first];
            const int sc = a >= 0 ? aln.sw_score[(size_t)a] : 0;
            if (sc > 0 && sc >= threshold && pl.score < sc) {
                pl.score = sc; pl.cigar.assign(aln.cigar[(size_t)a], (size_t)aln.cigar_len[(size_t)a]); pl.pos = aln.ref_begin[(size_t)a];
            }
        }
        std::sort(rg.haps.begin(), rg.haps.end());                       // realigner.cpp:105-106 (same comparator, same algorithm)
        struct_str_arr* res = new struct_str_arr();
        memset(res, 0, sizeof *res);
        for (int i = 0; i < (int)rg.reads.size(); ++i) {
            int pos = rg.in_pos[i];
            std::string cigar = rg.in_cigar[i];
            int best = -1;
            if (best_haplotype(rg, i, &best)) {
                const HapRecord& h = rg.haps[best];
                const Placement& pl = h.reads[i];
                if (pl.pos >= 0 && pl.pos < (int)h.shift.size()) {
                    const int new_pos = rg.ref_start + h.ref_pos + pl.pos + h.shift[pl.pos];
                    const std::vector<Op> ops = compose((int)rg.reads[i].length(), pl, h.ops);
                    if (!ops.empty()) { cigar = ops_to_string(ops); pos = new_pos; }
                }
            }
            if (i < 1000) {
                res->cigar_string[i] = new char[cigar.size() + 1];
                memcpy(res->cigar_string[i], cigar.c_str(), cigar.size() + 1);
                res->position[i] = pos;
            }
        }
        out[r] = res;
    });
    double t3 = now_s();
    st.t_fast = t1 - t0; st.t_gpu = t2 - t1; st.t_compose = t3 - t2;
    g_last = st;
    return 0;
}

}  // namespace

extern "C" struct_str_arr* realign_reads(char* seqs[], int* positions, char* cigars[], char* reference, char* haplotypes,
                                         int ref_start, int ref_prefix, int ref_suffix, int read_size)
{
    mpn_region rg{seqs, positions, cigars, read_size, reference, haplotypes, ref_start, ref_prefix, ref_suffix};
    struct_str_arr* out = nullptr;
    realign_many(&rg, 1, &out);
    return out;
}

extern "C" int mpn_realign_regions(const mpn_region* regions, int nregions, struct_str_arr** out)
{
    if (nregions < 0 || (nregions > 0 && (!regions || !out))) return -1;
    for (int r = 0; r < nregions; ++r) if (regions[r].read_size < 0 || regions[r].read_size > 1000) return -1;
    return realign_many(regions, nregions, out);
}

extern "C" int mpn_realign_regions_packed(const char* text, long long text_bytes, int nregions, const int* region_reads, const int* region_geom,
                                          const int* positions, int* out_positions, char** out_cigars, long long* out_cigars_bytes)
{
    if (nregions < 0 || !out_cigars || !out_cigars_bytes || (nregions > 0 && (!text || !region_reads || !region_geom || !positions || !out_positions))) return -1;
    // unpack: pointers into the caller's text, one mpn_region per region
    std::vector<mpn_region> regs((size_t)nregions);
    std::vector<char*> ptrs;
    size_t total_reads = 0;
    for (int r = 0; r < nregions; ++r) { if (region_reads[r] < 0 || region_reads[r] > 1000) return -1; total_reads += (size_t)region_reads[r]; }
    ptrs.reserve(2 * total_reads);
    const char* p = text; const char* const end = text + text_bytes;
    auto next = [&]() -> char* { if (p >= end) return nullptr; char* s0 = const_cast<char*>(p); p += strlen(p) + 1; return s0; };
    size_t rd = 0;
    for (int r = 0; r < nregions; ++r) {
        mpn_region& g = regs[(size_t)r];
        g.reference = next(); g.haplotypes = next();
        g.read_size = region_reads[r];
        const size_t first = ptrs.size();
        for (int k = 0; k < 2 * g.read_size; ++k) ptrs.push_back(next());
        if (!g.reference || !g.haplotypes || (g.read_size > 0 && !ptrs.back())) return -1;
        g.positions = const_cast<int*>(positions) + rd;
        g.ref_start = region_geom[3 * r]; g.ref_prefix = region_geom[3 * r + 1]; g.ref_suffix = region_geom[3 * r + 2];
        g.seqs = nullptr; g.cigars = nullptr;
        rd += (size_t)g.read_size;
        (void)first;
    }
    {   // the pointer array is complete (no more reallocation): hook the regions up
        size_t at = 0;
        for (int r = 0; r < nregions; ++r) { regs[(size_t)r].seqs = ptrs.data() + at; regs[(size_t)r].cigars = ptrs.data() + at + regs[(size_t)r].read_size; at += 2 * (size_t)regs[(size_t)r].read_size; }
    }
    std::vector<struct_str_arr*> res((size_t)nregions, nullptr);
    const int rc = realign_many(regs.data(), nregions, res.data());
    if (rc != 0) return rc;
    size_t bytes = 0;
    for (int r = 0; r < nregions; ++r) for (int i = 0; i < regs[(size_t)r].read_size; ++i) bytes += strlen(res[(size_t)r]->cigar_string[i]) + 1;
    char* blob = (char*)malloc(bytes + 1);
    char* w = blob; size_t k = 0;
    for (int r = 0; r < nregions; ++r) {
        for (int i = 0; i < regs[(size_t)r].read_size; ++i) {
            const size_t n = strlen(res[(size_t)r]->cigar_string[i]) + 1;
            memcpy(w, res[(size_t)r]->cigar_string[i], n); w += n;
            out_positions[k++] = res[(size_t)r]->position[i];
        }
        free_memory(res[(size_t)r], regs[(size_t)r].read_size);
    }
    *out_cigars = blob; *out_cigars_bytes = (long long)bytes;
    return 0;
}

extern "C" void mpn_realign_free(void* p) { free(p); }

extern "C" void free_memory(struct_str_arr* pointer, int size)
{
    if (!pointer) return;
    for (int i = 0; i < size && i < 1000; ++i) delete[] pointer->cigar_string[i];
    delete pointer;
}

/* test / A-B hook: the fast pass alone.  which = 0: GPU kernel (with the per-region host path for what it does not take), 1: host
 * k-mer index path for every region.  hap_scores: sum of nhap entries; places: per region nhap * nread pairs (score, pos). */
extern "C" int mpn_realign_fastpass_only(const mpn_region* regions, int nregions, int which, int* hap_scores, int* places, double* kernel_ms)
{
    std::vector<Region> rgs((size_t)std::max(nregions, 0));
    for (int r = 0; r < nregions; ++r) load_region(rgs[(size_t)r], regions[r]);
    if (kernel_ms) *kernel_ms = 0;
    if (which == 0) fast_pass_regions(rgs, kernel_ms);
    else mpn::parallel_for(nregions, 1, [&](int64_t r) { fast_pass(rgs[(size_t)r], false); });
    size_t hs = 0, pl = 0;
    for (const Region& rg : rgs)
        for (const HapRecord& h : rg.haps) {
            hap_scores[hs++] = h.score;
            for (const Placement& p : h.reads) { places[pl++] = p.score; places[pl++] = p.pos; }
        }
    return 0;
}

extern "C" int mpn_realign_last_stats(long long* pairs, long long* cells, double* seconds3)
{
    if (pairs) *pairs = g_last.pairs;
    if (cells) *cells = g_last.cells;
    if (seconds3) { seconds3[0] = g_last.t_fast; seconds3[1] = g_last.t_gpu; seconds3[2] = g_last.t_compose; }
    return 0;
}

extern "C" double mpn_realign_last_fastpass_kernel_ms() { return g_last.fp_kernel_ms; }
