// Window haplotype assembly (include/debruijn_graph.h): the de Bruijn step that feeds the realigner its candidate haplotypes.
// Behaviour follows the reference's debruijn_graph.cpp (DeepVariant's assembler on Boost.Graph); the structure does not:
//   * k-mers are never copied -- a vertex is a pointer into the caller's text, interned through a rolling hash + exact compare;
//   * vertices / edges are flat arrays, edges are interned by (from, to) in an open-addressing table;
//   * the cycle test is a topological peel (Kahn) instead of a DFS with a back-edge visitor;
//   * paths are parent-linked nodes in an arena, not copied vectors.
// Host only: a window has a few thousand k-mers; many windows run in parallel on the host threads (mpn_dbg_consensus_packed).
#include "../../include/debruijn_graph.h"
#include "host_shared.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <string>
#include <vector>

namespace {

constexpr int kMinK = 10, kMaxK = 101;          // KMinMaxFromReference, debruijn_graph.cpp:187-189
constexpr int kMaxLivePaths = 256;              // CandidatePaths, debruijn_graph.cpp:287-289
constexpr int kMinReadEdgeWeight = 2;           // Prune, debruijn_graph.cpp:345-349

struct View { const char* p; int n; };

inline uint64_t mix64(uint64_t k) { k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33; return k; }

// polynomial hash of k characters with O(1) slide
struct Roller {
    static constexpr uint64_t B = 0x9e3779b97f4a7c15ULL | 1ULL;
    int k; uint64_t top = 1;                     // B^(k-1)
    explicit Roller(int k_) : k(k_) { for (int i = 1; i < k; ++i) top *= B; }
    uint64_t start(const char* s) const { uint64_t h = 0; for (int i = 0; i < k; ++i) h = h * B + (uint8_t)s[i]; return h; }
    uint64_t slide(uint64_t h, const char* s) const { return (h - (uint8_t)s[0] * top) * B + (uint8_t)s[k]; }   // s = old window start
};

class Graph {
public:
    explicit Graph(int k) : k_(k), roll_(k) { slots_.assign(1024, 0); vmask_ = 1023; eslots_.assign(2048, 0); emask_ = 2047; }

    // AddKmersAndEdges, debruijn_graph.cpp:254-264: vertices for the k-mers starting at start..end, an edge between neighbours.
    // The vertex at `start` is made as soon as end > 0, even when end <= start (the reference's condition, kept).
    void add_run(const char* s, int start, int end, bool is_ref)
    {
        if (end <= 0) return;
        uint64_t h = roll_.start(s + start);
        int prev = ensure(s + start, h);
        for (int i = start + 1; i <= end; ++i) {
            h = roll_.slide(h, s + i - 1);
            const int cur = ensure(s + i, h);
            add_edge(prev, cur, is_ref);
            prev = cur;
        }
    }

    int find(const char* kmer) const
    {
        const uint64_t h = roll_.start(kmer);
        for (uint64_t q = mix64(h) & vmask_;; q = (q + 1) & vmask_) {
            const int id = slots_[q] - 1;
            if (id < 0) return -1;
            if (vhash_[(size_t)id] == h && memcmp(vkmer_[(size_t)id], kmer, (size_t)k_) == 0) return id;
        }
    }

    // HasCycle, debruijn_graph.cpp:153-159: any directed cycle (self loops included), over ALL edges, before pruning
    bool has_cycle() const
    {
        const int nv = (int)vkmer_.size();
        std::vector<int> indeg((size_t)nv, 0), first((size_t)nv + 1, 0), to(edges_.size());
        for (const Edge& e : edges_) { indeg[(size_t)e.to]++; first[(size_t)e.from + 1]++; }
        for (int v = 0; v < nv; ++v) first[(size_t)v + 1] += first[(size_t)v];
        std::vector<int> fill(first.begin(), first.end() - 1);
        for (const Edge& e : edges_) to[(size_t)fill[(size_t)e.from]++] = e.to;
        std::vector<int> stack;
        for (int v = 0; v < nv; ++v) if (indeg[(size_t)v] == 0) stack.push_back(v);
        int peeled = 0;
        while (!stack.empty()) {
            const int v = stack.back(); stack.pop_back(); ++peeled;
            for (int q = first[(size_t)v]; q < first[(size_t)v + 1]; ++q) if (--indeg[(size_t)to[(size_t)q]] == 0) stack.push_back(to[(size_t)q]);
        }
        return peeled != nv;
    }

    // Prune + CandidatePaths + HaplotypeForPath (debruijn_graph.cpp:343-376, :279-310, :312-321).  Unsorted; empty if too many paths.
    std::vector<std::string> haplotypes(int source, int sink) const
    {
        const int nv = (int)vkmer_.size();
        // edges that survive: reference edges and edges seen at least twice
        std::vector<int> ofirst((size_t)nv + 1, 0), ifirst((size_t)nv + 1, 0);
        for (const Edge& e : edges_) if (e.is_ref || e.weight >= kMinReadEdgeWeight) { ofirst[(size_t)e.from + 1]++; ifirst[(size_t)e.to + 1]++; }
        for (int v = 0; v < nv; ++v) { ofirst[(size_t)v + 1] += ofirst[(size_t)v]; ifirst[(size_t)v + 1] += ifirst[(size_t)v]; }
        std::vector<int> oto((size_t)ofirst[(size_t)nv]), ifrom((size_t)ifirst[(size_t)nv]);
        {
            std::vector<int> of(ofirst.begin(), ofirst.end() - 1), inf(ifirst.begin(), ifirst.end() - 1);
            for (const Edge& e : edges_) if (e.is_ref || e.weight >= kMinReadEdgeWeight) { oto[(size_t)of[(size_t)e.from]++] = e.to; ifrom[(size_t)inf[(size_t)e.to]++] = e.from; }
        }
        // vertices reachable forward from the source AND backward from the sink stay
        auto reach = [nv](int root, const std::vector<int>& first, const std::vector<int>& adj) {
            std::vector<uint8_t> seen((size_t)nv, 0);
            std::vector<int> stack{root};
            seen[(size_t)root] = 1;
            while (!stack.empty()) {
                const int v = stack.back(); stack.pop_back();
                for (int q = first[(size_t)v]; q < first[(size_t)v + 1]; ++q) { const int w = adj[(size_t)q]; if (!seen[(size_t)w]) { seen[(size_t)w] = 1; stack.push_back(w); } }
            }
            return seen;
        };
        const std::vector<uint8_t> fwd = reach(source, ofirst, oto), bwd = reach(sink, ifirst, ifrom);
        // successors among the kept vertices, ascending vertex id = order of first appearance.  The reference iterates a std::set of heap
        // node ADDRESSES (Boost listS descriptors), i.e. creation order under an allocator that hands out increasing addresses.  The output
        // is sorted, so the order is observable in one case only: the count tested before the last pop is (all paths) - (successors of the
        // last popped vertex - 1), so when the total exceeds 256 by less than that, the visiting order decides between "every path" and
        // "none" (debruijn_graph.cpp:293-299) -- i.e. only when the last branching sits right before the sink / a dead end.
        std::vector<int> kfirst((size_t)nv + 1, 0), kto;
        kto.reserve(oto.size());
        for (int v = 0; v < nv; ++v) {
            if (fwd[(size_t)v] && bwd[(size_t)v])
                for (int q = ofirst[(size_t)v]; q < ofirst[(size_t)v + 1]; ++q) { const int w = oto[(size_t)q]; if (fwd[(size_t)w] && bwd[(size_t)w]) kto.push_back(w); }
            kfirst[(size_t)v + 1] = (int)kto.size();
            std::sort(kto.begin() + kfirst[(size_t)v], kto.end());
        }
        // breadth-first path extension from the source; a path ends at the sink or at a vertex without successors
        struct Node { int v, parent; };
        std::vector<Node> nodes{Node{source, -1}};
        std::deque<int> open{0};
        std::vector<int> done;
        while (!open.empty()) {
            if ((int)(done.size() + open.size()) > kMaxLivePaths) return {};
            const int at = open.front(); open.pop_front();
            const int last = nodes[(size_t)at].v;
            for (int q = kfirst[(size_t)last]; q < kfirst[(size_t)last + 1]; ++q) {
                const int w = kto[(size_t)q];
                nodes.push_back(Node{w, at});
                const int id = (int)nodes.size() - 1;
                if (w == sink || kfirst[(size_t)w] == kfirst[(size_t)w + 1]) done.push_back(id); else open.push_back(id);
            }
        }
        std::vector<std::string> out;
        out.reserve(done.size());
        for (int id : done) {
            std::string h;
            for (int n = id; n >= 0; n = nodes[(size_t)n].parent) h.push_back(vkmer_[(size_t)nodes[(size_t)n].v][0]);
            std::reverse(h.begin(), h.end());
            h.append(vkmer_[(size_t)nodes[(size_t)id].v] + 1, (size_t)k_ - 1);
            out.push_back(std::move(h));
        }
        return out;
    }

private:
    struct Edge { int from, to, weight; bool is_ref; };

    int ensure(const char* kmer, uint64_t h)         // EnsureVertex, debruijn_graph.cpp:112-125
    {
        uint64_t q = mix64(h) & vmask_;
        for (;; q = (q + 1) & vmask_) {
            const int id = slots_[q] - 1;
            if (id < 0) break;
            if (vhash_[(size_t)id] == h && memcmp(vkmer_[(size_t)id], kmer, (size_t)k_) == 0) return id;
        }
        vkmer_.push_back(kmer); vhash_.push_back(h);
        slots_[q] = (int)vkmer_.size();
        if (vkmer_.size() * 2 > slots_.size()) {
            std::vector<int> bigger(slots_.size() * 2, 0);
            vmask_ = bigger.size() - 1;
            for (size_t id = 0; id < vkmer_.size(); ++id) { uint64_t r = mix64(vhash_[id]) & vmask_; while (bigger[r]) r = (r + 1) & vmask_; bigger[r] = (int)id + 1; }
            slots_.swap(bigger);
        }
        return (int)vkmer_.size() - 1;
    }

    void add_edge(int from, int to, bool is_ref)    // AddEdge, debruijn_graph.cpp:240-252: one edge per (from, to), weight = times seen
    {
        const uint64_t key = ((uint64_t)(uint32_t)from << 32) | (uint32_t)to;
        uint64_t q = mix64(key) & emask_;
        for (;; q = (q + 1) & emask_) {
            const int id = eslots_[q] - 1;
            if (id < 0) break;
            Edge& e = edges_[(size_t)id];
            if (e.from == from && e.to == to) { e.weight++; e.is_ref = e.is_ref || is_ref; return; }
        }
        edges_.push_back(Edge{from, to, 1, is_ref});
        eslots_[q] = (int)edges_.size();
        if (edges_.size() * 2 > eslots_.size()) {
            std::vector<int> bigger(eslots_.size() * 2, 0);
            emask_ = bigger.size() - 1;
            for (size_t id = 0; id < edges_.size(); ++id) {
                uint64_t r = mix64(((uint64_t)(uint32_t)edges_[id].from << 32) | (uint32_t)edges_[id].to) & emask_;
                while (bigger[r]) r = (r + 1) & emask_;
                bigger[r] = (int)id + 1;
            }
            eslots_.swap(bigger);
        }
    }

    int k_;
    Roller roll_;
    std::vector<const char*> vkmer_;
    std::vector<uint64_t> vhash_;
    std::vector<int> slots_, eslots_;
    uint64_t vmask_, emask_;
    std::vector<Edge> edges_;
};

// smallest k in [10, min(101, len-1)] for which the reference itself has no repeated k-mer (KMinMaxFromReference, :187-209); -1: none
int min_k_of_reference(View ref, int* max_k)
{
    *max_k = std::min(kMaxK, ref.n - 1);
    for (int k = kMinK; k <= *max_k; ++k) {
        // sort the window hashes, compare the characters of equal-hash neighbours
        const Roller r(k);
        const int last = ref.n - k;
        std::vector<std::pair<uint64_t, int>> hs;
        hs.reserve((size_t)last + 1);
        uint64_t h = r.start(ref.p);
        for (int i = 0;; ++i) { hs.push_back({h, i}); if (i == last) break; h = r.slide(h, ref.p + i); }
        std::sort(hs.begin(), hs.end());
        bool repeat = false;
        for (size_t a = 0; a + 1 < hs.size() && !repeat; ++a)
            for (size_t b = a + 1; b < hs.size() && hs[b].first == hs[a].first; ++b)
                if (memcmp(ref.p + hs[a].second, ref.p + hs[b].second, (size_t)k) == 0) { repeat = true; break; }
        if (!repeat) return k;
    }
    return -1;
}

// Build, debruijn_graph.cpp:212-238.  lowq[r] = sorted low-quality positions of read r.
std::vector<std::string> assemble(View ref, const std::vector<View>& reads, const std::vector<std::vector<int>>& lowq, int* used_k)
{
    *used_k = 0;
    int max_k = 0;
    const int min_k = min_k_of_reference(ref, &max_k);
    if (min_k < 0) return {};
    // next position at or after i that cannot be inside a k-mer: not A/C/G/T or flagged low quality (NextBadPosition, :270-282)
    std::vector<std::vector<int>> nextbad(reads.size());
    for (size_t r = 0; r < reads.size(); ++r) {
        const View rd = reads[r];
        std::vector<int>& nb = nextbad[r];
        nb.assign((size_t)rd.n + 1, rd.n);
        std::vector<uint8_t> bad((size_t)rd.n, 0);
        for (int i = 0; i < rd.n; ++i) { const char c = rd.p[i]; bad[(size_t)i] = !(c == 'A' || c == 'C' || c == 'G' || c == 'T'); }
        if (r < lowq.size()) for (int p : lowq[r]) if (p >= 0 && p < rd.n) bad[(size_t)p] = 1;
        for (int i = rd.n - 1; i >= 0; --i) nb[(size_t)i] = bad[(size_t)i] ? i : nb[(size_t)i + 1];
    }
    for (int k = min_k; k <= max_k; ++k) {
        Graph g(k);
        g.add_run(ref.p, 0, ref.n - k, true);                                  // AddEdgesForReference, :266-268
        const int source = g.find(ref.p), sink = g.find(ref.p + ref.n - k);      // :170-171
        for (size_t r = 0; r < reads.size(); ++r) {                              // AddEdgesForRead, :271-296
            const View rd = reads[r];
            const int stop = rd.n - k;
            for (int i = 0; i < stop;) {
                const int bad = nextbad[r][(size_t)i];
                g.add_run(rd.p, i, bad - k, false);
                i = bad + 1;
            }
        }
        if (g.has_cycle()) continue;
        std::vector<std::string> haps = g.haplotypes(source, sink);
        std::sort(haps.begin(), haps.end());
        *used_k = k;
        return haps;
    }
    return {};
}

// the reference splits with boost::split(..., is_any_of(",")): empty fields are kept, an empty string is one empty field
std::vector<View> split_commas(const char* s)
{
    std::vector<View> out;
    const char* at = s;
    for (;;) {
        const char* c = strchr(at, ',');
        if (!c) { out.push_back(View{at, (int)strlen(at)}); break; }
        out.push_back(View{at, (int)(c - at)});
        at = c + 1;
    }
    return out;
}

// `stringstream >> int` until it fails (debruijn_graph.cpp:407-416)
std::vector<int> parse_positions(View field)
{
    std::vector<int> out;
    std::string s(field.p, (size_t)field.n);
    const char* at = s.c_str();
    for (;;) {
        char* end = nullptr;
        const long v = strtol(at, &end, 10);
        if (end == at) break;
        out.push_back((int)v);
        at = end;
    }
    std::sort(out.begin(), out.end());
    return out;
}

std::vector<std::string> consensus_of(const char* reference, const char* c_reads, const char* c_base_quality, int* used_k)
{
    const View ref{reference, (int)strlen(reference)};
    const std::vector<View> reads = split_commas(c_reads);
    std::vector<std::vector<int>> lowq;
    for (const View& f : split_commas(c_base_quality)) lowq.push_back(parse_positions(f));
    return assemble(ref, reads, lowq, used_k);
}

}  // namespace

extern "C" dbg_str_arr* get_consensus(char* reference, char* c_reads, char* c_base_quality, int /*read_size*/)
{
    int used_k = 0;
    const std::vector<std::string> haps = consensus_of(reference, c_reads, c_base_quality, &used_k);
    dbg_str_arr* res = new dbg_str_arr();
    res->consensus_size = (int)std::min<size_t>(haps.size(), 500);            // the array has 500 slots (debruijn_graph.h:40-44)
    for (int i = 0; i < res->consensus_size; ++i) {
        res->consensus[i] = new char[haps[(size_t)i].size() + 1];
        memcpy(res->consensus[i], haps[(size_t)i].c_str(), haps[(size_t)i].size() + 1);
    }
    return res;
}

extern "C" void free_memory(dbg_str_arr* pointer, int size)
{
    if (!pointer) return;
    for (int i = 0; i < size && i < 500; ++i) delete[] pointer->consensus[i];
    delete pointer;
}

extern "C" int mpn_dbg_consensus_packed(const char* text, long long text_bytes, int nwindows, int* out_counts, int* out_k,
                                        char** out, long long* out_bytes)
{
    if (nwindows < 0 || !out || !out_bytes || (nwindows > 0 && (!text || !out_counts))) return -1;
    std::vector<const char*> field((size_t)nwindows * 3);
    {
        const char* at = text; const char* const end = text + text_bytes;
        for (size_t f = 0; f < field.size(); ++f) {
            if (at >= end) return -1;
            const void* z = memchr(at, 0, (size_t)(end - at));
            if (!z) return -1;
            field[f] = at;
            at = (const char*)z + 1;
        }
    }
    std::vector<std::vector<std::string>> res((size_t)nwindows);
    std::vector<int> ks((size_t)nwindows, 0);
    mpn::parallel_for(nwindows, 1, [&](int64_t w) {
        res[(size_t)w] = consensus_of(field[(size_t)w * 3], field[(size_t)w * 3 + 1], field[(size_t)w * 3 + 2], &ks[(size_t)w]);
    });
    size_t bytes = 0;
    for (int w = 0; w < nwindows; ++w) { out_counts[w] = (int)res[(size_t)w].size(); if (out_k) out_k[w] = ks[(size_t)w]; for (const std::string& h : res[(size_t)w]) bytes += h.size() + 1; }
    char* buf = (char*)malloc(bytes ? bytes : 1);
    if (!buf) return -1;
    size_t at = 0;
    for (const auto& hs : res) for (const std::string& h : hs) { memcpy(buf + at, h.c_str(), h.size() + 1); at += h.size() + 1; }
    *out = buf; *out_bytes = (long long)bytes;
    return 0;
}

extern "C" void mpn_dbg_free(void* p) { free(p); }
