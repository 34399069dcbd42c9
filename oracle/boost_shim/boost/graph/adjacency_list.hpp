// Minimal stand-in for the part of Boost.Graph that the reference's bin/realignment/realign/debruijn_graph.cpp uses -- TEST INFRASTRUCTURE.
// Boost is not installed in this image, so the reference's window assembler cannot be built as it is; with this directory on the include
// path (`-I oracle/boost_shim`) the UNMODIFIED debruijn_graph.cpp / .h compile into oracle/_ref/debruijn_graph_ref (oracle/Makefile),
// which is then the oracle of the product's assembler.  Nothing here is shipped or linked into the product.
//
// Implemented, with Boost's semantics: adjacency_list<setS, listS, bidirectionalS, VertexProp, EdgeProp> (no parallel edges; out-edges
// ordered by target descriptor), bundled properties g[v] / g[e], add_vertex, add_edge, edge, vertices, adjacent_vertices, out_degree,
// source, target, remove_edge_if, clear_vertex, remove_vertex, graph_traits, const_associative_property_map.
// One documented difference: with listS Boost's vertex descriptors are heap pointers and the out-edge sets are ordered by ADDRESS; here a
// descriptor compares by creation order (what address order is under an allocator that hands out increasing addresses).  The only place
// the order is observable is the moment CandidatePaths trips its 256-path cap (debruijn_graph.cpp:293-299).
#pragma once
#include <algorithm>
#include <cstddef>
#include <cstring>
#include <iterator>
#include <list>
#include <map>
#include <set>
#include <string>
#include <tuple>
#include <utility>
#include <vector>

namespace boost {

struct setS {}; struct listS {}; struct vecS {}; struct bidirectionalS {}; struct directedS {};

namespace shim {
struct vertex_base { std::size_t id; };
// vertex descriptor: a handle to the heap node, ordered by creation
struct vertex_handle {
    vertex_base* p = nullptr;
    bool operator==(const vertex_handle& o) const { return p == o.p; }
    bool operator!=(const vertex_handle& o) const { return p != o.p; }
    bool operator<(const vertex_handle& o) const { return (p ? p->id : 0) < (o.p ? o.p->id : 0); }
};
}  // namespace shim

template <class OutEdgeListS, class VertexListS, class DirectedS, class VertexProp, class EdgeProp>
class adjacency_list {
public:
    using vertex_descriptor = shim::vertex_handle;
    struct edge_rec { EdgeProp prop; };
    struct edge_descriptor {
        vertex_descriptor src, dst; edge_rec* rec = nullptr;
        bool operator==(const edge_descriptor& o) const { return rec == o.rec; }
    };
    struct node : shim::vertex_base {
        VertexProp prop;
        std::map<vertex_descriptor, edge_rec*> out, in;      // keyed by the other end
    };
    struct vertex_iterator {
        using iterator_category = std::forward_iterator_tag; using value_type = vertex_descriptor; using difference_type = std::ptrdiff_t;
        using pointer = const vertex_descriptor*; using reference = vertex_descriptor;
        typename std::list<node*>::const_iterator it;
        vertex_descriptor operator*() const { return vertex_descriptor{*it}; }
        vertex_iterator& operator++() { ++it; return *this; }
        vertex_iterator operator++(int) { vertex_iterator t = *this; ++it; return t; }
        bool operator==(const vertex_iterator& o) const { return it == o.it; }
        bool operator!=(const vertex_iterator& o) const { return it != o.it; }
    };
    struct adjacency_iterator {
        using iterator_category = std::forward_iterator_tag; using value_type = vertex_descriptor; using difference_type = std::ptrdiff_t;
        using pointer = const vertex_descriptor*; using reference = vertex_descriptor;
        typename std::map<vertex_descriptor, edge_rec*>::const_iterator it;
        vertex_descriptor operator*() const { return it->first; }
        adjacency_iterator& operator++() { ++it; return *this; }
        adjacency_iterator operator++(int) { adjacency_iterator t = *this; ++it; return t; }
        bool operator==(const adjacency_iterator& o) const { return it == o.it; }
        bool operator!=(const adjacency_iterator& o) const { return it != o.it; }
    };
    using edge_iterator = int;      // named by the reference's header, never used

    adjacency_list() = default;
    adjacency_list(const adjacency_list&) = delete;
    adjacency_list& operator=(const adjacency_list&) = delete;
    ~adjacency_list()
    {
        for (node* n : nodes_) { for (auto& kv : n->out) delete kv.second; delete n; }
    }
    VertexProp& operator[](vertex_descriptor v) { return static_cast<node*>(v.p)->prop; }
    const VertexProp& operator[](vertex_descriptor v) const { return static_cast<const node*>(v.p)->prop; }
    EdgeProp& operator[](const edge_descriptor& e) { return e.rec->prop; }
    const EdgeProp& operator[](const edge_descriptor& e) const { return e.rec->prop; }

    static node* N(vertex_descriptor v) { return static_cast<node*>(v.p); }
    std::list<node*> nodes_;
    std::size_t next_id_ = 1;
};

template <class G> struct graph_traits {
    using vertex_descriptor = typename G::vertex_descriptor;
    using edge_descriptor = typename G::edge_descriptor;
    using vertex_iterator = typename G::vertex_iterator;
    using edge_iterator = typename G::edge_iterator;
    using adjacency_iterator = typename G::adjacency_iterator;
};

#define SHIM_G adjacency_list<O, V, D, VP, EP>
#define SHIM_T template <class O, class V, class D, class VP, class EP>
SHIM_T typename SHIM_G::vertex_descriptor add_vertex(const VP& p, SHIM_G& g)
{
    auto* n = new typename SHIM_G::node();
    n->id = g.next_id_++; n->prop = p;
    g.nodes_.push_back(n);
    return typename SHIM_G::vertex_descriptor{n};
}
SHIM_T std::pair<typename SHIM_G::vertex_iterator, typename SHIM_G::vertex_iterator> vertices(const SHIM_G& g)
{
    return {typename SHIM_G::vertex_iterator{g.nodes_.begin()}, typename SHIM_G::vertex_iterator{g.nodes_.end()}};
}
SHIM_T std::pair<typename SHIM_G::edge_descriptor, bool> edge(typename SHIM_G::vertex_descriptor u, typename SHIM_G::vertex_descriptor v, const SHIM_G&)
{
    auto* n = SHIM_G::N(u);
    auto it = n->out.find(v);
    if (it == n->out.end()) return {typename SHIM_G::edge_descriptor{}, false};
    return {typename SHIM_G::edge_descriptor{u, v, it->second}, true};
}
SHIM_T std::pair<typename SHIM_G::edge_descriptor, bool> add_edge(typename SHIM_G::vertex_descriptor u, typename SHIM_G::vertex_descriptor v, const EP& p, SHIM_G& g)
{
    auto found = edge(u, v, g);
    if (found.second) return {found.first, false};                // setS: no parallel edges
    auto* r = new typename SHIM_G::edge_rec{p};
    SHIM_G::N(u)->out[v] = r; SHIM_G::N(v)->in[u] = r;
    return {typename SHIM_G::edge_descriptor{u, v, r}, true};
}
SHIM_T std::pair<typename SHIM_G::adjacency_iterator, typename SHIM_G::adjacency_iterator> adjacent_vertices(typename SHIM_G::vertex_descriptor v, const SHIM_G&)
{
    auto* n = SHIM_G::N(v);
    return {typename SHIM_G::adjacency_iterator{n->out.begin()}, typename SHIM_G::adjacency_iterator{n->out.end()}};
}
SHIM_T std::size_t out_degree(typename SHIM_G::vertex_descriptor v, const SHIM_G&) { return SHIM_G::N(v)->out.size(); }
SHIM_T typename SHIM_G::vertex_descriptor source(const typename SHIM_G::edge_descriptor& e, const SHIM_G&) { return e.src; }
SHIM_T typename SHIM_G::vertex_descriptor target(const typename SHIM_G::edge_descriptor& e, const SHIM_G&) { return e.dst; }
template <class Pred, class O, class V, class D, class VP, class EP> void remove_edge_if(Pred pred, SHIM_G& g)
{
    for (auto* n : g.nodes_)
        for (auto it = n->out.begin(); it != n->out.end();) {
            typename SHIM_G::edge_descriptor e{typename SHIM_G::vertex_descriptor{n}, it->first, it->second};
            if (pred(e)) { SHIM_G::N(it->first)->in.erase(typename SHIM_G::vertex_descriptor{n}); delete it->second; it = n->out.erase(it); }
            else ++it;
        }
}
SHIM_T void clear_vertex(typename SHIM_G::vertex_descriptor v, SHIM_G&)
{
    auto* n = SHIM_G::N(v);
    for (auto& kv : n->out) { if (kv.first != v) SHIM_G::N(kv.first)->in.erase(v); delete kv.second; }
    for (auto& kv : n->in) if (kv.first != v) SHIM_G::N(kv.first)->out.erase(v);      // the record of a self loop was deleted above
    for (auto& kv : n->in) if (kv.first != v) delete kv.second;
    n->out.clear(); n->in.clear();
}
SHIM_T void remove_vertex(typename SHIM_G::vertex_descriptor v, SHIM_G& g)
{
    auto* n = SHIM_G::N(v);
    g.nodes_.remove(n);
    delete n;
}
// generic access used by depth_first_search (also implemented for reverse_graph)
SHIM_T std::vector<typename SHIM_G::edge_descriptor> shim_out_edges(typename SHIM_G::vertex_descriptor v, const SHIM_G&)
{
    std::vector<typename SHIM_G::edge_descriptor> r;
    for (auto& kv : SHIM_G::N(v)->out) r.push_back(typename SHIM_G::edge_descriptor{v, kv.first, kv.second});
    return r;
}
SHIM_T std::vector<typename SHIM_G::vertex_descriptor> shim_vertices(const SHIM_G& g)
{
    std::vector<typename SHIM_G::vertex_descriptor> r;
    for (auto* n : g.nodes_) r.push_back(typename SHIM_G::vertex_descriptor{n});
    return r;
}
#undef SHIM_G
#undef SHIM_T

template <class Map> class const_associative_property_map {
public:
    const_associative_property_map() = default;
    explicit const_associative_property_map(const Map& m) : m_(&m) {}
    const Map* m_ = nullptr;
};

}  // namespace boost
