// boost::make_reverse_graph for the stand-in of adjacency_list.hpp: the same vertices, every edge turned around.  Test infrastructure.
#pragma once
#include "adjacency_list.hpp"
namespace boost {
template <class G> struct reverse_graph {
    const G& g;
    using vertex_descriptor = typename G::vertex_descriptor;
    struct edge_descriptor { typename G::edge_descriptor under; };
};
template <class G> reverse_graph<G> make_reverse_graph(const G& g) { return reverse_graph<G>{g}; }
template <class G> typename G::vertex_descriptor source(const typename reverse_graph<G>::edge_descriptor& e, const reverse_graph<G>&) { return e.under.dst; }
template <class G> typename G::vertex_descriptor target(const typename reverse_graph<G>::edge_descriptor& e, const reverse_graph<G>&) { return e.under.src; }
template <class G> std::vector<typename reverse_graph<G>::edge_descriptor> shim_out_edges(typename G::vertex_descriptor v, const reverse_graph<G>&)
{
    std::vector<typename reverse_graph<G>::edge_descriptor> r;
    for (auto& kv : G::N(v)->in) r.push_back(typename reverse_graph<G>::edge_descriptor{typename G::edge_descriptor{kv.first, v, kv.second}});
    return r;
}
template <class G> std::vector<typename G::vertex_descriptor> shim_vertices(const reverse_graph<G>& rg) { return shim_vertices(rg.g); }
}  // namespace boost
