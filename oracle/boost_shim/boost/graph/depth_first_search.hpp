// boost::depth_first_search with named parameters (visitor / root_vertex / vertex_index_map) and dfs_visitor<>, for the stand-in of
// adjacency_list.hpp.  Same event order as Boost's iterative implementation: the root (if given) first, then every still-white vertex in
// vertices(g) order; per vertex the out-edges in container order; tree / back / forward-or-cross classification by colour.  Test infrastructure.
#pragma once
#include "adjacency_list.hpp"
namespace boost {
struct null_visitor {};
template <class Base = null_visitor> struct dfs_visitor {
    template <class V, class G> void initialize_vertex(V, const G&) {}
    template <class V, class G> void start_vertex(V, const G&) {}
    template <class V, class G> void discover_vertex(V, const G&) {}
    template <class E, class G> void examine_edge(E, const G&) {}
    template <class E, class G> void tree_edge(E, const G&) {}
    template <class E, class G> void back_edge(E, const G&) {}
    template <class E, class G> void forward_or_cross_edge(E, const G&) {}
    template <class E, class G> void finish_edge(E, const G&) {}
    template <class V, class G> void finish_vertex(V, const G&) {}
};
template <class Vis> struct dfs_params {
    Vis vis; shim::vertex_handle root; bool has_root = false;
    dfs_params root_vertex(shim::vertex_handle v) const { dfs_params p = *this; p.root = v; p.has_root = true; return p; }
    template <class M> dfs_params vertex_index_map(const M&) const { return *this; }
};
template <class Vis> dfs_params<Vis> visitor(Vis v) { return dfs_params<Vis>{v, shim::vertex_handle{}, false}; }

template <class G, class Vis> void depth_first_search(const G& g, dfs_params<Vis> params)
{
    using V = shim::vertex_handle;
    Vis& vis = params.vis;
    std::map<V, int> colour;                      // 0 white, 1 gray, 2 black
    const std::vector<V> all = shim_vertices(g);
    for (V v : all) { colour[v] = 0; vis.initialize_vertex(v, g); }
    auto visit = [&](V start) {
        using E = decltype(shim_out_edges(start, g));
        struct frame { V v; E edges; std::size_t next; };
        std::vector<frame> stack;
        colour[start] = 1; vis.discover_vertex(start, g);
        stack.push_back(frame{start, shim_out_edges(start, g), 0});
        while (!stack.empty()) {
            frame& f = stack.back();
            if (f.next < f.edges.size()) {
                auto e = f.edges[f.next++];
                vis.examine_edge(e, g);
                const V t = target(e, g);
                const int c = colour[t];
                if (c == 0) {
                    vis.tree_edge(e, g);
                    colour[t] = 1; vis.discover_vertex(t, g);
                    stack.push_back(frame{t, shim_out_edges(t, g), 0});      // (invalidates f)
                } else if (c == 1) vis.back_edge(e, g);
                else vis.forward_or_cross_edge(e, g);
            } else {
                colour[f.v] = 2; vis.finish_vertex(f.v, g);
                stack.pop_back();
            }
        }
    };
    if (params.has_root) { vis.start_vertex(params.root, g); visit(params.root); }
    for (V v : all) if (colour[v] == 0) { vis.start_vertex(v, g); visit(v); }
}
}  // namespace boost
