// DeBruijnGraph::GraphViz (debruijn_graph.cpp:338-351) is never called on the assembler's path; these declarations only let it compile.
// Test infrastructure.
#pragma once
#include "adjacency_list.hpp"
#include <ostream>
namespace boost {
struct default_writer { template <class... A> void operator()(A&&...) const {} };
template <class T> struct shim_label_writer { T v; };
template <class T> shim_label_writer<T> make_label_writer(T v) { return shim_label_writer<T>{v}; }
template <class M, class G> int get(M, const G&) { return 0; }
template <class G, class VW, class EW, class GW, class IM> void write_graphviz(std::ostream&, const G&, VW, EW, GW, IM) {}
}  // namespace boost
