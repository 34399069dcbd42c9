// boost::split(container, input, boost::is_any_of(chars)) with the default token_compress_off: every delimiter ends a token, so adjacent
// delimiters and an empty input give empty strings (Boost's behaviour, which the reference's get_consensus relies on for the per-read
// low-quality lists, debruijn_graph.cpp:394-400).  Test infrastructure.
#pragma once
#include <string>
#include <vector>
namespace boost {
struct shim_any_of { std::string chars; };
inline shim_any_of is_any_of(const char* c) { return shim_any_of{c}; }
template <class Out> void split(Out& out, const std::string& in, const shim_any_of& pred)
{
    out.clear();
    std::string cur;
    for (char c : in) {
        if (pred.chars.find(c) != std::string::npos) { out.push_back(cur); cur.clear(); }
        else cur.push_back(c);
    }
    out.push_back(cur);
}
template <class Out> void split(Out& out, const char* in, const shim_any_of& pred) { split(out, std::string(in), pred); }
template <class Out> void split(Out& out, char* in, const shim_any_of& pred) { split(out, std::string(in), pred); }
}  // namespace boost
