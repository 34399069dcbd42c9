// see adjacency_list.hpp in this directory (stand-in for the part of Boost.Graph the reference's debruijn_graph.cpp uses; test infrastructure)
#pragma once
#include "adjacency_list.hpp"
