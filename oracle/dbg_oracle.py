"""CPU restatement of the reference's window haplotype assembler.  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: the reference's debruijn_graph.cpp needs Boost.Graph, which this image does not have (no boost headers anywhere),
so it cannot be compiled into oracle/_ref and the reference holds no tests or golden vectors for it.  This file is a literal,
string-keyed restatement of /root/reference/bin/realignment/realign/debruijn_graph.cpp, written independently of the product's
implementation (plain dicts and copied path lists, DFS colouring for the cycle test); the product is checked against it.

One reference behaviour is not deterministic and therefore not restated bit for bit: Boost iterates a vertex's successors in the
order of heap addresses (`setS` edge container over `listS` vertex descriptors, debruijn_graph.h:61-66).  The output is sorted
(debruijn_graph.cpp:232), so that order can only change at which moment the 256-live-paths cap (:287-289) trips in windows that
hover around the cap.  Successors are visited here in order of vertex creation, which is what ascending heap addresses give.
"""


def k_bounds(ref):
    """KMinMaxFromReference, debruijn_graph.cpp:187-209: (min_k, max_k); min_k = -1 when every k repeats a k-mer of the reference."""
    max_k = min(101, len(ref) - 1)
    for k in range(10, max_k + 1):
        seen = set()
        ok = True
        for i in range(len(ref) - k + 1):
            kmer = ref[i:i + k]
            if kmer in seen:
                ok = False
                break
            seen.add(kmer)
        if ok:
            return k, max_k
    return -1, max_k


class Graph:
    """DeBruijnGraph, debruijn_graph.cpp:161-178"""

    def __init__(self, ref, reads, lowq, k):
        self.k = k
        self.vertex = {}                 # kmer -> id (creation order), kmer_to_vertex_
        self.kmers = []
        self.edges = {}                  # (from, to) -> [weight, is_ref]
        self.add_run(ref, 0, len(ref) - k, True)                                   # AddEdgesForReference :266-268
        self.source = self.vertex[ref[:k]]
        self.sink = self.vertex[ref[len(ref) - k:]]
        for i, read in enumerate(reads):
            self.add_read(read, lowq[i] if i < len(lowq) else set())

    def ensure(self, kmer):              # EnsureVertex :112-125
        v = self.vertex.get(kmer)
        if v is None:
            v = len(self.kmers)
            self.vertex[kmer] = v
            self.kmers.append(kmer)
        return v

    def add_edge(self, a, b, is_ref):    # AddEdge :240-252
        e = self.edges.setdefault((a, b), [0, False])
        e[0] += 1
        e[1] = e[1] or is_ref

    def add_run(self, bases, start, end, is_ref):      # AddKmersAndEdges :254-264
        if end > 0:
            prev = self.ensure(bases[start:start + self.k])
            for i in range(start + 1, end + 1):
                cur = self.ensure(bases[i:i + self.k])
                self.add_edge(prev, cur, is_ref)
                prev = cur

    def add_read(self, bases, low):      # AddEdgesForRead :271-296
        def next_bad(start):
            for i in range(start, len(bases)):
                if bases[i] not in "ACGT" or i in low:
                    return i
            return len(bases)
        stop = len(bases) - self.k
        i = 0
        while i < stop:
            nb = next_bad(i)
            self.add_run(bases, i, nb - self.k, False)
            i = nb + 1

    def successors(self):
        succ = {v: [] for v in range(len(self.kmers))}
        for (a, b) in self.edges:
            succ[a].append(b)
        for v in succ:
            succ[v].sort()
        return succ

    def has_cycle(self):                 # HasCycle :153-159: depth-first search, a back edge (grey target) is a cycle
        succ = self.successors()
        colour = [0] * len(self.kmers)
        for root in range(len(self.kmers)):
            if colour[root]:
                continue
            colour[root] = 1
            stack = [(root, iter(succ[root]))]
            while stack:
                v, it = stack[-1]
                w = next(it, None)
                if w is None:
                    colour[v] = 2
                    stack.pop()
                elif colour[w] == 1:
                    return True
                elif colour[w] == 0:
                    colour[w] = 1
                    stack.append((w, iter(succ[w])))
        return False

    def prune(self):                     # Prune :343-376
        self.edges = {e: wr for e, wr in self.edges.items() if wr[1] or wr[0] >= 2}

        def reach(root, nbrs):
            seen = {root}
            todo = [root]
            while todo:
                v = todo.pop()
                for w in nbrs.get(v, ()):
                    if w not in seen:
                        seen.add(w)
                        todo.append(w)
            return seen
        fwd, bwd = {}, {}
        for (a, b) in self.edges:
            fwd.setdefault(a, []).append(b)
            bwd.setdefault(b, []).append(a)
        keep = reach(self.source, fwd) & reach(self.sink, bwd)
        self.keep = keep
        self.edges = {(a, b): wr for (a, b), wr in self.edges.items() if a in keep and b in keep}

    def candidate_paths(self):           # CandidatePaths :279-310
        succ = {}
        for (a, b) in self.edges:
            succ.setdefault(a, []).append(b)
        for v in succ:
            succ[v].sort()
        terminated, extendable = [], [[self.source]]
        while extendable:
            if len(terminated) + len(extendable) > 256:
                return []
            path = extendable.pop(0)
            for w in succ.get(path[-1], ()):
                ext = path + [w]
                if w == self.sink or not succ.get(w):
                    terminated.append(ext)
                else:
                    extendable.append(ext)
        return terminated

    def haplotype(self, path):           # HaplotypeForPath :312-321
        return "".join(self.kmers[v][0] for v in path) + self.kmers[path[-1]][1:]


def build(ref, reads, lowq):
    """DeBruijnGraph::Build :212-238 -> (sorted haplotypes, k used or 0)"""
    min_k, max_k = k_bounds(ref)
    if min_k == -1:
        return [], 0
    for k in range(min_k, max_k + 1):
        g = Graph(ref, reads, lowq, k)
        if g.has_cycle():
            continue
        g.prune()
        return sorted(g.haplotype(p) for p in g.candidate_paths()), k
    return [], 0


def get_consensus(reference, c_reads, c_base_quality):
    """get_consensus :388-426 with its string formats: reads joined by ',', per-read low-quality positions joined by ' ' then ','"""
    reads = c_reads.split(",")
    lowq = []
    for field in c_base_quality.split(","):
        s = set()
        for tok in field.split():
            try:
                s.add(int(tok))
            except ValueError:
                break
        lowq.append(s)
    return build(reference, reads, lowq)
