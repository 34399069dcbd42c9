/*
 * debruijn_graph.h -- C ABI of the window haplotype assembler.
 *
 * Drop-in for the `debruijn_graph` shared object MegaPath-Nano builds from debruijn_graph.cpp (Boost.Graph) and loads with
 * ctypes at realign_illumina_reads.py:30,33:
 *   get_consensus   reference debruijn_graph.cpp:388-426   (DeBruijnGraph::Build :212-238)
 *   free_memory     reference debruijn_graph.cpp:430-437
 *   struct          reference debruijn_graph.h:40-44 `struct_str_arr` <-> `DBGPointer` in realign_illumina_reads.py:46-48
 * Same names, argument order, string formats and ownership.  Built as megapath-nano_b200/realign/debruijn_graph (no suffix,
 * like the reference).  It is a separate shared object because the realigner exports a different `free_memory`.
 *
 * Host code (C++17, no Boost): this is the step BEFORE the Smith-Waterman hot path -- it produces the candidate haplotypes
 * that realign_reads (include/realigner.h) aligns against.  Graph work of a few thousand vertices per window; no GPU part.
 */
#ifndef MPN_DEBRUIJN_GRAPH_H
#define MPN_DEBRUIJN_GRAPH_H
#ifdef __cplusplus
extern "C" {
#endif

/* reference debruijn_graph.h:40-44 (there the tag is struct_str_arr, which collides with realigner.h's type of the same name) */
typedef struct dbg_str_arr {
    int consensus_size;
    char* consensus[500];
} dbg_str_arr;

/*
 * reference        the window's reference bases, NUL terminated
 * c_reads          the reads, separated by ','
 * c_base_quality   per read (separated by ','): the 0-based read positions of low-quality bases, separated by white space
 * read_size        number of reads (unused, as in the reference)
 * Returns the candidate haplotypes sorted bytewise; consensus_size == 0 when no k in [10, min(101, len-1)] gives an acyclic
 * graph or when more than 256 paths are alive (debruijn_graph.cpp:287-289).  Release with free_memory(p, p->consensus_size).
 */
dbg_str_arr* get_consensus(char* reference, char* c_reads, char* c_base_quality, int read_size);
void free_memory(dbg_str_arr* pointer, int size);

/*
 * NEW: many windows in one call, spread over the host threads.  `text` holds, per window, three NUL-terminated strings back
 * to back with exactly the formats above (reference, reads, low-quality positions).  out_counts[w] = number of haplotypes of
 * window w, *out_k[w] (if out_k != NULL) = the k that was used (0: none); *out = one malloc'ed buffer with all haplotypes as
 * NUL-terminated strings in window order (release with mpn_dbg_free).  Returns 0, or -1 on malformed input.
 */
int mpn_dbg_consensus_packed(const char* text, long long text_bytes, int nwindows, int* out_counts, int* out_k,
                             char** out, long long* out_bytes);
void mpn_dbg_free(void* p);

#ifdef __cplusplus
}
#endif
#endif
