/*
 * realigner.h -- C ABI of the region realigner on the B200 Smith-Waterman engine.
 *
 * Drop-in for the `realigner` shared object MegaPath-Nano builds from ssw_cpp.cpp + ssw.c + realigner.cpp
 * (reference README.md:45, build.sh:12) and loads with ctypes at realign_illumina_reads.py:29,32:
 *   realign_reads   reference realigner.cpp:854-859 (ReAligner::realign_reads :782-850)
 *   free_memory     reference realigner.cpp:861-869
 *   struct_str_arr  reference realigner.h:42-46  <->  `StructPointer` in realign_illumina_reads.py:40-43
 * Same names, same argument order, same ownership (the caller releases the result with free_memory).
 *
 * What changes underneath: the (haplotype, reference) and (read, haplotype) Smith-Waterman calls that the reference
 * issues one at a time (realigner.cpp:325-349 and :351-384) are collected per region -- or across many regions with
 * mpn_realign_regions -- and submitted as ONE batch to the CUDA kernels (include/mpn_ssw_batch.h).  The k-mer fast pass
 * and the CIGAR algebra stay on the host.  No CPU alignment code exists in this library.
 */
#ifndef MPN_REALIGNER_H
#define MPN_REALIGNER_H
#ifdef __cplusplus
extern "C" {
#endif

/* result of one region: new position and CIGAR per read, in input order; at most 1000 reads per region (reference realigner.h:42-46) */
typedef struct struct_str_arr {
    int position[1000];
    char* cigar_string[1000];
} struct_str_arr;

/*
 * seqs / positions / cigars   the reads of the region (ASCII bases, current position, current CIGAR), read_size of them
 * reference                   reference window, NUL terminated
 * haplotypes                  candidate haplotypes separated by white space
 * ref_start                   position of reference[0] on the chromosome
 * ref_prefix / ref_suffix     flanks added around the window (not required to be covered by reads, realigner.cpp:248-252)
 */
struct_str_arr* realign_reads(char* seqs[], int* positions, char* cigars[], char* reference, char* haplotypes,
                              int ref_start, int ref_prefix, int ref_suffix, int read_size);
void free_memory(struct_str_arr* pointer, int size);

/* NEW: many regions, one GPU submit.  out[r] receives what realign_reads would return for regions[r]. */
typedef struct {
    char** seqs; int* positions; char** cigars; int read_size;
    char* reference; char* haplotypes;
    int ref_start, ref_prefix, ref_suffix;
} mpn_region;
int mpn_realign_regions(const mpn_region* regions, int nregions, struct_str_arr** out);

/*
 * NEW: the same with flat buffers, for bindings where building arrays of C strings is the bottleneck (Python: one join instead of 10^5
 * c_char_p objects).  `text` holds NUL-terminated strings back to back, per region: reference, haplotypes (white-space separated, one
 * string), then its read sequences, then its read CIGARs.  region_reads[r] = number of reads; region_geom[3*r ..] = ref_start, ref_prefix,
 * ref_suffix; positions = current read positions of all regions back to back.  Results: out_positions (same layout as positions) and
 * *out_cigars = one malloc'ed buffer of NUL-terminated CIGAR strings in read order (release with mpn_realign_free).
 */
int mpn_realign_regions_packed(const char* text, long long text_bytes, int nregions, const int* region_reads, const int* region_geom,
                               const int* positions, int* out_positions, char** out_cigars, long long* out_cigars_bytes);
void mpn_realign_free(void* p);

/* counters of the last realign_reads / mpn_realign_regions call of this thread's process: Smith-Waterman pairs submitted,
 * forward-matrix cells, and seconds spent in {host k-mer pass, GPU batch (copies included), host CIGAR algebra} */
int mpn_realign_last_stats(long long* pairs, long long* cells, double* seconds3);


/* the fast pass alone (test and A/B hook): which = 0 the GPU kernel mpn_fastpass (include/mpn_ssw_batch.h), 1 the host k-mer index.
 * hap_scores: one int per haplotype of every region; places: per region nhap * nread pairs (score, pos), pos -1 = not placed. */
int mpn_realign_fastpass_only(const mpn_region* regions, int nregions, int which, int* hap_scores, int* places, double* kernel_ms);
/* device milliseconds of the fast-pass kernel inside the last realign_reads / mpn_realign_regions call */
double mpn_realign_last_fastpass_kernel_ms(void);

#ifdef __cplusplus
}
#endif
#endif
