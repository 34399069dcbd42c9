// host_shim.cpp -- TEST INFRASTRUCTURE ONLY.  Never linked into the product.
//
// Lets the `-m "not gpu"` tests exercise the product's HOST logic -- the k-mer fast pass and CIGAR algebra of
// csrc/realign_region.cpp and the '=' / 'X' post-processing of csrc/ssw_cpp_layer.cpp -- on a machine without a GPU, by
// standing in for the one entry point those files call, mpn_align_batch (include/mpn_ssw_batch.h), with the CPU checkers:
// the compiled reference ssw.c when MPN_SHIM_REF points at oracle/_ref/libssw_ref.so, else the scalar restatement.
// oracle/Makefile links this file with the two product sources into oracle/_hosttest/realigner_hosttest.so, which only
// tests/test_realigner_host.py loads.  The product libraries (libmpn_ssw.so, realign/realigner) contain the CUDA engine instead.
#include "../include/mpn_ssw_batch.h"
#include "../megapath-nano_b200/csrc/host_shared.h"
#include <dlfcn.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

extern "C" double oracle_run_batch(void* init, void* align, void* idestroy, void* adestroy,
                        const int8_t* reads, const int64_t* read_off, const int8_t* refs, const int64_t* ref_off, const int32_t* masklen,
                        const int8_t* mat, int32_t n, int32_t gapO, int32_t gapE, int32_t flag, int32_t filters, int32_t filterd, int32_t score_size,
                        int64_t npairs, int32_t threads, int32_t* out, uint32_t* cigar, int32_t cigar_cap);
extern "C" double oracle_run_batch_port(const int8_t* reads, const int64_t* read_off, const int8_t* refs, const int64_t* ref_off, const int32_t* masklen,
                             const int8_t* mat, int32_t n, int32_t gapO, int32_t gapE, int32_t flag, int32_t filters, int32_t filterd, int32_t score_size,
                             int64_t npairs, int32_t threads, int32_t* out, uint32_t* cigar, int32_t cigar_cap);

namespace mpn {
std::mutex& shared_engine_mutex() { static std::mutex mu; return mu; }
mpn_engine* shared_engine_locked() { return reinterpret_cast<mpn_engine*>(0x1); }      // opaque token; the shim has no engine
}

extern "C" int mpn_align_batch(mpn_engine*, const mpn_params* p, const int8_t* reads, const int64_t* read_off, const int8_t* refs,
                               const int64_t* ref_off, const int32_t* masklen, int64_t npairs, mpn_result* out, uint32_t* cigar, int64_t cigar_cap);

// spans -> CSR (the checkers take one pair per call anyway)
extern "C" int mpn_align_batch_spans(mpn_engine* e, const mpn_params* p, const int8_t* seq, int64_t seq_bytes, const int64_t* rd_start, const int32_t* rd_len,
                                     const int64_t* rf_start, const int32_t* rf_len, const int32_t* masklen, int64_t npairs,
                                     mpn_result* out, uint32_t* cigar, int64_t cigar_cap)
{
    std::vector<int64_t> ro(1, 0), fo(1, 0);
    std::vector<int8_t> reads, refs;
    for (int64_t i = 0; i < npairs; ++i) {
        if (rd_start[i] < 0 || rf_start[i] < 0 || rd_start[i] + rd_len[i] > seq_bytes || rf_start[i] + rf_len[i] > seq_bytes) return MPN_E_ARG;
        reads.insert(reads.end(), seq + rd_start[i], seq + rd_start[i] + rd_len[i]);
        refs.insert(refs.end(), seq + rf_start[i], seq + rf_start[i] + rf_len[i]);
        ro.push_back((int64_t)reads.size()); fo.push_back((int64_t)refs.size());
    }
    reads.push_back(0); refs.push_back(0);
    return mpn_align_batch(e, p, reads.data(), ro.data(), refs.data(), fo.data(), masklen, npairs, out, cigar, cigar_cap);
}

extern "C" int mpn_align_batch(mpn_engine*, const mpn_params* p, const int8_t* reads, const int64_t* read_off, const int8_t* refs,
                               const int64_t* ref_off, const int32_t* masklen, int64_t npairs, mpn_result* out, uint32_t* cigar, int64_t cigar_cap)
{
    const int cap = 2048;
    std::vector<int32_t> res((size_t)npairs * 8);
    std::vector<uint32_t> cig((size_t)npairs * cap);
    const char* ref_so = getenv("MPN_SHIM_REF");
    if (ref_so && *ref_so) {
        static void* h = dlopen(ref_so, RTLD_NOW | RTLD_LOCAL);
        if (!h) { fprintf(stderr, "host_shim: cannot load %s\n", ref_so); abort(); }
        oracle_run_batch(dlsym(h, "ssw_init"), dlsym(h, "ssw_align"), dlsym(h, "init_destroy"), dlsym(h, "align_destroy"), reads, read_off, refs, ref_off, masklen,
                         p->mat, p->n, p->gapO, p->gapE, p->flag, p->filters, p->filterd, p->score_size, npairs, 4, res.data(), cig.data(), cap);
    } else {
        oracle_run_batch_port(reads, read_off, refs, ref_off, masklen, p->mat, p->n, p->gapO, p->gapE, p->flag, p->filters, p->filterd, p->score_size,
                              npairs, 4, res.data(), cig.data(), cap);
    }
    int64_t used = 0;
    for (int64_t i = 0; i < npairs; ++i) {
        const int32_t* r = &res[(size_t)i * 8];
        mpn_result& o = out[i];
        memset(&o, 0, sizeof o);
        if (r[7] < 0) { o.status = MPN_ST_NULL; continue; }
        o.score1 = (uint16_t)r[0]; o.score2 = (uint16_t)r[1]; o.ref_begin1 = r[2]; o.ref_end1 = r[3]; o.read_begin1 = r[4]; o.read_end1 = r[5]; o.ref_end2 = r[6];
        o.cigar_len = r[7];
        if (r[7] > 0) {
            if (r[7] > cap || used + r[7] > cigar_cap) return MPN_E_CIGAR_SPACE;
            o.cigar_off = used;
            memcpy(cigar + used, &cig[(size_t)i * cap], sizeof(uint32_t) * (size_t)r[7]);
            used += r[7];
        }
    }
    return 0;
}
