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

// ---- stand-in for mpn_fastpass (the GPU k-mer fast pass): the same "diagonal run" formulation, scalar and character-exact.
// It lets the CPU tier check that formulation end to end against the compiled reference realigner (which uses a k-mer hash index).
extern "C" int mpn_fastpass(mpn_engine*, const char* text, int64_t, const int64_t* hap_start, const int32_t* hap_len, const uint8_t* hap_is_ref, int32_t nhaps,
                            const int64_t* read_start, const int32_t* read_len, int32_t nreads, const mpn_fp_region* regions, int32_t nregions,
                            mpn_placement* places, int32_t* hap_score, uint8_t* region_flag)
{
    (void)nhaps; (void)nreads;
    const int K = 32;
    for (int g = 0; g < nregions; ++g) {
        const mpn_fp_region& rg = regions[g];
        region_flag[g] = 0;
        mpn::parallel_for(rg.nhap, 1, [&](int64_t h) {
            const char* hap = text + hap_start[rg.hap_first + h];
            const int L = hap_len[rg.hap_first + h];
            mpn_placement* out = places + rg.place_first + h * rg.nread;
            std::vector<int> cov((size_t)std::max(L, 1), 1 << 30);
            std::vector<char> occ((size_t)std::max(L, 1), 0);
            long long total = 0;
            for (int r = 0; r < rg.nread; ++r) {
                out[r].score = 0; out[r].pos = -1;
                const char* read = text + read_start[rg.read_first + r];
                const int n = read_len[rg.read_first + r];
                if (n <= K || L < K) continue;
                // per start: earliest trigger (t, o); starts are 0 .. L - n
                struct Ev { int t, o; };
                std::vector<Ev> ev((size_t)std::max(L - n + 1, 0), Ev{-1, -1});
                for (int d = -(n - K); d <= L - K; ++d) {
                    int run = 0;
                    for (int x = std::max(0, -d); x < n && d + x < L; ++x) {
                        run = (read[x] == hap[d + x]) ? run + 1 : 0;
                        if (run >= K) {
                            const int o = x - K + 1, i = d + o;
                            occ[(size_t)i] = 1;
                            const int s = std::max(0, d);
                            if (s + n <= L) { Ev& e = ev[(size_t)s]; if (e.t < 0 || i < e.t || (i == e.t && o < e.o)) e = Ev{i, o}; }
                        }
                    }
                }
                int best = 0, bt = 0, bo = 0, bpos = -1;
                for (int s = 0; s + n <= L; ++s) {
                    if (ev[(size_t)s].t < 0) continue;
                    int mism = 0;
                    for (int x = 0; x < n; ++x) if (read[x] != hap[s + x] && read[x] != 'N' && hap[s + x] != 'N') ++mism;
                    if (mism > 2) continue;
                    const int sc = (n - mism) * 4 - mism * 6, t = ev[(size_t)s].t, o = ev[(size_t)s].o;
                    for (int p = s; p < s + n; ++p) cov[(size_t)p] = std::min(cov[(size_t)p], t);
                    if (sc > best || (sc == best && (t < bt || (t == bt && o < bo)))) { best = sc; bt = t; bo = o; bpos = s; }
                }
                if (best > 0) { out[r].score = best; out[r].pos = bpos; total += best; }
            }
            bool drop = false;
            if (!hap_is_ref[rg.hap_first + h])
                for (int i = 0; i + K <= L; ++i)
                    if (i >= rg.prefix && (size_t)i < (size_t)L - (size_t)rg.suffix && occ[(size_t)i] && cov[(size_t)i] > i) drop = true;
            hap_score[rg.hap_first + h] = drop ? 0 : (int32_t)total;
        });
    }
    return 0;
}

extern "C" float mpn_fastpass_last_kernel_ms(const mpn_engine*) { return 0.f; }
