// Legacy per-pair ABI (include/ssw.h) on top of the batched engine: ssw_init / ssw_align / init_destroy / align_destroy
// with the reference's names, ownership and error behaviour (reference ssw.c:733-857).  Each ssw_align is a batch of one
// on the GPU; there is deliberately no CPU code path here.
#include "../../include/ssw.h"
#include "../../include/mpn_ssw_batch.h"
#include <cstdlib>
#include "host_shared.h"
#include <cstdio>

struct _profile {
    const int8_t* read;       // borrowed (reference ssw.c:749)
    const int8_t* mat;        // borrowed (reference ssw.c:750)
    int32_t readLen;
    int32_t n;
    int8_t score_size;
};

namespace mpn {
std::mutex& shared_engine_mutex() { static std::mutex mu; return mu; }
mpn_engine* shared_engine_locked()
{
    static mpn_engine* engine = nullptr;
    if (!engine) {
        engine = mpn_engine_create(-1);
        if (!engine) {
            fprintf(stderr, "[libssw] cannot create the GPU engine: this library has no CPU fallback\n");
            abort();
        }
    }
    return engine;
}
}  // namespace mpn

extern "C" s_profile* ssw_init(const int8_t* read, const int32_t readLen, const int8_t* mat, const int32_t n, const int8_t score_size)
{
    s_profile* p = (s_profile*)calloc(1, sizeof(struct _profile));
    p->read = read; p->mat = mat; p->readLen = readLen; p->n = n; p->score_size = score_size;
    return p;
}

extern "C" void init_destroy(s_profile* p) { free(p); }

extern "C" s_align* ssw_align(const s_profile* prof, const int8_t* ref, int32_t refLen, const uint8_t weight_gapO, const uint8_t weight_gapE,
                              const uint8_t flag, const uint16_t filters, const int32_t filterd, const int32_t maskLen)
{
    if (maskLen < 15)      // reference ssw.c:782-784
        fprintf(stderr, "When maskLen < 15, the function ssw_align doesn't return 2nd best alignment information.\n");
    if (prof->score_size < 0 || prof->score_size > 2) {      // no profile was built (reference ssw.c:801-804)
        fprintf(stderr, "Please call the function ssw_init before ssw_align.\n");
        return NULL;
    }
    mpn_params pr;
    pr.mat = prof->mat; pr.n = prof->n; pr.gapO = weight_gapO; pr.gapE = weight_gapE; pr.score_size = prof->score_size;
    pr.flag = flag; pr.filters = filters; pr.filterd = filterd;
    const int64_t read_off[2] = {0, prof->readLen}, ref_off[2] = {0, refLen};
    const int32_t mask[1] = {maskLen};
    mpn_result res;
    const int64_t cap = 2ll * ((int64_t)prof->readLen + refLen) + 16;
    uint32_t* cig = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)cap);
    int rc;
    {
        mpn::SharedEngineLock lk;
        rc = mpn_align_batch(lk.engine(), &pr, prof->read, read_off, ref, ref_off, mask, 1, &res, cig, cap);
    }
    if (rc != 0) {
        fprintf(stderr, "[libssw] GPU alignment failed (code %d); no CPU fallback\n", rc);
        abort();
    }
    if (res.status != MPN_ST_OK) {
        // the cases in which the reference returns NULL: 8-bit overflow without a 16-bit profile (ssw.c:793-796), traceback error (ssw.c:840-843)
        if (res.status == MPN_ST_NULL && prof->score_size == 0)
            fprintf(stderr, "Please set 2 to the score_size parameter of the function ssw_init, otherwise the alignment results will be incorrect.\n");
        free(cig);
        return NULL;
    }
    s_align* r = (s_align*)calloc(1, sizeof(s_align));
    r->score1 = res.score1; r->score2 = res.score2;
    r->ref_begin1 = res.ref_begin1; r->ref_end1 = res.ref_end1;
    r->read_begin1 = res.read_begin1; r->read_end1 = res.read_end1; r->ref_end2 = res.ref_end2;
    if (res.cigar_len > 0) {
        if (res.cigar_off) memmove(cig, cig + res.cigar_off, sizeof(uint32_t) * (size_t)res.cigar_len);
        r->cigar = cig; r->cigarLen = res.cigar_len;
    } else {
        free(cig);
        r->cigar = NULL; r->cigarLen = 0;
    }
    return r;
}

extern "C" void align_destroy(s_align* a)
{
    if (!a) return;
    free(a->cigar);
    free(a);
}
