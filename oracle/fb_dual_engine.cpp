// fb_dual_engine.cpp -- TEST INFRASTRUCTURE: a C-ABI engine that forwards every call to BOTH the CUDA
// library and the CPU oracle library (dlopen, RTLD_LOCAL) and compares every work-item result.  The host
// program linked against it (oracle/_build/fillgaps_dual) therefore checks the device engine item by item
// on real inputs, and reports the first item whose discrete outputs differ or whose weights drift > 1e-5.
// Results returned to the caller are the CUDA engine's.  Used only by -m gpu tests / debugging.
#include <dlfcn.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../include/figbird_b200.h"

namespace {
struct Api {
    void* h = nullptr;
    fb_status (*ctx_create)(int32_t, fb_ctx**);
    void (*ctx_destroy)(fb_ctx*);
    const char* (*last_error)(const fb_ctx*);
    fb_status (*model_upload)(fb_ctx*, const FbModel*);
    fb_status (*batch_upload)(fb_ctx*, const FbGapBatch*);
    fb_status (*em_run)(fb_ctx*, const FbWorkItem*, int32_t, const FbItemOut**);
    fb_status (*get_counters)(const fb_ctx*, FbCounters*);
    fb_status (*microbench)(fb_ctx*, double*);
    fb_status (*set_latency)(fb_ctx*, int32_t);
    bool load(const char* path) {
        h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
        if (!h) { fprintf(stderr, "fb_dual: dlopen %s: %s\n", path, dlerror()); return false; }
#define SYM(f, n) *(void**)(&f) = dlsym(h, n); if (!f) { fprintf(stderr, "fb_dual: %s lacks %s\n", path, n); return false; }
        SYM(ctx_create, "fb_ctx_create") SYM(ctx_destroy, "fb_ctx_destroy") SYM(last_error, "fb_last_error") SYM(model_upload, "fb_model_upload")
        SYM(batch_upload, "fb_batch_upload") SYM(em_run, "fb_em_run") SYM(get_counters, "fb_get_counters") SYM(microbench, "fb_microbench_fp64") SYM(set_latency, "fb_ctx_set_latency_critical")
#undef SYM
        return true;
    }
};
Api& dev() { static Api a; return a; }
Api& ora() { static Api a; return a; }
bool init() {
    static int ok = -1;
    if (ok < 0) {
        const char* d = getenv("FB_DUAL_DEVICE_LIB"); const char* o = getenv("FB_DUAL_ORACLE_LIB");
        ok = (d && o && dev().load(d) && ora().load(o)) ? 1 : 0;
    }
    return ok == 1;
}
long g_mismatch = 0, g_items = 0;
double g_worst = 0;      // largest relative deviation of a count-matrix entry (per-position base weight) seen so far
// Summary for the tests: written at process exit (engine contexts are pooled by the host program and never destroyed)
struct Summary { ~Summary() {
    if (const char* p = getenv("FB_DUAL_SUMMARY")) { FILE* f = fopen(p, "w"); if (f) { fprintf(f, "{\"items\": %ld, \"mismatching\": %ld, \"worst_weight_rel\": %.3e}\n", g_items, g_mismatch, g_worst); fclose(f); } }
    fprintf(stderr, "fb_dual: %ld items compared, %ld mismatching, worst relative weight deviation %.3e\n", g_items, g_mismatch, g_worst);
} } g_summary;
}  // namespace

struct fb_ctx { fb_ctx* d; fb_ctx* o; std::string err; int cutoff = 0; };

extern "C" const char* fb_engine_name(void) { return "dual(cuda-sm100a|oracle-cpu)"; }
extern "C" fb_status fb_ctx_create(int32_t device, fb_ctx** out) {
    if (!init()) return FB_ERR_STATE;
    fb_ctx* c = new fb_ctx(); c->d = nullptr; c->o = nullptr;
    *out = c;
    fb_status s = dev().ctx_create(device, &c->d);
    if (s != FB_OK) { c->err = "device engine: ctx_create failed"; return s; }
    return ora().ctx_create(0, &c->o);
}
extern "C" void fb_ctx_destroy(fb_ctx* c) {
    if (!c) return;
    if (c->d) dev().ctx_destroy(c->d);
    if (c->o) ora().ctx_destroy(c->o);
    delete c;
}
extern "C" const char* fb_last_error(const fb_ctx* c) { return c ? (c->err.empty() ? dev().last_error(c->d) : c->err.c_str()) : "null"; }
extern "C" fb_status fb_model_upload(fb_ctx* c, const FbModel* m) { c->cutoff = m->prob_cutoff; fb_status s = dev().model_upload(c->d, m); return s ? s : ora().model_upload(c->o, m); }
extern "C" fb_status fb_batch_upload(fb_ctx* c, const FbGapBatch* b) { fb_status s = dev().batch_upload(c->d, b); return s ? s : ora().batch_upload(c->o, b); }
extern "C" fb_status fb_get_counters(const fb_ctx* c, FbCounters* o) { return dev().get_counters(c->d, o); }
extern "C" fb_status fb_microbench_fp64(fb_ctx* c, double* o) { return dev().microbench(c->d, o); }
extern "C" fb_status fb_ctx_set_latency_critical(fb_ctx* c, int32_t on) { return dev().set_latency(c->d, on); }
extern "C" int32_t fb_fillgaps_main(int32_t, const char* const*);   // provided by the host sources linked into this library

extern "C" fb_status fb_em_run(fb_ctx* c, const FbWorkItem* items_in, int32_t n, const FbItemOut** out) {
    // ask both engines for the count matrices of every EM item so that weights are compared too
    static thread_local std::vector<FbWorkItem> forced;
    forced.assign(items_in, items_in + n);
    for (auto& w : forced) if (w.kind == FB_ITEM_EM) w.flags |= FB_FLAG_WANT_COUNTS;
    const FbWorkItem* items = forced.data();
    fb_status s = dev().em_run(c->d, items, n, out);
    if (s) return s;
    const FbItemOut** oo = (const FbItemOut**)malloc(sizeof(void*) * (n > 0 ? n : 1));
    s = ora().em_run(c->o, items, n, oo);
    if (s) { free(oo); c->err = "oracle engine failed"; return s; }
    for (int i = 0; i < n; i++) {
        const FbItemOut* A = oo[i]; const FbItemOut* B = out[i];
        const unsigned char* a = (const unsigned char*)A; const unsigned char* b = (const unsigned char*)B;
        std::string why; bool countsOnly = false;
        g_items++;
        if (A->calls != B->calls) why += " calls";
        if (A->comp_count != B->comp_count) why += " comp_count";
        if (A->flags != B->flags) why += " flags";
        if (A->placements != B->placements) why += " placements";
        const int R = A->n_reads, Lg = A->cand_len, slots = A->n_slots > 1 ? A->calls : 1;
        if (why.empty()) {
            const double* ap2 = (const double*)(a + A->off_p2max); const double* bp2 = (const double*)(b + B->off_p2max);
            const int32_t* aps = (const int32_t*)(a + A->off_pos2); const int32_t* bps = (const int32_t*)(b + B->off_pos2);
            const double* ap1 = (const double*)(a + A->off_p1max); const double* bp1 = (const double*)(b + B->off_p1max);
            // EM items: p2max / pos2 are guaranteed only for accepted reads; a rejected read must be rejected by both engines
            const bool em = items[i].kind == FB_ITEM_EM; const double cut = (double)c->cutoff;
            auto accepted = [&](double p) { return p > 0 && -log10(p) < cut; };
            for (int k = 0; k < slots * R; k++) {
                if (em && !accepted(ap2[k])) {
                    if (accepted(bp2[k])) { char buf[200]; snprintf(buf, sizeof buf, " accept[%d] oracle %.17g (rejected) device %.17g", k, ap2[k], bp2[k]); why += buf; break; }
                } else {
                if (ap2[k] != bp2[k]) { char buf[200]; snprintf(buf, sizeof buf, " p2max[%d] oracle %.17g pos %d device %.17g pos %d", k, ap2[k], aps[k], bp2[k], bps[k]); why += buf; break; }
                if (aps[k] != bps[k]) { why += " pos2[" + std::to_string(k) + "]"; break; }
                }
                if ((ap1[k] > 0) != (bp1[k] > 0) || (ap1[k] > 0 && fabs(ap1[k] - bp1[k]) > 1e-9 * ap1[k])) { char buf[200]; snprintf(buf, sizeof buf, " p1max[%d] oracle %.17g device %.17g", k, ap1[k], bp1[k]); why += buf; break; }
            }
            if (memcmp(a + A->off_soft, b + B->off_soft, Lg)) {
                why += " soft";
                for (int x = 0; x < Lg; x++) if (a[A->off_soft + x] != b[B->off_soft + x]) {
                    why += "@row" + std::to_string(x);
                    if (A->off_counts >= 0 && B->off_counts >= 0) {
                        const double* ac = (const double*)(a + A->off_counts) + 5 * x; const double* bc = (const double*)(b + B->off_counts) + 5 * x;
                        char buf[400]; snprintf(buf, sizeof buf, " oracle[%.17g %.17g %.17g %.17g %.17g] device[%.17g %.17g %.17g %.17g %.17g]", ac[0], ac[1], ac[2], ac[3], ac[4], bc[0], bc[1], bc[2], bc[3], bc[4]);
                        why += buf;
                    }
                    break;
                }
            }
            if (memcmp(a + A->off_hard, b + B->off_hard, Lg)) why += " hard";
            if (memcmp(a + A->off_cov, b + B->off_cov, sizeof(int32_t) * Lg)) why += " cov";
            if (A->off_counts >= 0 && B->off_counts >= 0) {
                const double* ac = (const double*)(a + A->off_counts); const double* bc = (const double*)(b + B->off_counts);
                for (int k = 0; k < 5 * Lg; k++) if (ac[k] != 0) { const double rel = fabs(ac[k] - bc[k]) / fabs(ac[k]); if (rel > g_worst) g_worst = rel; }
                for (int k = 0; k < 5 * Lg; k++) if (fabs(ac[k] - bc[k]) > 1e-5 * fabs(ac[k])) {
                    char buf[200]; snprintf(buf, sizeof buf, " counts[row %d col %d] oracle %.17g device %.17g", k / 5, k % 5, ac[k], bc[k]);
                    if (why.empty()) { countsOnly = true; } why += buf; break; }
            }
        }
        if (!why.empty()) {
            g_mismatch++;
            static long nCounts = 0, nDisc = 0;
            long& cls = countsOnly ? nCounts : nDisc;
            if (++cls <= 12)
                fprintf(stderr, "fb_dual MISMATCH item kind=%d gap=%d Lg=%d rounds=%d flags=%d reads=%d: oracle calls=%d comp=%d | device calls=%d comp=%d :%s\n",
                        items[i].kind, items[i].gap, items[i].cand_len, items[i].max_rounds, items[i].flags, R, A->calls, A->comp_count, B->calls, B->comp_count, why.c_str());
        }
    }
    free(oo);
    return FB_OK;
}
