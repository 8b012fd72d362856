// TEST DOUBLE for libsfb200 (tests/test_host_quant_cli.py::test_driver_host_flow_with_stub_device): the entry points sfb200-quant
// calls, returning canned device results and logging the calls to the file named by $SFB200_STUB_LOG, so that the driver's HOST
// flow (option handling, effective lengths, FLD hand-over, output files) runs in the CPU suite.  It computes nothing and is
// never linked into anything that ships.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "../include/sfb200.h"

struct sfb200_ctx { uint32_t T; uint32_t max_frag_len; uint64_t reads; };

static void logf(const char* fmt, ...) __attribute__((format(printf, 1, 2)));
#include <cstdarg>
static void logf(const char* fmt, ...) {
    const char* p = getenv("SFB200_STUB_LOG");
    if (!p) return;
    FILE* f = fopen(p, "a");
    va_list ap; va_start(ap, fmt); vfprintf(f, fmt, ap); va_end(ap);
    fputc('\n', f); fclose(f);
}

extern "C" {
int sfb200_ctx_create(int, sfb200_ctx** out) { *out = new sfb200_ctx{0, 1000, 0}; return SFB200_OK; }
void sfb200_ctx_destroy(sfb200_ctx* c) { delete c; }
void* sfb200_host_alloc(size_t bytes) { return std::malloc(bytes ? bytes : 1); }
void sfb200_host_free(void* p) { std::free(p); }
int sfb200_bind_host_near_device(int) { return 0; }
const char* sfb200_last_error(const sfb200_ctx*) { return "stub"; }
int sfb200_index_build(sfb200_ctx* c, const char*, const uint64_t*, const uint32_t*, uint32_t n, int k) { c->T = n; logf("index_build %u %d", n, k); return SFB200_OK; }
int sfb200_map_begin(sfb200_ctx* c, const sfb200_map_opts* o) { c->max_frag_len = o->max_frag_len; c->reads = 0; logf("map_begin %d", o->lib_format_id); return SFB200_OK; }
int sfb200_index_save(sfb200_ctx*, const char* p) { logf("index_save %s", p); FILE* f = fopen(p, "wb"); if (f) fclose(f); return SFB200_OK; }
int sfb200_index_load(sfb200_ctx*, const char* p) { logf("index_load %s", p); return SFB200_OK; }
uint64_t sfb200_map_clipped(sfb200_ctx*) { return getenv("SFB200_STUB_CLIPPED") ? 7 : 0; }
int sfb200_map_set_bias(sfb200_ctx*, int s, int g, int32_t n) { logf("map_set_bias %d %d %d", s, g, n); return SFB200_OK; }
int sfb200_map_batch(sfb200_ctx* c, const char*, const uint64_t*, const char* b2, const uint64_t*, uint64_t n) {
    if (getenv("SFB200_STUB_FAIL_MAP")) { logf("map_batch FAILS"); return SFB200_EFULL; }       // a device error in the middle of the run
    c->reads += n; logf("map_batch %llu %d", (unsigned long long)n, b2 ? 1 : 0); return SFB200_OK; }
int sfb200_map_batch_fixed(sfb200_ctx* c, const char*, uint32_t, const char* b2, uint32_t, uint64_t n) {
    if (getenv("SFB200_STUB_FAIL_MAP")) { logf("map_batch FAILS"); return SFB200_EFULL; }
    c->reads += n; logf("map_batch %llu %d", (unsigned long long)n, b2 ? 1 : 0); return SFB200_OK; }
// records = complete four-line groups; *consumed = the bytes the first n of them cover
static uint64_t stub_records(const char* t, uint64_t n) { uint64_t nl = 0; for (uint64_t i = 0; i < n; ++i) nl += t[i] == '\n'; return nl / 4; }
static uint64_t stub_consumed(const char* t, uint64_t n, uint64_t recs) { uint64_t nl = 0; for (uint64_t i = 0; i < n; ++i) if (t[i] == '\n' && ++nl == 4 * recs) return i + 1; return 0; }
int sfb200_map_fastq(sfb200_ctx* c, const char* t1, uint64_t n1, const char* t2, uint64_t n2, uint64_t max_records, uint64_t* n_records, uint64_t* c1, uint64_t* c2) {
    uint64_t n = stub_records(t1, n1);
    if (t2) { const uint64_t m = stub_records(t2, n2); if (m < n) n = m; }
    if (max_records && n > max_records) n = max_records;
    *n_records = n; *c1 = stub_consumed(t1, n1, n);
    if (t2) *c2 = stub_consumed(t2, n2, n);
    for (uint64_t i = 0, line = 0; n && i < *c1; ++i) {                       // the driver must hand over text that starts at a record boundary
        if (line % 4 == 0 && (i == 0 || t1[i - 1] == '\n') && t1[i] != '@') { logf("map_fastq BAD_START at %llu", (unsigned long long)i); return SFB200_EINVAL; }
        if (t1[i] == '\n') ++line;
    }
    c->reads += n;
    logf("map_fastq %llu %llu %llu paired=%d", (unsigned long long)n, (unsigned long long)*c1, (unsigned long long)(t2 ? *c2 : 0), t2 ? 1 : 0);
    return SFB200_OK;
}
int sfb200_map_finish(sfb200_ctx* c, uint64_t counters[6], uint32_t* fld, uint64_t* E, uint64_t* nnz) {
    const uint64_t v[6] = {c->reads, 3, 5, 4, 2, 1};
    std::memcpy(counters, v, sizeof v);
    for (uint32_t i = 0; i < c->max_frag_len; ++i) fld[i] = (i >= 180 && i < 220) ? 300 : 0;      // 12000 sampled fragment lengths
    *E = 2; *nnz = 3;
    return SFB200_OK;
}
int sfb200_map_get_bias(sfb200_ctx*, uint32_t* rb, uint32_t* og) { for (uint32_t i = 0; i < 4096; ++i) rb[i] = i + 1; for (uint32_t i = 0; i < 101; ++i) og[i] = i + 1; return SFB200_OK; }
int sfb200_eq_export(sfb200_ctx*, uint64_t* rp, uint32_t* lab, uint64_t* cnt) { rp[0] = 0; rp[1] = 1; rp[2] = 3; lab[0] = 0; lab[1] = 0; lab[2] = 1; cnt[0] = 2; cnt[1] = 1; return SFB200_OK; }
void sfb200_em_default_opts(sfb200_em_opts* o) { std::memset(o, 0, sizeof *o); o->prior_alpha = 0.01; o->tol = 0.01; o->min_iter = 50; o->max_iter = 10000; o->check_cutoff = 1e-2; o->min_alpha = 1e-8; }
static void alphas(uint32_t n, double* a) { for (uint32_t i = 0; i < n; ++i) a[i] = i == 0 ? 2.0 : i == 1 ? 1.0 : 0.0; }
int sfb200_em_run(sfb200_ctx*, const double* eff, uint32_t n, uint64_t nm, const sfb200_em_opts* o, double* a, uint32_t* it, double* mrd) {
    logf("em_run %llu vb=%d eff0=%.6f eff1=%.6f", (unsigned long long)nm, o->use_vb, eff[0], n > 1 ? eff[1] : 0.0);
    alphas(n, a);
    if (it) *it = 51;
    if (mrd) *mrd = 0.001;
    return SFB200_OK;
}
int sfb200_em_run_bias(sfb200_ctx*, const double* eff, uint32_t n, uint64_t nm, const sfb200_em_opts*, const sfb200_bias_model* m, double* a, double* eff_out,
                       uint32_t* it, double* mrd) {
    logf("em_run_bias %llu mode=%d gc_samp=%u fwd=%lld rc=%lld n_cdf=%u fld_max=%u rb7=%u og100=%u cdf_last=%.6f", (unsigned long long)nm, m->mode, m->gc_samp,
         (long long)m->num_fwd, (long long)m->num_rc, m->n_cdf, m->fld_max, m->read_bias[7], m->observed_gc[100], m->n_cdf ? m->fld_cdf[m->n_cdf - 1] : -1.0);
    alphas(n, a);
    for (uint32_t i = 0; i < n; ++i) eff_out[i] = 0.5 * eff[i];
    if (it) *it = 77;
    if (mrd) *mrd = 0.002;
    return SFB200_OK;
}
int sfb200_bootstrap_run(sfb200_ctx*, const double*, uint32_t n, const sfb200_em_opts*, uint32_t nb, uint64_t, sfb200_f64_row_cb cb, void* u) {
    std::string row(n * sizeof(double), 0);
    double* r = reinterpret_cast<double*>(&row[0]);
    for (uint32_t b = 0; b < nb; ++b) { for (uint32_t i = 0; i < n; ++i) r[i] = b + 1.0; if (cb(u, r, n)) return SFB200_ECALLBACK; }
    return SFB200_OK;
}
int sfb200_gibbs_run(sfb200_ctx*, const double*, const double*, uint32_t n, uint64_t, uint32_t ns, uint64_t, sfb200_i32_row_cb cb, void* u) {
    std::string row(n * sizeof(int32_t), 0);
    int32_t* r = reinterpret_cast<int32_t*>(&row[0]);
    for (uint32_t s = 0; s < ns; ++s) { for (uint32_t i = 0; i < n; ++i) r[i] = (int32_t)s; if (cb(u, r, n)) return SFB200_ECALLBACK; }
    return SFB200_OK;
}
}
