/*
 * oracle.h -- C API of the CPU oracle for the Sailfish quantification hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in sailfish_b200/ (the product) may include, link or
 * call this.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs use it, and only as the checker / the CPU arm.
 *
 * The oracle restates, in plain C++ on the CPU, the algorithm of the reference
 * (kingsfordgroup/sailfish v0.10.0, paths below are relative to /root/reference):
 *
 *   orc_xxh64            src/xxhash.c:346-455  (XXH64)                       PINNED (oracle/_ref + KATs)
 *   orc_compat_* / orc_hit_type / orc_parse_libtype
 *                        src/SailfishUtils.cpp:63-97,157-289                 PINNED (tests/LibraryTypeTests.cpp truth tables)
 *   eq-class counting    include/EquivalenceClassBuilder.hpp:64-108,
 *                        src/TranscriptGroup.cpp:9-19,53-55                  PINNED (oracle/_ref runs the real builder)
 *   hit -> label logic   src/SailfishQuantify.cpp:215-439 (PE), 530-631 (SE) restated; PARITY UNPINNED (no reference test)
 *   quasi-mapping        RapMap sf-v0.10.1 (NOT in /root/reference)          PARITY UNPINNED: specified by DESIGN.md "mapping spec v1"
 *   effective lengths    src/SailfishQuantify.cpp:648-838,937-992,1034-1043  restated; PARITY UNPINNED
 *   EM / VBEM            src/CollapsedEMOptimizer.cpp:33-44,93-217,711-893   PINNED when oracle/_ref/libsfref_em.so built
 *                                                                            (the reference TU itself, compiled against stub
 *                                                                            headers for TBB/Boost/RapMap), else UNPINNED
 *   bootstrap            src/CollapsedEMOptimizer.cpp:438-525,557-709,
 *                        include/MultinomialSampler.hpp:13-64                restated (RNG differs: distributional parity)
 *   Gibbs                src/CollapsedGibbsSampler.cpp:35-186,199-270        restated (RNG differs: distributional parity)
 *   TPM                  src/GZipWriter.cpp:194-248                          restated
 */
#ifndef SFB200_ORACLE_H
#define SFB200_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- A5: hashing ------------------------------------------------------------------------- */
uint64_t orc_xxh64(const void* data, size_t len, uint64_t seed);

/* ---- A4: library-format compatibility ------------------------------------------------------
 * format id = LibraryFormat::formatID() (include/LibraryFormat.hpp:89-98):
 *   bit0 type (0 SE,1 PE) | bits1-2 orientation (0 SAME,1 AWAY,2 TOWARD,3 NONE) | bits3-5 strandedness
 *   (0 SA,1 AS,2 S,3 A,4 U).   mate status: 0 SINGLE_END, 1 PAIRED_END_LEFT, 2 PAIRED_END_RIGHT, 3 PAIRED_END_PAIRED */
int orc_parse_libtype(const char* s);                       /* "IU","ISF",... -> format id, -1 if unknown */
int orc_compat_single(int expected_fmt, int32_t start, int is_fwd, int mate_status);
int orc_compat_paired(int expected_fmt, int observed_fmt);
int orc_hit_type(int32_t end1_start, int end1_fwd, uint32_t len1, int32_t end2_start, int end2_fwd,
                 uint32_t len2, int can_dovetail);           /* -> observed format id */

/* ---- index (mapping spec v1) ---------------------------------------------------------------- */
typedef struct orc_index orc_index;
/* seq: concatenated transcript sequences (ASCII), transcript t = seq[txp_off[t] .. txp_off[t]+txp_len[t]) */
orc_index* orc_index_build(const char* seq, const uint64_t* txp_off, const uint32_t* txp_len,
                           uint32_t n_txp, int k, int n_threads);
/* wrap an existing packed text + suffix order (positions sorted by (k-mer value, position)); used by bench.py's CPU arm */
orc_index* orc_index_from_arrays(const uint64_t* words, uint64_t text_len, const uint32_t* txp_len,
                                 uint32_t n_txp, int k, const uint32_t* sa_pos, uint64_t n_sa);
/* same with transcript ids and the k-mer table supplied as well: table16 = n_slots x {k-mer u64, first entry u32, count u32},
 * empty = all-ones k-mer, slot = XXH64(k-mer) & (n_slots-1), linear probing */
orc_index* orc_index_from_table(const uint64_t* words, uint64_t text_len, const uint32_t* txp_len, uint32_t n_txp, int k,
                                const uint32_t* sa_pos, const uint32_t* sa_tid, uint64_t n_sa, const uint64_t* table16,
                                uint64_t n_slots);
void orc_index_free(orc_index*);
/* per-hit pieces of the bias sample collection: ReadKmerDist<6>::update's bin for a hit (-1 = no sample), Transcript::gcFrac */
int32_t orc_bias_context_index(const orc_index*, uint32_t tid, int32_t pos, int fwd, uint32_t read_len);
int32_t orc_gc_frac(const orc_index*, uint32_t tid, int32_t s, int32_t e);
uint64_t orc_index_n_sa(const orc_index*);       /* number of valid suffix positions */
uint64_t orc_index_n_kmers(const orc_index*);    /* number of distinct k-mers */
uint64_t orc_index_text_len(const orc_index*);   /* packed coordinate space length */
/* copy out: sa_pos[n_sa] (packed coordinate), sa_tid[n_sa]; kmers[n_kmers] ascending, lb[n_kmers], cnt[n_kmers] */
void orc_index_export(const orc_index*, uint32_t* sa_pos, uint32_t* sa_tid, uint64_t* kmers, uint32_t* lb, uint32_t* cnt);
/* packed 2-bit text words (32 bases / u64, base i at bits 2*(i%32)); n_words = ceil(text_len/32)+1 */
void orc_index_export_text(const orc_index*, uint64_t* words);

/* ---- mapping + eq-class counting -------------------------------------------------------------- */
typedef struct {
    uint32_t max_read_occs;      /* --maxReadOcc, default 200            (SailfishQuantify.cpp:217,533) */
    uint32_t max_frag_len;       /* --maxFragLen, default 1000           (:115,427) */
    int32_t  num_frag_samples;   /* --numFragSamples, default 10000      (:901) */
    int32_t  lib_format_id;      /* LibraryFormat::formatID of -l        */
    int32_t  strict_intersect;   /* --strictIntersect                    (:204) */
    int32_t  allow_orphans;      /* !--discardOrphans                    (:139,226) */
    int32_t  allow_dovetail;     /* --allowDovetail                      (:165) */
    int32_t  ignore_compat;      /* --ignoreLibCompat                    (:156) */
    int32_t  enforce_compat;     /* --enforceLibCompat                   (:160) */
    uint32_t max_interval;       /* mapping spec v1: k-mer buckets larger than this are ignored (1000) */
} orc_map_opts;

typedef struct orc_run orc_run;
orc_run* orc_run_create(const orc_index*, const orc_map_opts*);
void orc_run_free(orc_run*);
void orc_run_keep_labels(orc_run*, int on);   /* debug: remember the label of every read of the last batch */
/* bases: concatenated ASCII; read i = bases[off[i]..off[i+1]).  bases2/off2 NULL for single-end.
 * Reads are processed in global order (== reference at -p 1); n_threads>1 splits the batch into
 * contiguous chunks (FLD sampling stays by global read index). */
int orc_map_batch(orc_run*, const char* bases1, const uint64_t* off1, const char* bases2, const uint64_t* off2,
                  uint64_t n_reads, int n_threads);
/* counters: [0] numObservedFragments [1] numMappedFragments [2] numFragHits [3] upperBoundHits [4] numFwd [5] numRC */
int orc_map_finish(orc_run*, uint64_t counters[6], uint32_t* fld_hist /*max_frag_len*/, uint64_t* n_classes, uint64_t* nnz);
/* --biasCorrect / --gcBiasCorrect sample collection (SailfishQuantify.cpp:255-287,372-389,555-583); call before orc_map_batch */
void orc_run_set_bias(orc_run*, int seq_bias, int gc_bias, int32_t num_bias_samples);
int orc_map_finish_bias(const orc_run*, uint32_t* read_bias /*4096, pseudo-count 1 included*/, uint32_t* observed_gc /*101*/);
/* classes in canonical order (label-lexicographic): row_ptr[E+1], labels[nnz], counts[E] */
int orc_eq_export(const orc_run*, uint64_t* row_ptr, uint32_t* labels, uint64_t* counts);
/* mapping work counters for B_map: [0] table probes P [1] SA entries S [2] text bases compared X */
void orc_map_work(const orc_run*, uint64_t work[3]);
/* per-read debug: label of the last batch's read i (returns length, writes up to cap ids); -1 if unmapped */
int orc_last_label(const orc_run*, uint64_t i, uint32_t* out, int cap);

/* ---- effective lengths (A9) ------------------------------------------------------------------- */
/* mode 0: smoothed from fld_hist if n_samples_seen >= num_frag_samples (remainingFLOps<=0), else Gaussian prior;
 *         single_end!=0 forces the prior;  mode 1: --noEffectiveLengthCorrection;  mode 2: --unsmoothedFLD */
int orc_eff_lens(const uint32_t* txp_len, uint32_t n_txp, const uint32_t* fld_hist, uint32_t max_frag_len,
                 int32_t num_frag_samples, int single_end, int mode, double prior_mean, double prior_sd,
                 double* eff_out);

/* ---- EM / VBEM (A10-A14) ---------------------------------------------------------------------- */
typedef struct {
    int32_t  use_vb;          /* --useVBOpt */
    double   prior_alpha;     /* 0.01  (CollapsedEMOptimizer.cpp:786) */
    double   tol;             /* 0.01  (SailfishQuantify.cpp:1343) */
    uint32_t min_iter;        /* 50    (CollapsedEMOptimizer.cpp:716) */
    uint32_t max_iter;        /* 10000 (SailfishQuantify.cpp:1343) */
    uint32_t fixed_iters;     /* >0: run exactly this many iterations, ignore convergence (bench / parity switch) */
    double   check_cutoff;    /* 1e-2  (:811) */
    double   min_alpha;       /* 1e-8  (:810) */
} orc_em_opts;
void orc_em_default_opts(orc_em_opts*);
/* returns 0 ok, -1 no active transcripts (:794-798), -2 alpha sum too small (:877-881) */
int orc_em_run(uint32_t n_txp, uint64_t n_classes, const uint64_t* row_ptr, const uint32_t* labels,
               const uint64_t* counts, const double* eff_lens, uint64_t num_mapped, const orc_em_opts*,
               int n_threads, double* alphas_out, uint32_t* iters_out, double* max_rel_diff_out);
/* GZipWriter.cpp:194-248 numerics */
void orc_tpm(uint32_t n_txp, const double* alphas, const double* eff_lens, uint64_t num_mapped, double* tpm_out);
double orc_digamma(double x);

/* ---- bootstrap (A16) / Gibbs (A17) ------------------------------------------------------------- */
typedef int (*orc_f64_row_cb)(void* user, const double* row, size_t n);
typedef int (*orc_i32_row_cb)(void* user, const int32_t* row, size_t n);
int orc_bootstrap(uint32_t n_txp, uint64_t n_classes, const uint64_t* row_ptr, const uint32_t* labels,
                  const uint64_t* counts, const double* eff_lens, const orc_em_opts*, uint32_t n_boot,
                  uint64_t seed, orc_f64_row_cb cb, void* user);
/* doBootstrap's inner loop (:476-514) on caller-supplied resampled counts: no min_iter, convergence gate on the
 * previous alpha (:499).  Lets a test feed the SAME resampled counts to the oracle and the GPU. */
int orc_bootstrap_em(uint32_t n_txp, uint64_t n_classes, const uint64_t* row_ptr, const uint32_t* labels,
                     const uint64_t* samp_counts, const double* eff_lens, const orc_em_opts*,
                     double* alphas_out, uint32_t* iters_out);
/* masses = Transcript::mass() after optimize (alpha/sum alpha) */
int orc_gibbs(uint32_t n_txp, uint64_t n_classes, const uint64_t* row_ptr, const uint32_t* labels,
              const uint64_t* counts, const double* eff_lens, const double* masses, uint64_t num_mapped,
              uint32_t n_samples, uint64_t seed, orc_i32_row_cb cb, void* user);

#ifdef __cplusplus
}
#endif
#endif
